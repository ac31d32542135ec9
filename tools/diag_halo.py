"""Diagnostic: per-tap correctness of the halo-reuse conv kernel (single-tap identity weights)."""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import layers as L

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
H, W, n = 16, 16, 1
x = torch.randn(n, 64, H, W, generator=g)
xs = L.Split.from_nchw(x.to(dev))
b = torch.zeros(64)
print("mode", os.environ.get("ESR_TC_HALO_MODE", "0"), "v1" if os.environ.get("ESR_TC_V1") else "v3")
for tap in range(9):
    w = torch.zeros(64, 64, 3, 3)
    w[torch.arange(64), torch.arange(64), tap // 3, tap % 3] = 1.0
    want = F.conv2d(x, w, b, padding=1)
    out = L.Split(n, 64, H, W, dev)
    L.conv_tc([xs], L.pack_weight(w.to(dev)), L.pad_bias(b.to(dev), 64), 64, act=None, out=out)
    got = out.to_nchw().cpu()
    err = (got - want).abs()
    bad = (err > 1e-3).float()
    # which output rows / cols / channels are wrong
    print(f"tap {tap} (dy={tap//3},dx={tap%3}) rel={err.max().item()/want.abs().max().item():.3e} bad_frac={bad.mean().item():.3f}",
          "bad cols:", [int(v) for v in bad.sum((0, 1, 2)).tolist()], "bad ch/8:", [int(v) for v in bad.sum((0, 2, 3)).view(8, 8).sum(1).tolist()])
