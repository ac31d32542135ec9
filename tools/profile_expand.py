"""Kernel-level timing of the redistribution stage (esr_b200.expand) at a bench workload's shapes:
    python tools/profile_expand.py [cfg2] -> table of CUDA kernels (torch.profiler) + wall/GPU time per call."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from esr_b200.expand import expand  # noqa: E402

wl = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "cfg2"]
B, L, s = wl["B"], wl["L"], wl["scale"]
H, W = wl["lr"][0] * s, wl["lr"][1] * s
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
cnt = torch.poisson(torch.full(((L - 2) * B, 2, H, W), 0.3, device=dev), generator=g)
for _ in range(3):
    out = expand(cnt, 0, 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(20):
    out = expand(cnt, 0, 0)
e1.record()
torch.cuda.synchronize()
print(f"expand: {e0.elapsed_time(e1) / 20:.3f} ms GPU timeline, {(time.perf_counter() - t0) / 20 * 1e3:.3f} ms wall, events {int((out[..., 3] != 0).sum())}, shape {tuple(out.shape)}")
from torch.profiler import ProfilerActivity, profile
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    expand(cnt, 0, 0)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
