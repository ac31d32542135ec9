"""Where the redistribution time goes: kernels of esr_cnt2event_fused (CUDA events) vs the whole expand() call (wall clock,
host sizing included), fused vs general chain.  Run on the GPU box: python tools/bench_expand.py"""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from esr_b200 import _lib, expand as ex   # noqa: E402

dev = torch.device("cuda:0")
L = _lib.lib()
g = torch.Generator(device=dev).manual_seed(1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
CASES = [(48, 256, 256, 0.3), (48, 256, 256, 1.0), (1, 256, 256, 0.76), (1, 256, 256, 7.6), (8, 1024, 1024, 0.3)]
for (B, H, W, lam) in CASES[:int(os.environ.get("NCASES", "99"))]:
    cnt = torch.poisson(torch.full((B, 2, H, W), lam, device=dev), generator=g)
    E = int(cnt.sum().item())
    mcap = min(64, 1 << int(2 * cnt.max().item() - 1).bit_length())
    tables, desc = ex._xf_tables(dev)
    cap = int(E * 1.3) + 4096 + B * 100000
    out = torch.empty((cap * 4,), dtype=torch.float32, device=dev)
    stats = torch.empty((B, 4), dtype=torch.int64, device=dev)
    nb = L.esr_cnt2event_fused_workspace_bytes(B, H, W)
    ws = torch.empty((nb,), dtype=torch.uint8, device=dev)

    def kern():
        _lib.check(L.esr_cnt2event_fused(_lib.ptr(cnt), B, H, W, _lib.ptr(tables), desc.ctypes.data_as(ctypes.c_void_p), mcap, _lib.ptr(stats),
                                         _lib.ptr(out), cap, _lib.ptr(ws), nb, _lib.stream_ptr()), "fused")
    res = {}
    for name, fn in (("kernels_fused", kern), ("expand_fused", lambda: ex.expand(cnt, 0, 0))):
        for _ in range(3):
            fn()
        ts, wall = [], []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            ts.append(e0.elapsed_time(e1))
        res[name] = (sorted(ts)[3], sorted(wall)[3])
    os.environ["ESR_EXPAND_FUSED"] = "0"
    fn = lambda: ex.expand(cnt, 0, 0)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    del os.environ["ESR_EXPAND_FUSED"]
    res["expand_general"] = (sorted(ts)[3], 0.0)
    mx = int(cnt.max().item())
    alg = 8.0 * H * W * B + 16.0 * E
    print(f"B={B} {H}x{W} lam={lam} events={E} max={mx} alg_MB={alg/1e6:.1f}: " +
          " | ".join(f"{k} {v[0]*1e3:.1f} us (wall {v[1]*1e3:.1f})  {alg/v[0]/1e6:.0f} GB/s" for k, v in res.items()), flush=True)
