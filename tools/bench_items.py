"""Per-frame tensor factory: esr_b200.dataset.create_item on the GPU vs the CPU restatement of the
reference's H5Dataset.__getitem__ factory (oracle/items.py: C encodings + torch CPU F.interpolate), same synthetic events.
Prints one JSON line.   python tools/bench_items.py [--lr 128 128] [--scale 2] [--events 2048] [--iters 200]
Also times the batched path for the three tensors the trainer reads (dataset.collate_sequence: 3 launches per batch)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lr", type=int, nargs=2, default=[128, 128])
    ap.add_argument("--scale", type=int, default=2)
    ap.add_argument("--events", type=int, default=2048)
    ap.add_argument("--time-bins", type=int, default=1)
    ap.add_argument("--iters", type=int, default=200)
    args = ap.parse_args()
    from esr_b200 import dataset
    from oracle import items as oi
    H, W = args.lr
    rng = np.random.default_rng(0)

    def ev(n, h, w):
        return np.stack([rng.integers(0, w, n).astype(np.float64), rng.integers(0, h, n).astype(np.float64),
                         np.sort(rng.random(n)), rng.choice([-1.0, 1.0], n)])
    frames = [(ev(args.events, H, W), ev(args.events * args.scale ** 2, H * args.scale, W * args.scale)) for _ in range(16)]
    dev = torch.device("cuda:0")
    for i in range(10):
        dataset.create_item(*frames[i % 16], (H, W), args.scale, args.time_bins, device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(args.iters):
        item = dataset.create_item(*frames[i % 16], (H, W), args.scale, args.time_bins, device=dev)
    torch.cuda.synchronize()
    gpu_s = (time.perf_counter() - t0) / args.iters
    torch.set_num_threads(1)                       # one DataLoader worker = one core (more threads only slow these tiny ops down)
    n_cpu = max(10, args.iters // 10)
    oi.create_item(*frames[0], (H, W), args.scale, args.time_bins)
    t0 = time.perf_counter()
    for i in range(n_cpu):
        oi.create_item(*frames[i % 16], (H, W), args.scale, args.time_bins)
    cpu_s = (time.perf_counter() - t0) / n_cpu
    # batched: 8 sequences x 8 frames -> inp_cnt / inp_scaled_cnt / gt_cnt banks
    B, L = 8, 8
    inp = [[frames[(b * L + l) % 16][0] for l in range(L)] for b in range(B)]
    gt = [[frames[(b * L + l) % 16][1] for l in range(L)] for b in range(B)]
    for _ in range(3):
        dataset.collate_sequence(inp, gt, (H, W), (H * args.scale, W * args.scale), device=dev)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        dataset.collate_sequence(inp, gt, (H, W), (H * args.scale, W * args.scale), device=dev)
    torch.cuda.synchronize()
    col_s = (time.perf_counter() - t0) / 20
    print(json.dumps({"metric": "dataset items/sec (per-frame tensor factory, 20 tensors per item)", "value": 1.0 / gpu_s,
                      "unit": "items/s", "ms_per_item": gpu_s * 1e3,
                      "config": {"lr": [H, W], "scale": args.scale, "events_per_frame": args.events, "time_bins": args.time_bins,
                                 "data": "synthetic"},
                      "cpu_baseline": {"value": 1.0 / cpu_s, "unit": "items/s", "ms_per_item": cpu_s * 1e3, "kind": "port",
                                       "cores": 1, "sample": f"{n_cpu} items, one worker thread (the reference runs this in 4 DataLoader workers)"},
                      "collate_sequence": {"frames_per_s": B * L / col_s, "ms_per_batch": col_s * 1e3, "batch": [B, L],
                                           "note": "inp_cnt + inp_scaled_cnt + gt_cnt banks for a batch of sequences, 3 scatter launches, incl. host flattening + H2D"},
                      "note": "wall clock incl. H2D of the raw events; create_item keeps the reference's per-frame call structure (host-latency bound)"}))


if __name__ == "__main__":
    main()
