#!/usr/bin/env python
"""bench.py -- LR event-frames/sec of the ESR hot path on B200 (BASELINE.json metric), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2|cfg3] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" = one batch of B sequences x L LR event frames taken from raw events to redistributed SR events:
  encode L frames (LR->HR lift + count scatter) -> L-2 DeepRecurrNet forwards with carried ConvGRU state
  -> cnt2event of the L-2 SR count tensors.          (infer_ours_cnt.py:54-75)
value   : events already resident in HBM when a timed step starts.
e2e     : the same step through the public API with pinned HOST buffers: H2D of the events and D2H of the
          resulting event lists inside the timed region.
Multi-GPU: the path shards by batch with no data-path collective (inference); every rank runs the per-GPU batch of
the workload on its own shard (weak scaling) and the time is the max over ranks.
`--impl reference` times the CPU oracle port of the same path on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1]: 2x SR, seq_len=8, LR 128x128, batch=8 per GPU
    "cfg2": dict(scale=2, L=8, lr=(128, 128), B=8, desc="2x SR, seq_len=8, LR 128x128 synthetic events, batch=8/GPU"),
    # BASELINE.json configs[2]: 4x SR, seq_len=8, LR 128x128, batch=32 over 8 GPUs = 4 per GPU
    "cfg3": dict(scale=4, L=8, lr=(128, 128), B=4, desc="4x SR, seq_len=8, LR 128x128 synthetic events, batch=4/GPU"),
    # BASELINE.json configs[3]: 4x SR, seq_len=16, LR 256x256, batch=16 over 8 GPUs = 2 per GPU (long-sequence stress)
    "cfg4": dict(scale=4, L=16, lr=(256, 256), B=2, desc="4x SR, seq_len=16, LR 256x256 synthetic events, batch=2/GPU"),
}
EVENTS_PER_FRAME = 2048          # shipped WINDOW (config/train_ours_enfssyn.yml:9)
FLOP_PER_HR_PIXEL = 184.7e3


def synth_events(B, L, lr, seed):
    """SURVEY 8d synthetic input: n=2048 events per frame, x~U{0..W-1}, y~U{0..H-1}, p~U{-1,+1}."""
    g = torch.Generator().manual_seed(seed)
    n = B * L * EVENTS_PER_FRAME
    xs = torch.randint(0, lr[1], (n,), generator=g).float()
    ys = torch.randint(0, lr[0], (n,), generator=g).float()
    ps = (torch.randint(0, 2, (n,), generator=g) * 2 - 1).float()
    off = torch.arange(0, n + 1, EVENTS_PER_FRAME, dtype=torch.int64)
    return xs, ys, ps, off


def synth_weights(seed=0):
    """Random-init weights of the shipped architecture (fan-in scaled normal; non-zero conv_offset_mask so that the
    deformable sampling is exercised), keyed like the reference state_dict.  No checkpoint exists offline."""
    from esr_b200.model import DeepRecurrNet
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for k, v in DeepRecurrNet(inch=2, basech=8, num_frame=3).state_dict().items():
        shp = tuple(v.shape)
        if k.endswith(".weight"):
            fan_in = 1
            for d in shp[1:]:
                fan_in *= d
            std = 0.02 if "conv_offset_mask" in k else (1.0 / fan_in) ** 0.5
            sd[k] = torch.randn(shp, generator=g) * std
        else:
            sd[k] = torch.randn(shp, generator=g) * (0.3 if "conv_offset_mask" in k else 0.05)
    return sd


def synth_sr_bias(B, L, hr, seed):
    """A random-init network's output rounds to zero events, so the redistribution stage is fed
    `model output + Poisson(0.3)` synthetic counts to do representative work."""
    g = torch.Generator().manual_seed(seed + 17)
    return torch.poisson(torch.full(((L - 2) * B, 2, hr[0], hr[1]), 0.3), generator=g)


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons of ONE GPU during the timed region.  NVML in-process (nvidia_ml_py): a query costs
    microseconds; spawning `nvidia-smi` per sample from every rank (the first version) took the driver lock for ~0.5 s each
    and visibly slowed the timed steps at N >= 4.  Falls back to one nvidia-smi call per second if NVML is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu_index, [], False
        self.nvml, self.handle = None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(gpu_index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[gpu_index]) if vis and vis.split(",")[gpu_index].isdigit() else gpu_index
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _sample_nvml(self):
        n, h = self.nvml, self.handle
        sm = n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)
        mx = n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)
        try:
            bits = n.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            bits = n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        flag = lambda m: "Active" if bits & m else "Not Active"
        # bit masks (nvml.h): SwPowerCap 0x4, HwSlowdown 0x8, SwThermalSlowdown 0x20, HwThermalSlowdown 0x40
        self.rows.append([str(self.gpu), str(sm), str(mx), "", hex(bits), flag(0x8), flag(0x40), flag(0x20), flag(0x4)])

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self._sample_nvml()
                else:
                    out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.gpu)],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 1.0)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def usable_cores(cap=32):
    """Host threads the CPU legs may use: affinity and cgroup quota, capped (torch's small convs stop scaling, and
    oversubscribing a 128-way box made the oracle ~100x slower when measured)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        if q != "max":
            n = min(n, max(1, int(int(q) / int(per))))
    except Exception:
        pass
    return max(1, min(n, cap))


_JSON_OUT = sys.stdout


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1443.7), d.get("hbm_gbs", 6574.1), "measured (MEASURED_PEAKS.json)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------------------------
def cpu_oracle_step(wl, B_sample, sd, seed):
    """One step of the CPU oracle port on a sample of B_sample sequences.  Returns seconds."""
    from oracle import events as oe
    from oracle import model_ref
    scale, L, lr = wl["scale"], wl["L"], wl["lr"]
    hr = (lr[0] * scale, lr[1] * scale)
    xs, ys, ps, off = synth_events(B_sample, L, lr, seed)
    xs, ys, ps, off = xs.numpy(), ys.numpy(), ps.numpy(), off.numpy()
    bias = synth_sr_bias(B_sample, L, hr, seed)
    net = model_ref.OracleNet(sd)
    t0 = time.perf_counter()
    frames = np.empty((B_sample * L, 2, hr[0], hr[1]), np.float32)
    for f in range(B_sample * L):
        a, b = off[f], off[f + 1]
        frames[f] = oe.events_to_channels(oe.lift_coords(xs[a:b], lr[1], hr[1]), oe.lift_coords(ys[a:b], lr[0], hr[0]),
                                          ps[a:b], hr)
    bank = torch.from_numpy(frames).view(B_sample, L, 2, hr[0], hr[1])
    net.reset_states()
    outs = [net(bank[:, w:w + 3].contiguous()) for w in range(L - 2)]
    sr = torch.cat(outs, 0) + bias
    ev = oe.cnt2event(sr.numpy(), 0)
    dt = time.perf_counter() - t0
    return dt, int((ev[:, :, 3] != 0).sum())


def cpu_oracle_train_step(wl, sd, seed):
    """One training iteration (windows with carried state, summed MSE, backward, torch Adam amsgrad) of the CPU oracle on
    ONE sequence.  Returns seconds."""
    import torch.nn.functional as F
    from oracle import model_ref
    scale, L, lr = wl["scale"], wl["L"], wl["lr"]
    H, W = lr[0] * scale, lr[1] * scale
    g = torch.Generator().manual_seed(seed)
    frames = torch.poisson(torch.full((1, L, 2, H, W), 0.1), generator=g)
    gt = torch.poisson(torch.full((1, L, 2, H, W), 0.1), generator=g)
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    opt = torch.optim.Adam(list(params.values()), lr=1e-3, weight_decay=1e-4, amsgrad=True)
    t0 = time.perf_counter()
    opt.zero_grad()
    states, loss = None, 0
    for w in range(L - 2):
        pred, states = model_ref.forward(params, frames[:, w:w + 3], states)
        loss = loss + F.mse_loss(pred, gt[:, w + 1])
    loss.backward()
    opt.step()
    return time.perf_counter() - t0


def measure_training(args, wl, net_sd, dev, rank, world, dist):
    """the training iteration (forward of all windows, backward through time, gradient all-reduce
    when N > 1, Adam amsgrad) on the same workload; N = 1 replays one CUDA graph, N > 1 runs eagerly with an NCCL all-reduce
    of the flat 1.8 M-element gradient.  Device-resident synthetic count tensors (the trainer's input after the dataloader)."""
    from esr_b200 import train
    from esr_b200.model import DeepRecurrNet
    scale, L, lr, B = wl["scale"], wl["L"], wl["lr"], wl["B"]
    H, W = lr[0] * scale, lr[1] * scale
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    net.load_state_dict(net_sd)
    net = net.to(dev)
    opt = train.Adam(net.parameters(), lr=1e-3, weight_decay=1e-4, amsgrad=True)
    g = torch.Generator().manual_seed(200 + rank)
    frames = torch.poisson(torch.full((B, L, 2, H, W), 0.1), generator=g).to(dev)
    gt = torch.poisson(torch.full((B, L, 2, H, W), 0.1), generator=g).to(dev)
    if world > 1:
        def allred(flat):
            dist.all_reduce(flat)
            flat /= world
        step = lambda: train.train_step(net, opt, frames, gt, all_reduce=allred)
    else:
        gstep = train.GraphedTrainStep(net, opt, tuple(frames.shape), dev)
        step = lambda: gstep(frames, gt)
    steps = max(3, min(args.steps, 10))
    for _ in range(2):
        step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    first = last = None
    for i in range(steps):
        l = step()
        if i == 0:
            first = l.clone()
        last = l
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = t.item() / steps
    out = {"metric": "training LR event-frames/sec (forward + backward + Adam)", "value": world * B * L / (ms * 1e-3), "unit": "frames/s",
           "ms_per_step": ms, "steps": steps, "mode": "one CUDA graph per iteration" if world == 1 else "eager + NCCL all-reduce of the flat gradient",
           "batch_per_gpu": B, "loss_first": float(first), "loss_last": float(last),
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        torch.set_num_threads(cores)
        tcpu = cpu_oracle_train_step(wl, net_sd, 3)
        out["cpu_baseline"] = {"value": L / tcpu, "unit": "frames/s", "cores": cores, "kind": "port",
                               "sample": f"1 sequence x {L} LR frames, one iteration, torch CPU autograd through the oracle + torch Adam"}
    return out


def run_reference(args, wl, rank, world):
    if rank != 0:
        return
    cores = usable_cores()
    torch.set_num_threads(cores)
    sd = synth_weights(0)
    # each step = a bounded sample of the workload: as many sequences as take ~2 s on this host (at most the GPU arm's batch);
    # the whole run is held under ~3 minutes (K is honoured unless that bound would be exceeded)
    t1 = cpu_oracle_step(wl, 1, sd, 1)[0]
    B_sample = int(min(wl["B"], max(1, round(2.0 / max(t1, 1e-3)))))
    for _ in range(max(0, min(args.warmup, 2) - 1)):
        cpu_oracle_step(wl, B_sample, sd, 1)
    times, budget = [], 170.0
    for _ in range(max(1, args.steps)):
        times.append(cpu_oracle_step(wl, B_sample, sd, 1)[0])
        budget -= times[-1]
        if budget < times[-1]:
            break
    t = float(np.mean(times))
    val = B_sample * wl["L"] / t
    sample = f"{B_sample} sequence(s) x {wl['L']} LR frames per step (B={wl['B']} in the GPU arm), {len(times)} steps, fp32, torch CPU + C oracle"
    line = {"impl": "reference", "metric": "LR event-frames/sec", "value": val, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": min(args.warmup, 2), "ms_per_step": t * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": wl["desc"], "impl_note": "CPU restatement of the reference path (oracle/), "
                       "the reference's own Python cannot travel to the GPU box"},
            "cpu_baseline": {"value": val, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _JSON_OUT.write(json.dumps(line) + "\n")
    _JSON_OUT.flush()


# ------------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from the host instead of replaying a CUDA graph")
    ap.add_argument("--profile-out", default=None, help="write the per-launch timing table of one step here")
    ap.add_argument("--no-train", action="store_true", help="skip the training-iteration measurement (the `train` key)")
    args = ap.parse_args()
    # stdout carries exactly ONE line (the JSON): anything a library prints there (NCCL's version banner at N > 1) goes to stderr
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    global _JSON_OUT
    _JSON_OUT = os.fdopen(json_fd, "w")
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, wl, rank, world)
        return

    import torch.distributed as dist
    from esr_b200 import _lib
    from esr_b200.model import DeepRecurrNet
    from esr_b200.pipeline import EventSRPipeline

    assert torch.cuda.is_available(), "bench.py --impl ours needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # NCCL's version banner / debug lines must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=dev)

    scale, L, lr, B = wl["scale"], wl["L"], wl["lr"], wl["B"]
    hr = (lr[0] * scale, lr[1] * scale)
    sd = synth_weights(0)
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    pipe = EventSRPipeline(net, B, L, lr, scale, dev)
    pipe.sr_bias = synth_sr_bias(B, L, hr, 100 + rank).to(dev)

    xs, ys, ps, off = synth_events(B, L, lr, 100 + rank)
    h_xs, h_ys, h_ps, h_off = (t.pin_memory() for t in (xs, ys, ps, off))
    d_xs, d_ys, d_ps, d_off = (t.to(dev) for t in (xs, ys, ps, off))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """per-step CUDA-event timing with an (untimed) L2 flush between steps; returns total ms"""
        tot = 0.0
        for _ in range(steps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            tot += e0.elapsed_time(e1)
        return tot

    dev_step = lambda: pipe.run_device(d_xs, d_ys, d_ps, d_off, EVENTS_PER_FRAME)
    host_step = lambda: pipe.run_host(h_xs, h_ys, h_ps, h_off, EVENTS_PER_FRAME)

    dev_step()
    if not args.no_graph:
        pipe.capture()
    for _ in range(args.warmup):
        dev_step()
        host_step()
        pipe.collect(pipe.submit_host(h_xs, h_ys, h_ps, h_off, EVENTS_PER_FRAME))
    sampler = ClockSampler(local)
    sampler.start()
    barrier()
    l0 = _lib.lib().esr_launch_count()
    ms_dev = timed(dev_step, args.steps)
    launches = _lib.lib().esr_launch_count() - l0 + args.steps * pipe.graph_launches
    barrier()
    ms_e2e_sync = timed(host_step, args.steps)          # one batch at a time (latency view)
    barrier()
    # throughput view of the same end-to-end path: submit(i+1) is issued before finish(i), so the GPU runs batch i+1's network while
    # the host waits for / sizes batch i's event list, and the D2H of batch i-1 drains on a side stream (three in flight).
    # One timed region around all K steps (inputs + workspace exceed L2; no flush inside, it would serialise the overlap).
    def pipelined(steps):
        pend_a, pend_b = None, None                # submitted (network queued) / finished (D2H in flight)
        for _ in range(steps):
            h = pipe.submit_host(h_xs, h_ys, h_ps, h_off, EVENTS_PER_FRAME)
            if pend_a is not None:
                hb = pipe.finish(pend_a)            # host sizes batch i-1's output while the GPU runs batch i's network
                if pend_b is not None:
                    pipe.collect(pend_b)
                pend_b = hb
            pend_a = h
        for hdl in (pend_b, pend_a):
            if hdl is not None:
                pipe.collect(hdl)

    # warm-up of THIS loop form: with three batches in flight the caching allocator needs one more set of event / workspace
    # blocks than the one-at-a-time form; their first cudaMalloc synchronises the device and would land in the timed region
    pipelined(max(args.warmup, 3) + 1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    flush.zero_()
    torch.cuda.synchronize()
    e0.record()
    pipelined(args.steps)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1)
    barrier()
    sampler.stop_flag = True
    sampler.join(timeout=2)

    t = torch.tensor([ms_dev, ms_e2e, ms_e2e_sync], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_dev, ms_e2e, ms_e2e_sync = t.tolist()
    frames_per_step = world * B * L
    value = frames_per_step * args.steps / (ms_dev / 1e3)
    e2e = frames_per_step * args.steps / (ms_e2e / 1e3)

    # ---- roofline of the dominant kernel (k_conv_tc): per-launch CUDA events over one sequence batch (all windows) of the same workload
    roofline, prof_rows = None, []
    if rank == 0:
        import ctypes
        plan = net._plans[(B, L, hr[0], hr[1])]
        bank = torch.poisson(torch.full((B * L, 2, hr[0], hr[1]), 0.1)).to(dev)
        out = torch.empty(((L - 2) * B, 2, hr[0], hr[1]), device=dev)
        cap = 256
        cls = (ctypes.c_int * cap)()
        ms = (ctypes.c_float * cap)()
        fl = (ctypes.c_double * cap)()
        cnt = ctypes.c_int(0)
        acc = {0: [0.0, 0.0, 0], 1: [0.0, 0.0, 0], 2: [0.0, 0.0, 0], 3: [0.0, 0.0, 0]}
        per_launch = {}                                              # launch index -> [ms sum, flops]
        reps = 5
        for r in range(reps + 1):
            _lib.check(_lib.lib().esr_net_forward_profiled(plan.handle, _lib.ptr(bank), None, _lib.ptr(out),
                                                           cap, ctypes.byref(cnt), cls, ms, fl, _lib.stream_ptr()), "profiled forward")
            if r == 0:
                continue                                             # warm-up
            for i in range(cnt.value):
                a = acc[cls[i]]
                a[0] += ms[i]; a[1] += fl[i]; a[2] += 1
                pl = per_launch.setdefault(i, [0.0, fl[i], int(cls[i])])
                pl[0] += ms[i]
            if r == reps:
                prof_rows = [(i, int(cls[i]), float(ms[i]), float(fl[i])) for i in range(cnt.value)]
        tc_ms, tc_fl, tc_n = (acc[0][i] + acc[3][i] for i in range(3))     # tensor-core work: convs + GRU chain kernel
        peak_tf, peak_hbm, peak_src = measured_peaks()
        achieved = tc_fl / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
        tot_ms = sum(a[0] for a in acc.values())
        # the dominant single launch: the k_conv_tc instance with the most work (local_fusion.0.conv1/conv2, 192->192 3x3
        # over all window slots), excluding the cooperative GRU kernel, timed live above
        top = max((v for v in per_launch.values() if v[2] == 0), key=lambda v: v[1], default=None)
        top_ms = top[0] / reps if top else 0.0
        top_tf = top[1] / (top_ms * 1e-3) / 1e12 if top_ms > 0 else 0.0
        roofline = {"kernel": "k_conv_tc_persist (tcgen05 implicit-GEMM conv, persistent, two TMEM accumulators): largest launch = "
                              "local_fusion 192->192 3x3 over all window slots",
                    "bound": "tensor", "achieved": top_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": top_tf / peak_tf,
                    "peak_source": peak_src + ", bf16 sustained",
                    "launch_ms": top_ms, "algorithmic_gflop_per_launch": top[1] / 1e9 if top else None,
                    # dram__bytes_read.sum + dram__bytes_write.sum of this launch at cfg2, ncu --set full,
                    # profiles/r1_ncu_tcpersist.md (114.86 MB + 63.87 MB); algorithmic bytes: 144 x 32 x 32 x 192 x 4 B in + out = 226.5 MB
                    "traffic": 178.73e6 if args.workload == "cfg2" else None,
                    "tensor_pipe_active_pct_ncu": 75.0 if args.workload == "cfg2" else None,
                    "all_tc_launches": {"achieved": achieved, "frac": achieved / peak_tf},
                    "gru_chain_ms_per_step": acc[3][0] / reps,
                    "launches_per_step": tc_n // reps, "avg_launch_us": tc_ms / max(tc_n, 1) * 1e3,
                    "algorithmic_gflop_per_step": tc_fl / reps / 1e9,
                    "share_of_step_kernel_time": tc_ms / tot_ms if tot_ms else None,
                    "note": "algorithmic FLOPs (1x); the fp32-parity 3-pass split-bf16 product issues 3x that on the tensor pipe",
                    "cuda_core_conv_ms_per_step": acc[1][0] / reps, "elementwise_ms_per_step": acc[2][0] / reps,
                    "tc_ms_per_step": tc_ms / reps}
        if args.profile_out:
            with open(args.profile_out, "w") as f:
                f.write("idx,class(0=tc,1=direct,2=other,3=gru_chain),ms,algorithmic_flops\n")
                for row in prof_rows:
                    f.write("%d,%d,%.5f,%.0f\n" % row)

    # ---- stage breakdown of one step (device-resident inputs), CUDA events on the launching stream
    stages = None
    if rank == 0:
        from esr_b200 import encodings as enc
        from esr_b200.expand import expand
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        acc3 = [0.0, 0.0, 0.0]
        for _ in range(5):
            torch.cuda.synchronize()
            ev[0].record()
            with torch.no_grad():
                enc.encode_frames(d_xs, d_ys, d_ps, d_off, lr_size=lr, hr_size=hr, n_max_frame=EVENTS_PER_FRAME, out=pipe.bank)
            ev[1].record()
            with torch.no_grad():
                if pipe._graph is not None:
                    pipe._graph.replay()
                    sr_ = pipe._graph_sr
                else:
                    sr_ = pipe._windows()
            ev[2].record()
            expand(sr_, 0, 0)
            ev[3].record()
            torch.cuda.synchronize()
            for i in range(3):
                acc3[i] += ev[i].elapsed_time(ev[i + 1]) / 5
        stages = {"encode_ms": acc3[0], "network_ms": acc3[1], "redistribute_ms": acc3[2]}

    # ---- CPU baseline (oracle port) on rank 0, N=1 only
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = usable_cores()
        torch.set_num_threads(cores)
        t1 = cpu_oracle_step(wl, 1, sd, 1)[0]                        # warm-up + sizing: one sequence
        b_s = int(min(B, max(1, round(4.0 / max(t1, 1e-3)))))        # sequences per pass: ~4 s of CPU work, at most the GPU arm's batch
        ts, spent = [], 0.0
        while spent < 12.0 or len(ts) < 2:                           # ~12-16 s in total
            ts.append(cpu_oracle_step(wl, b_s, sd, 1)[0])
            spent += ts[-1]
        tcpu = float(np.mean(ts))
        cpu_baseline = {"value": b_s * L / tcpu, "unit": "frames/s", "cores": cores, "kind": "port",
                        "sample": f"{b_s} sequence(s) x {L} LR frames per pass (the GPU arm's step is B={B}), mean of {len(ts)} passes = {spent:.1f} s of CPU work, fp32"}

    # ---- training iteration: reported next to the inference headline, never mixed into `value`
    train_res = None
    if not args.no_train:
        train_res = measure_training(args, wl, sd, dev, rank, world, dist)

    if rank == 0:
        n_ev = EVENTS_PER_FRAME * B * L
        ev_out = pipe.run_device(d_xs, d_ys, d_ps, d_off, EVENTS_PER_FRAME)[1]
        line = {"metric": "LR event-frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "bf16x3 (split-bf16 tensor-core operands, fp32 accumulate) / fp32 elsewhere",
                "data": "synthetic",
                "config": {"workload": wl["desc"], "scale": scale, "seq_len": L, "lr": list(lr), "batch_per_gpu": B,
                           "events_per_frame": EVENTS_PER_FRAME, "windows_per_sequence": L - 2,
                           "redistribute_input": "model output + Poisson(0.3) synthetic counts",
                           "l2": "256 MiB buffer rewritten between timed steps (outside the timed intervals)",
                           "parallelism": f"dp{world} (batch shards, no data-path collective)",
                           "cuda_graph": not args.no_graph,
                           "plan": "sequence plan: per-frame and state-independent layers batched over all windows, ConvGRU chain serial"},
                "sr_frames_per_s": value * (L - 2) / L,
                "clocks": sampler.summary(),
                "e2e": {"value": e2e, "unit": "frames/s", "ms_per_step": ms_e2e / args.steps,
                        "mode": "software-pipelined, three batches in flight: network of batch i+1 | host sizing + emit/sort of batch i | D2H of batch i-1",
                        "one_at_a_time": {"value": frames_per_step * args.steps / (ms_e2e_sync / 1e3),
                                          "ms_per_step": ms_e2e_sync / args.steps},
                        "h2d_bytes_per_step": int(n_ev * 12 + (B * L + 1) * 8), "d2h_bytes_per_step": int(ev_out.numel() * 4)},
                "gpu_launches": int(launches),
                "tensor_roofline_whole_path": {"algorithmic_tflops": FLOP_PER_HR_PIXEL * hr[0] * hr[1] * B * (L - 2) * world /
                                               (ms_dev / args.steps / 1e3) / 1e12},
                "stages_ms_per_step": stages,
                "train": train_res,
                "roofline": roofline, "cpu_baseline": cpu_baseline}
        _JSON_OUT.write(json.dumps(line) + "\n")
        _JSON_OUT.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
