"""Generate tests/golden/events_golden.npz from the REFERENCE's own code (run in the build container only).

Needs /root/reference and oracle/_ref (python oracle/build_ref.py).  The reference's pure-Python
dataloader/encodings.py is imported from /root/reference (namespace package `dataloader`), its Cython
helpers from the oracle/_ref build of the unmodified .pyx files.  Output: small seeded input/output
vectors for
  - events_to_channels        (dataloader/encodings.py:289-304), incl. the out-of-range / fractional quirks
  - the LR->HR lift + scatter  (dataloader/h5dataset.py:508-528 semantics: x / W_lr * W_hr, fp32)
  - cnt2event mode 0 / 1       (dataloader/cython_cnt2event/cnt2event.pyx:18-116)
  - event_redistribute_*       (dataloader/cython_event_redistribute/event_redistribute.pyx:17-153)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(1, "/root/reference")

from dataloader import encodings as ref_enc  # noqa: E402
from dataloader.cython_cnt2event import cnt2event as ref_c2e  # noqa: E402
from dataloader.cython_event_redistribute import event_redistribute as ref_er  # noqa: E402


def main():
    out = {}
    rng = np.random.default_rng(20220907)

    # --- events_to_channels: the SURVEY 8c known answer + seeded cases with out-of-range / fractional coords
    cases = []
    xs = np.array([0, 5, 5, 2.9, 6, -1, 3], np.float32)
    ys = np.array([0, 3, 3, 1.2, 1, 2, 4], np.float32)
    ps = np.array([1, -1, -1, 1, 1, 1, -1], np.float32)
    cases.append((xs, ys, ps, (4, 6)))
    for (n, H, W) in [(0, 5, 7), (1, 3, 3), (1000, 16, 24), (5000, 45, 80), (20000, 128, 128)]:
        xs = (rng.random(n) * (W + 6) - 3).astype(np.float32)
        ys = (rng.random(n) * (H + 6) - 3).astype(np.float32)
        ps = rng.choice(np.array([-1, 1], np.float32), n)
        if n >= 1000:
            xs[::7] = np.floor(xs[::7])
            ps[::11] = 0.0
            ps[5::13] *= 2.0          # non-unit polarities add ps**2
        cases.append((xs, ys, ps, (H, W)))
    out["n_e2c"] = len(cases)
    for i, (xs, ys, ps, (H, W)) in enumerate(cases):
        tx, ty, tp = torch.from_numpy(xs.copy()), torch.from_numpy(ys.copy()), torch.from_numpy(ps.copy())
        if len(xs) == 0:
            img = torch.zeros(2, H, W)   # index_put_ with empty index is fine, but keep it explicit
            img = ref_enc.events_to_channels(tx, ty, tp, sensor_size=(H, W))
        else:
            img = ref_enc.events_to_channels(tx, ty, tp, sensor_size=(H, W))
        out[f"e2c{i}_xs"], out[f"e2c{i}_ys"], out[f"e2c{i}_ps"] = xs, ys, ps
        out[f"e2c{i}_hw"] = np.array([H, W])
        out[f"e2c{i}_out"] = img.numpy()
        out[f"e2c{i}_xs_after"], out[f"e2c{i}_ys_after"] = tx.numpy(), ty.numpy()

    # --- LR->HR lift (create_normalized_events + create_scaled_encoding 'cnt')
    lifts = [(64, 64, 2), (128, 128, 4), (346, 260, 2), (80, 45, 4)]
    out["n_lift"] = len(lifts)
    for i, (W, H, k) in enumerate(lifts):
        n = 4096
        xs = rng.integers(0, W, n).astype(np.float32)
        ys = rng.integers(0, H, n).astype(np.float32)
        ps = rng.choice(np.array([-1, 1], np.float32), n)
        tx, ty, tp = torch.from_numpy(xs.copy()), torch.from_numpy(ys.copy()), torch.from_numpy(ps.copy())
        xn, yn = tx / W, ty / H                                       # h5dataset.py:515
        img = ref_enc.events_to_channels(xn * (W * k), yn * (H * k), tp, sensor_size=(H * k, W * k))  # :526
        out[f"lift{i}_xs"], out[f"lift{i}_ys"], out[f"lift{i}_ps"] = xs, ys, ps
        out[f"lift{i}_dims"] = np.array([H, W, k])
        out[f"lift{i}_out"] = img.numpy()

    # --- events_to_stack_no_polarity (+ binary_search_torch_tensor) and events_to_voxel
    tcases = []
    for (n, H, W, TB) in [(2, 4, 4, 3), (50, 8, 10, 1), (3000, 24, 32, 5), (4000, 45, 80, 1), (2500, 16, 16, 4)]:
        xs = (rng.random(n) * (W + 4) - 2).astype(np.float32)
        ys = (rng.random(n) * (H + 4) - 2).astype(np.float32)
        ps = rng.choice(np.array([-1, 1], np.float32), n)
        ts = np.sort(rng.random(n)).astype(np.float32)
        if n >= 2500:
            ts = np.round(ts * 40) / 40            # many equal timestamps: exercises the search's early exits
            ts = np.sort(ts).astype(np.float32)
        if n > 3:
            ts = ((ts - ts[0]) / (ts[-1] - ts[0] + np.float32(1e-6))).astype(np.float32)   # event_formatting normalisation
        tcases.append((xs, ys, ts, ps, H, W, TB))
    out["n_stack"] = len(tcases)
    for i, (xs, ys, ts, ps, H, W, TB) in enumerate(tcases):
        tx, ty, tt, tp = (torch.from_numpy(a.copy()) for a in (xs, ys, ts, ps))
        st = ref_enc.events_to_stack_no_polarity(tx, ty, tt, tp, TB, sensor_size=(H, W))
        out[f"stk{i}_xs"], out[f"stk{i}_ys"], out[f"stk{i}_ts"], out[f"stk{i}_ps"] = xs, ys, ts, ps
        out[f"stk{i}_dims"] = np.array([H, W, TB])
        out[f"stk{i}_out"] = st.numpy()
        out[f"stk{i}_xs_after"], out[f"stk{i}_ps_after"] = tx.numpy(), tp.numpy()
        tx, ty, tt, tp = (torch.from_numpy(a.copy()) for a in (xs, ys, ts, ps))
        vx = ref_enc.events_to_voxel(tx, ty, tt, tp, max(TB, 2), sensor_size=(H, W))
        out[f"stk{i}_voxel"] = vx.numpy()
        out[f"stk{i}_voxel_xs_after"] = tx.numpy()
        tx, ty, tp = (torch.from_numpy(a.copy()) for a in (xs, ys, ps))
        pm = tp.clone()                                               # unit polarities (the reference's own use)
        out[f"stk{i}_mask_ps"] = pm.numpy().copy()
        out[f"stk{i}_mask"] = ref_enc.events_to_mask(tx, ty, pm, sensor_size=(H, W)).numpy()
        out[f"stk{i}_mask_ps_after"] = pm.numpy()

    # --- cnt2event
    c = np.zeros((1, 2, 2, 3), np.float32)
    c[0, 0, 0, 1], c[0, 0, 1, 2], c[0, 1, 0, 0], c[0, 1, 1, 1] = 2.5, 3.0, 1.0, 3.5
    cnts = [c, np.zeros((2, 2, 4, 4), np.float32)]
    for (B, H, W, lam) in [(1, 8, 8, 0.5), (3, 17, 23, 0.7), (2, 32, 48, 0.3), (2, 16, 16, 3.0)]:
        v = rng.poisson(lam, (B, 2, H, W)).astype(np.float32) + (rng.random((B, 2, H, W)).astype(np.float32) - 0.5) * 0.98
        v = np.maximum(v, 0).astype(np.float32)
        if B == 3:
            v[1] = 0.2       # a sample that rounds to empty
        cnts.append(v)
    h = np.full((1, 2, 3, 3), 0.5, np.float32)
    h[0, 0, 1, 1] = 1.5
    h[0, 1, 2, 2] = 2.5      # half-to-even: 0.5->0, 1.5->2, 2.5->2
    cnts.append(h)
    out["n_c2e"] = len(cnts)
    for i, v in enumerate(cnts):
        out[f"c2e{i}_in"] = v
        for mode in (0, 1):
            out[f"c2e{i}_out{mode}"] = ref_c2e.cnt2event(v, mode)

    # --- event_redistribute
    st = np.zeros((1, 3, 2, 2), np.float32)
    st[0, 0, 0, 0], st[0, 1, 1, 0], st[0, 2, 0, 1], st[0, 0, 1, 1] = 2, -3, 1, -1.5
    stacks = [st, np.zeros((2, 2, 2, 3, 3), np.float32)]
    for shape in [(1, 2, 3, 9, 11), (3, 2, 1, 8, 8), (2, 5, 12, 10), (3, 1, 6, 6)]:
        v = (rng.poisson(0.6, shape) * rng.choice([-1, 1], shape)).astype(np.float32)
        v += ((rng.random(shape) - 0.5) * 0.9).astype(np.float32)
        stacks.append(v.astype(np.float32))
    z = np.zeros((2, 2, 2, 2), np.float32)
    z[0, 0, 0, 0], z[0, 1, 1, 1] = 2, -2       # sample 0 sums to zero -> treated as empty by the reference
    z[1, 0, 0, 1] = 1
    stacks.append(z)
    out["n_er"] = len(stacks)
    for i, v in enumerate(stacks):
        out[f"er{i}_in"] = v
        fn = ref_er.event_redistribute_PolarityStack if v.ndim == 5 else ref_er.event_redistribute_NoPolarityStack
        for mode in (0, 1):
            out[f"er{i}_out{mode}"] = fn(v, mode)

    path = os.path.join(HERE, "events_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
