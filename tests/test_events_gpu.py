"""GPU parity of the event path: CUDA kernels (through the C ABI) vs the C oracle and the reference-generated
golden vectors.  Bit-exact (integer / index work; fp32 timestamps are produced by identical float64 arithmetic)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available()
    return torch.device("cuda:0")


def test_events_to_channels_golden(golden_events, dev):
    from esr_b200 import encodings as enc
    g = golden_events
    for i in range(int(g["n_e2c"])):
        xs, ys, ps = (torch.from_numpy(g[f"e2c{i}_{k}"].copy()) for k in ("xs", "ys", "ps"))
        H, W = g[f"e2c{i}_hw"]
        out = enc.events_to_channels(xs, ys, ps, sensor_size=(int(H), int(W)))      # CPU tensors in, like the reference
        assert not out.is_cuda
        assert np.array_equal(out.numpy(), g[f"e2c{i}_out"]), i
        assert np.array_equal(xs.numpy(), g[f"e2c{i}_xs_after"]) and np.array_equal(ys.numpy(), g[f"e2c{i}_ys_after"])
        # CUDA tensors in place
        cx, cy, cp = (torch.from_numpy(g[f"e2c{i}_{k}"].copy()).to(dev) for k in ("xs", "ys", "ps"))
        out = enc.events_to_channels(cx, cy, cp, sensor_size=(int(H), int(W)))
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), g[f"e2c{i}_out"])
        assert np.array_equal(cx.cpu().numpy(), g[f"e2c{i}_xs_after"])


def test_lift_golden(golden_events, dev):
    from esr_b200 import encodings as enc
    g = golden_events
    for i in range(int(g["n_lift"])):
        H, W, k = (int(v) for v in g[f"lift{i}_dims"])
        xs, ys, ps = (torch.from_numpy(g[f"lift{i}_{n}"]).to(dev) for n in ("xs", "ys", "ps"))
        off = torch.tensor([0, xs.numel()], dtype=torch.int64, device=dev)
        out = enc.encode_frames(xs, ys, ps, off, lr_size=(H, W), hr_size=(H * k, W * k))
        assert np.array_equal(out[0].cpu().numpy(), g[f"lift{i}_out"]), i


def test_events_to_image_vs_oracle(dev):
    from esr_b200 import encodings as enc
    rng = np.random.default_rng(5)
    n, H, W = 30000, 37, 61
    xs = (rng.random(n) * (W + 8) - 4).astype(np.float32)
    ys = (rng.random(n) * (H + 8) - 4).astype(np.float32)
    ps = rng.integers(-3, 4, n).astype(np.float32)
    tx, ty, tp = torch.from_numpy(xs.copy()), torch.from_numpy(ys.copy()), torch.from_numpy(ps.copy())
    img = enc.events_to_image(tx, ty, tp, sensor_size=(H, W))
    oor = (xs >= W) | (xs < 0) | (ys >= H) | (ys < 0)
    want = np.zeros((H, W), np.float32)
    np.add.at(want, (ys[~oor].astype(np.int64), xs[~oor].astype(np.int64)), ps[~oor])
    assert np.array_equal(img.numpy(), want)
    assert (tx.numpy()[oor] == 0).all() and (tp.numpy()[oor] == 0).all() and np.array_equal(tx.numpy()[~oor], xs[~oor])


def test_encode_frames_batched_vs_oracle(dev):
    from esr_b200 import encodings as enc
    from oracle import events as oe
    rng = np.random.default_rng(11)
    F, Hl, Wl, k = 9, 45, 80, 2
    lens = rng.integers(0, 5000, F)
    lens[3] = 0
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
    n = int(off[-1])
    xs = rng.integers(-2, Wl + 2, n).astype(np.float32)
    ys = rng.integers(-2, Hl + 2, n).astype(np.float32)
    ps = rng.choice(np.array([-1, 1], np.float32), n)
    out = enc.encode_frames(torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), torch.from_numpy(ps).to(dev),
                            torch.from_numpy(off).to(dev), lr_size=(Hl, Wl), hr_size=(Hl * k, Wl * k),
                            n_max_frame=int(lens.max())).cpu().numpy()
    for f in range(F):
        a, b = off[f], off[f + 1]
        want = oe.events_to_channels(oe.lift_coords(xs[a:b], Wl, Wl * k), oe.lift_coords(ys[a:b], Hl, Hl * k),
                                     ps[a:b], (Hl * k, Wl * k))
        assert np.array_equal(out[f], want), f


@pytest.mark.parametrize("mode", [0, 1])
def test_cnt2event_golden(golden_events, mode, dev):
    from esr_b200 import cnt2event as c2e
    g = golden_events
    for i in range(int(g["n_c2e"])):
        got = c2e.cnt2event(g[f"c2e{i}_in"], mode)
        assert got.dtype == np.float32 and np.array_equal(got, g[f"c2e{i}_out{mode}"]), i


@pytest.mark.parametrize("mode", [0, 1])
def test_event_redistribute_golden(golden_events, mode, dev):
    from esr_b200 import event_redistribute as er
    g = golden_events
    for i in range(int(g["n_er"])):
        v = g[f"er{i}_in"]
        fn = er.event_redistribute_PolarityStack if v.ndim == 5 else er.event_redistribute_NoPolarityStack
        assert np.array_equal(fn(v, mode), g[f"er{i}_out{mode}"]), i


@pytest.mark.parametrize("shape,lam", [((1, 2, 64, 64), 0.3), ((8, 2, 256, 256), 0.3), ((3, 2, 100, 37), 2.5),
                                       ((2, 2, 8, 8), 40.0)])
def test_cnt2event_vs_oracle(shape, lam, dev):
    from esr_b200 import cnt2event as c2e
    from oracle import events as oe
    rng = np.random.default_rng(hash(shape) % 1000)
    cnt = rng.poisson(lam, shape).astype(np.float32) + ((rng.random(shape) - 0.5) * 0.9).astype(np.float32)
    cnt = np.maximum(cnt, 0).astype(np.float32)
    for mode in (0, 1):
        got = c2e.cnt2event_cuda(torch.from_numpy(cnt).to(dev), mode).cpu().numpy()
        assert np.array_equal(got, oe.cnt2event(cnt, mode)), (shape, mode)


def test_redistribute_vs_oracle_and_roundtrip(dev):
    from esr_b200 import encodings as enc
    from oracle import events as oe
    rng = np.random.default_rng(3)
    st = (rng.poisson(0.5, (4, 2, 5, 40, 56)) * rng.choice([-1, 1], (4, 2, 5, 40, 56))).astype(np.float32)
    for mode, name in ((0, "linear"), (1, "random")):
        got = enc.cython_event_redistribute(torch.from_numpy(st), name).numpy()
        assert np.array_equal(got, oe.event_redistribute(st, mode))
        got1 = enc.multiprocess_cython(torch.from_numpy(st), name).numpy()
        per = [oe.event_redistribute(st[i:i + 1], mode)[0] for i in range(4)]
        for i in range(4):
            assert np.array_equal(got1[i, :len(per[i])], per[i])
    # cnt -> events -> cnt round trip at a BASELINE-sized grid (size-independent property)
    cnt = rng.poisson(0.3, (2, 2, 512, 512)).astype(np.float32)
    ev = __import__("esr_b200.cnt2event", fromlist=["x"]).cnt2event_cuda(torch.from_numpy(cnt).to(dev), 0)
    for b in range(2):
        e = ev[b]
        e = e[e[:, 3] != 0]
        back = enc.events_to_channels(e[:, 0].contiguous(), e[:, 1].contiguous(), e[:, 3].contiguous(), (512, 512))
        assert torch.equal(back.cpu(), torch.from_numpy(cnt[b]))
        t = ev[b, :, 2]
        n = int((ev[b, :, 3] != 0).sum())
        assert bool((t[1:n] >= t[:n - 1]).all())      # sortedness


def test_empty_and_error_cases(dev):
    from esr_b200 import cnt2event as c2e
    z = np.zeros((3, 2, 5, 5), np.float32)
    assert c2e.cnt2event(z, 0).shape == (3, 1, 4)
    c = np.zeros((1, 2, 2, 2), np.float32)
    c[0, 0, 0, 0], c[0, 1, 1, 1] = 3, -1
    with pytest.raises(ValueError):
        c2e.cnt2event(c, 0)
    with pytest.raises(ValueError):
        c2e.cnt2event(z.astype(np.float64), 0)


def test_stack_and_voxel_golden(golden_events, dev):
    """Row 3 of the scope table: events_to_stack_no_polarity with the reference's binary-search bin edges, and
    events_to_voxel.  The stack is bit-exact (per-bin sums of +-1); voxel weights are fp32 sums accumulated with
    atomics, so they are compared to 1e-5 relative (summation order differs)."""
    from esr_b200 import encodings as enc
    g = golden_events
    for i in range(int(g["n_stack"])):
        H, W, TB = (int(v) for v in g[f"stk{i}_dims"])
        for on_gpu in (False, True):
            xs, ys, ts, ps = (torch.from_numpy(g[f"stk{i}_{k}"].copy()) for k in ("xs", "ys", "ts", "ps"))
            if on_gpu:
                xs, ys, ts, ps = (t.to(dev) for t in (xs, ys, ts, ps))
            out = enc.events_to_stack_no_polarity(xs, ys, ts, ps, TB, sensor_size=(H, W))
            assert out.is_cuda == on_gpu
            assert np.array_equal(out.cpu().numpy(), g[f"stk{i}_out"]), (i, on_gpu)
            assert np.array_equal(xs.cpu().numpy(), g[f"stk{i}_xs_after"])
            assert np.array_equal(ps.cpu().numpy(), g[f"stk{i}_ps_after"])
        xs, ys, ts, ps = (torch.from_numpy(g[f"stk{i}_{k}"].copy()) for k in ("xs", "ys", "ts", "ps"))
        vox = enc.events_to_voxel(xs, ys, ts, ps, max(TB, 2), sensor_size=(H, W)).numpy()
        want = g[f"stk{i}_voxel"]
        assert np.abs(vox - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), i
        assert np.array_equal(xs.numpy(), g[f"stk{i}_voxel_xs_after"])


@pytest.mark.parametrize("peak", [1, 2, 3, 16, 17, 128, 255, 256, 300])
def test_cnt2event_compact_and_raw_sort_keys(dev, peak):
    """Counts up to 255 take the compact-rank sort keys, larger ones the raw fp32 timestamp bits; both must equal the
    oracle bit for bit (ties between pixels with different counts are where a wrong rank table would show)."""
    from esr_b200 import cnt2event as c2e
    from oracle import events as oe
    rng = np.random.default_rng(peak)
    cnt = rng.integers(0, min(peak, 6) + 1, (2, 2, 12, 10)).astype(np.float32)
    cnt[0, 0, 3, 4] = peak
    cnt[1, 1, 0, 0] = max(1, peak - 1)
    got = c2e.cnt2event_cuda(torch.from_numpy(cnt).to(dev), 0).cpu().numpy()
    assert np.array_equal(got, oe.cnt2event(cnt, 0))


@pytest.mark.parametrize("peak", [3, 300])
def test_cnt2event_ragged_segments(dev, peak):
    """The sort is segmented by sample with tiles aligned to the sample starts: empty samples (leading, in the middle and
    trailing), a 3-event sample, samples of exactly one tile (2048 events) and one tile + 1, and a many-tile sample, for
    the 1-pass rank keys (peak 3), the 4-pass raw keys (peak 300) and the random mode."""
    from esr_b200 import cnt2event as c2e
    from oracle import events as oe
    rng = np.random.default_rng(peak)
    H, W = 48, 64
    cnt = np.zeros((8, 2, H, W), np.float32)
    cnt[1, 0, 5, 7] = 3                                                   # 3 events
    flat = cnt[2].reshape(-1)
    flat[rng.choice(flat.size, 2048, replace=False)] = 1                  # exactly one tile
    flat = cnt[4].reshape(-1)
    flat[rng.choice(flat.size, 2047, replace=False)] = 1
    flat[np.flatnonzero(flat)[0]] = 2                                     # one tile + 1
    cnt[5] = rng.integers(0, 4, (2, H, W))                                # ~9k events, several tiles
    cnt[5, 1, 0, 0] = peak
    cnt[6, 1, H - 1, W - 1] = 1                                           # single event; samples 0, 3, 7 stay empty
    for mode in (0, 1):
        got = c2e.cnt2event_cuda(torch.from_numpy(cnt).to(dev), mode).cpu().numpy()
        assert np.array_equal(got, oe.cnt2event(cnt, mode)), mode


@pytest.mark.parametrize("H,W,n", [(32, 32, 60000), (64, 48, 200000), (128, 128, 400000)])
def test_dense_frames_scatter_vs_oracle(dev, H, W, n):
    """Dense frames (many events per image cell, heavy atomic contention): unit polarities are integer sums and must be
    bit-exact; non-integer weights are fp32 atomics in arbitrary order (tolerance)."""
    from esr_b200 import encodings as enc
    from oracle import events as oe
    rng = np.random.default_rng(H + n)
    xs = rng.integers(-3, W + 3, n).astype(np.float32)
    ys = rng.integers(-3, H + 3, n).astype(np.float32)
    ps = rng.choice(np.array([-1, 1], np.float32), n)
    off = np.array([0, n // 3, n], np.int64)                     # two frames, both dense
    got = enc.encode_frames(torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), torch.from_numpy(ps).to(dev),
                            torch.from_numpy(off).to(dev), hr_size=(H, W), n_max_frame=int(n - n // 3)).cpu().numpy()
    for f in range(2):
        a, b = off[f], off[f + 1]
        want = oe.events_to_channels(xs[a:b].copy(), ys[a:b].copy(), ps[a:b], (H, W))
        assert np.array_equal(got[f], want), f
    ps2 = ps.copy()
    ps2[::5] *= 1.5                                              # ps*ps = 2.25: not an integer
    got = enc.encode_frames(torch.from_numpy(xs).to(dev), torch.from_numpy(ys).to(dev), torch.from_numpy(ps2).to(dev),
                            torch.from_numpy(off).to(dev), hr_size=(H, W), n_max_frame=int(n - n // 3)).cpu().numpy()
    want = oe.events_to_channels(xs[:off[1]].copy(), ys[:off[1]].copy(), ps2[:off[1]], (H, W))
    np.testing.assert_allclose(got[0], want, rtol=1e-5, atol=1e-3)


def test_events_to_mask_golden(golden_events, dev):
    """accumulate=False: the last event on a pixel wins.  torch's CPU index_put_ is sequential only for small inputs
    (larger ones are split over threads, so WHICH duplicate wins is not defined by the reference itself); with unit
    polarities every in-range writer stores 1, and only pixel (0,0) -- where zeroed out-of-range events also land -- is
    order dependent, so it is checked for membership instead of equality."""
    from esr_b200 import encodings as enc
    g = golden_events
    for i in range(int(g["n_stack"])):
        H, W, _ = (int(v) for v in g[f"stk{i}_dims"])
        xs, ys, ps = (torch.from_numpy(g[f"stk{i}_{k}"].copy()) for k in ("xs", "ys", "mask_ps"))
        out = enc.events_to_mask(xs, ys, ps, sensor_size=(H, W))
        got, want = out.numpy().copy(), g[f"stk{i}_mask"].copy()
        assert got[0, 0] in (0.0, 1.0)
        got[0, 0] = want[0, 0] = 0.0
        assert np.array_equal(got, want), i
        assert np.array_equal(ps.numpy(), g[f"stk{i}_mask_ps_after"])


def test_dataset_glue_matches_reference_semantics(dev, golden_events):
    """event_formatting / create_normalized_events / create_scaled_encoding('cnt') = the lift fixtures of the reference."""
    from esr_b200 import dataset as ds
    g = golden_events
    for i in range(int(g["n_lift"])):
        H, W, k = (int(v) for v in g[f"lift{i}_dims"])
        xs, ys, ps = g[f"lift{i}_xs"], g[f"lift{i}_ys"], g[f"lift{i}_ps"]
        ts = np.linspace(5.0, 6.0, len(xs))
        ev = ds.event_formatting(np.stack([xs, ys, ts, ps]), device=dev)
        assert ev.is_cuda and ev.dtype == torch.float32 and float(ev[2, 0]) == 0.0 and float(ev[2, -1]) < 1.0
        norm = ds.create_normalized_events(ev, (H, W))
        cnt = ds.create_scaled_encoding(norm, (H * k, W * k), 'cnt')
        assert np.array_equal(cnt.cpu().numpy(), g[f"lift{i}_out"]), i
    fr = torch.arange(2 * 5 * 3, device=dev).view(2, 5, 3)
    wins = ds.sliding_windows(fr)
    assert len(wins) == 3 and torch.equal(wins[1], fr[:, 1:4])


def test_collate_sequence_matches_per_frame_reference_semantics(dev):
    """The batched GPU collate = per frame create_stack_encoding (sanitises out-of-range events in place), create_cnt_encoding,
    create_normalized_events + create_scaled_encoding('cnt') in the order of H5Dataset.__getitem__ (oracle restatement of
    dataloader/h5dataset.py:337-354, 508-528, 611-619), windowed like custom_collate."""
    from esr_b200 import dataset as ds
    from oracle import events as oe
    rng = np.random.default_rng(4)
    B, L, H, W, k = 2, 5, 45, 80, 4

    def frame(n, h, w):
        ev = np.stack([rng.integers(-1, w + 1, n), rng.integers(-1, h + 1, n), np.sort(rng.random(n)), rng.choice([-1.0, 1.0], n)])
        return ev.astype(np.float64)

    inp = [[frame(int(rng.integers(0, 700)), H, W) for _ in range(L)] for _ in range(B)]
    gt = [[frame(int(rng.integers(1, 3000)), H * k, W * k) for _ in range(L)] for _ in range(B)]
    wins = ds.collate_sequence(inp, gt, (H, W), (H * k, W * k), device=dev)
    assert len(wins) == L - 2 and wins[0]['inp_scaled_cnt'].shape == (B, 3, 2, H * k, W * k)
    for b in range(B):
        for l in range(L):
            # the order of H5Dataset.__getitem__ (h5dataset.py:337-354): the stack encoding runs first and zeroes out-of-range
            # events in place (frames of more than 3 events), so the count encodings below see the sanitised arrays
            def formatted(ev):
                x, y, t, p = (ev[c].astype(np.float32) for c in range(4))
                if len(t):
                    t = (t - t[0]) / (t[-1] - t[0] + np.float32(1e-6))
                return x, y, t.astype(np.float32), p
            x, y, t, p = formatted(inp[b][l])
            oe.events_to_stack_no_polarity(x, y, t, p, 1, (H, W))
            want_cnt = oe.events_to_channels(x.copy(), y.copy(), p, (H, W))
            want_scaled = oe.events_to_channels(oe.lift_coords(x, W, W * k), oe.lift_coords(y, H, H * k), p, (H * k, W * k))
            gx, gy, gtt, gp = formatted(gt[b][l])
            oe.events_to_stack_no_polarity(gx, gy, gtt, gp, 1, (H * k, W * k))
            want_gt = oe.events_to_channels(gx.copy(), gy.copy(), gp, (H * k, W * k))
            w0 = min(l, L - 3)
            assert np.array_equal(wins[w0]['inp_cnt'][b, l - w0].cpu().numpy(), want_cnt), (b, l)
            assert np.array_equal(wins[w0]['inp_scaled_cnt'][b, l - w0].cpu().numpy(), want_scaled), (b, l)
            assert np.array_equal(wins[w0]['gt_cnt'][b, l - w0].cpu().numpy(), want_gt), (b, l)


def test_more_than_256_samples_draw_one_random_stream():
    """with B > 256 the batch is processed in 256-sample parts; in random-timestamp mode they must share ONE numpy stream
    (the reference seeds once per call, cnt2event.pyx:25), not restart it per part."""
    from esr_b200 import cnt2event as c2e
    from oracle import events as oe
    rng = np.random.default_rng(5)
    cnt = rng.poisson(0.4, (300, 2, 6, 7)).astype(np.float32)
    cnt[260] = 0                                                     # an empty sample in the second part
    for mode in (0, 1):
        got = c2e.cnt2event_cuda(torch.from_numpy(cnt).cuda(), mode).cpu().numpy()
        want = oe.cnt2event(cnt, mode)
        assert got.shape == want.shape and np.array_equal(got, want), mode


@pytest.mark.parametrize("shape,peak", [((2, 2, 5, 5), 3), ((3, 2, 33, 31), 7), ((2, 2, 64, 64), 64), ((2, 2, 64, 64), 65),
                                        ((5, 2, 96, 128), 20), ((1, 2, 256, 256), 33)])
def test_cnt2event_fused_path(dev, shape, peak, monkeypatch):
    """cnt2event / linear through esr_cnt2event_fused (three launches, rows written before the host knows maxlen) against the
    oracle and against the general chain: odd sizes (no 16-byte loads), a ragged last tile, counts at the 64 limit and one above
    (fallback), an inactive sample whose +1 / -1 cancel (cnt2event.pyx:56: rounded sum == 0 emits one zero row), a repeat call
    (capacity taken from the previous call) and a capacity that is too small (fallback)."""
    from esr_b200 import cnt2event as c2e, expand as ex
    from oracle import events as oe
    rng = np.random.default_rng(peak * 7 + shape[0])
    cnt = rng.poisson(0.4, shape).astype(np.float32) + ((rng.random(shape) - 0.5) * 0.9).astype(np.float32)
    cnt = np.maximum(cnt, 0).astype(np.float32)
    cnt[0, 1, 2, 3] = peak
    cnt[-1, 0, 0, 0] = max(1, peak - 1)
    if shape[0] >= 3:
        cnt[1] = 0
        cnt[1, 0, 1, 1], cnt[1, 1, 2, 2] = 1, -1                     # sums to zero: the reference treats the sample as empty
    want = oe.cnt2event(cnt, 0)
    x = torch.from_numpy(cnt).to(dev)
    ex._XF_ROWS.clear()
    got = c2e.cnt2event_cuda(x, 0)
    assert np.array_equal(got.cpu().numpy(), want)
    again = c2e.cnt2event_cuda(x, 0)                                  # capacity now comes from the first call
    assert np.array_equal(again.cpu().numpy(), want)
    key = (str(x.device), shape[0], shape[2], shape[3])
    assert ex._XF_ROWS[key][0] == want.shape[0] * want.shape[1]
    ex._XF_ROWS[key] = (max(1, want.shape[1] // 2), peak)                     # too small a guess: general chain, same answer
    monkeypatch.setattr(ex, "_XF_SLACK", 0)
    small = c2e.cnt2event_cuda(x, 0)
    assert np.array_equal(small.cpu().numpy(), want)
    monkeypatch.setenv("ESR_EXPAND_FUSED", "0")
    legacy = c2e.cnt2event_cuda(x, 0)
    assert np.array_equal(legacy.cpu().numpy(), want)


def test_cnt2event_fused_full_size_properties(dev):
    """BASELINE-sized grid (8 x 2 x 256 x 256, ~2 M events): fused path == general chain bit for bit, rows sorted by time inside
    each sample, counts recovered by scattering the rows back (size-independent properties; the oracle covers the small cases)."""
    from esr_b200 import cnt2event as c2e, encodings as enc
    rng = np.random.default_rng(11)
    cnt = rng.poisson(1.0, (8, 2, 256, 256)).astype(np.float32)
    cnt[3] = 0
    x = torch.from_numpy(cnt).to(dev)
    got = c2e.cnt2event_cuda(x, 0)
    os.environ["ESR_EXPAND_FUSED"] = "0"
    try:
        ref = c2e.cnt2event_cuda(x, 0)
    finally:
        del os.environ["ESR_EXPAND_FUSED"]
    assert torch.equal(got, ref)
    for b in range(8):
        e = got[b]
        n = int((e[:, 3] != 0).sum())
        assert n == int(cnt[b].sum())
        assert n < 2 or bool((e[1:n, 2] >= e[:n - 1, 2]).all())
        assert not bool(e[n:].any())
        back = enc.events_to_channels(e[:n, 0].contiguous(), e[:n, 1].contiguous(), e[:n, 3].contiguous(), (256, 256))
        assert torch.equal(back.cpu(), torch.from_numpy(cnt[b]))
