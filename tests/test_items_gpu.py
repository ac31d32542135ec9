"""the per-frame tensor factory of H5Dataset.__getitem__ on the GPU (esr_b200.dataset.create_item) against
fixtures produced by the REFERENCE's own, unmodified __getitem__ (tests/golden/make_golden_items.py), and the two plane
resizes against torch's CPU F.interpolate (the implementation the reference calls)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "items_golden.npz")
EXACT = ("inp_stack", "inp_cnt", "inp_near_cnt", "inp_near_stack", "inp_scaled_cnt", "inp_scaled_stack", "inp_down_cnt",
         "inp_down_scaled_cnt", "inp_custom_cnt", "inp_custom_scaled_cnt", "inp_custom_down_cnt", "inp_custom_down_scaled_cnt",
         "gt_custom_cnt", "gt_stack", "gt_cnt", "gt_img", "gt_inp_size_img", "frame")
BICUBIC = ("inp_bicubic_cnt", "inp_bicubic_stack")


def test_items_golden_is_complete():
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        for k in EXACT + BICUBIC:
            assert f"c{i}_{k}" in g.files, (i, k)
        H, W, scale, tb = (int(v) for v in g[f"c{i}_cfg"])
        assert g[f"c{i}_inp_cnt"].shape == (2, H, W) and g[f"c{i}_gt_cnt"].shape == (2, H * scale, W * scale)
        assert g[f"c{i}_inp_stack"].shape == (tb, H, W)
        assert g[f"c{i}_inp_down_cnt"].shape == (2, round(H / scale), round(W / scale))


def test_oracle_item_factory_matches_reference_getitem():
    """The CPU restatement (oracle/items.py) reproduces the reference's own __getitem__ output: integer-valued encodings and
    the resizes (same torch CPU kernels) bit for bit."""
    from oracle import items as oi
    g = np.load(GOLD)
    for i in range(int(g["n_cases"])):
        H, W, scale, tb = (int(v) for v in g[f"c{i}_cfg"])
        item = oi.create_item(g[f"c{i}_inp_events"], g[f"c{i}_gt_events"], (H, W), scale, tb)
        assert set(item) == set(EXACT + BICUBIC)
        for k in EXACT + BICUBIC:
            assert item[k].dtype == np.float32 and np.array_equal(item[k], g[f"c{i}_{k}"]), (i, k)


@pytest.mark.gpu
def test_create_item_matches_reference_getitem():
    """Integer-valued encodings bit-exact; the bicubic up-samplings to 1e-6 of the plane maximum (fp32 summation of 16 taps,
    possible FMA contraction in ATen's CPU build)."""
    from esr_b200 import dataset
    g = np.load(GOLD)
    dev = torch.device("cuda:0")
    for i in range(int(g["n_cases"])):
        H, W, scale, tb = (int(v) for v in g[f"c{i}_cfg"])
        item = dataset.create_item(g[f"c{i}_inp_events"], g[f"c{i}_gt_events"], (H, W), scale, time_bins=tb, device=dev)
        assert set(item) == set(EXACT + BICUBIC)
        for k in EXACT:
            got, want = item[k].cpu().numpy(), g[f"c{i}_{k}"]
            assert got.shape == want.shape and got.dtype == want.dtype, (i, k)
            assert np.array_equal(got, want), (i, k)
        for k in BICUBIC:
            got, want = item[k].cpu().numpy(), g[f"c{i}_{k}"]
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= 1e-6 * max(1.0, np.abs(want).max()), (i, k)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,size", [((2, 32, 48), (64, 96)), ((3, 45, 80), (180, 320)), ((1, 17, 23), (40, 31)),
                                        ((2, 64, 64), (64, 64)), ((1, 50, 70), (20, 33)), ((1, 1, 1), (4, 4))])
def test_interpolate_planes_vs_torch_cpu(shape, size):
    """bicubic / nearest, up- and down-scaling, non-integer ratios, identity, 1x1 input (all taps clamped)."""
    from esr_b200 import encodings as enc
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.poisson(torch.full(shape, 1.5), generator=g) - 1.0
    for mode, kw in (("bicubic", {"align_corners": False}), ("nearest", {})):
        want = F.interpolate(x.unsqueeze(0), size=size, mode=mode, **kw).squeeze(0)
        got = enc.interpolate_planes(x.cuda(), size, mode).cpu()
        assert got.shape == want.shape
        if mode == "nearest":
            assert torch.equal(got, want)
        else:
            assert (got - want).abs().max() <= 1e-6 * max(1.0, want.abs().max().item())
    got_cpu_in = enc.interpolate_planes(x, size, "nearest")
    assert not got_cpu_in.is_cuda
    with pytest.raises(ValueError):
        enc.interpolate_planes(x, size, "bilinear")


@pytest.mark.gpu
def test_collate_sequence_follows_getitem_order_for_out_of_range_events():
    """H5Dataset.__getitem__ builds the stack encoding first, which zeroes out-of-range events in place (frames of
    more than 3 events), so they add nothing to inp_cnt / inp_scaled_cnt / gt_cnt -- unlike the standalone events_to_channels,
    where an out-of-range negative event lands on neg[0, 0].  The batched path must equal the per-frame factory (create_item, pinned
    to the reference's own __getitem__ output) on malformed events, and keep the standalone quirk for frames of <= 3 events."""
    from esr_b200 import dataset as ds
    from esr_b200 import encodings as enc
    rng = np.random.default_rng(7)
    H, W, k = 24, 32, 2

    def events(n, Hs, Ws, bad):
        xs = rng.integers(0, Ws, n).astype(np.float64)
        ys = rng.integers(0, Hs, n).astype(np.float64)
        ps = rng.choice([-1.0, 1.0], n)
        idx = rng.choice(n, bad, replace=False)
        xs[idx[: bad // 2]] = Ws + 3                          # out of range to the right
        ys[idx[bad // 2:]] = -2                               # out of range above
        ps[idx] = -1.0                                        # negative: the case the quirk moves to neg[0, 0]
        return np.stack([xs, ys, np.sort(rng.random(n)) + 5.0, ps])
    L = 3
    inp = [[events(400, H, W, 12) for _ in range(L)]]
    gt = [[events(1600, H * k, W * k, 20) for _ in range(L)]]
    wins = ds.collate_sequence(inp, gt, (H, W), (H * k, W * k))
    bank = wins[0]["bank"]
    for l in range(L):
        item = ds.create_item(inp[0][l].copy(), gt[0][l].copy(), [H, W], k, 1)
        for key in ("inp_cnt", "inp_scaled_cnt", "gt_cnt"):
            assert torch.equal(bank[key][0, l], item[key]), (key, l)
        assert float(bank["inp_cnt"][0, l, 1, 0, 0]) == float(item["inp_cnt"][1, 0, 0])
    # frames of <= 3 events: the stack encoding returns early without touching the events -> the standalone quirk applies
    dev = torch.device("cuda:0")
    xs = torch.tensor([1.0, W + 5.0, 2.0], device=dev); ys = torch.tensor([1.0, 3.0, 2.0], device=dev); ps = torch.tensor([1.0, -1.0, -1.0], device=dev)
    off = torch.tensor([0, 3], dtype=torch.int64, device=dev)
    a = enc.encode_frames(xs.clone(), ys.clone(), ps, off, None, (H, W), 3, sanitised=True)
    b = enc.encode_frames(xs.clone(), ys.clone(), ps, off, None, (H, W), 3, sanitised=False)
    assert torch.equal(a, b) and float(a[0, 1, 0, 0]) == 1.0
