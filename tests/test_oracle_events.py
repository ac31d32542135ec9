"""The C oracle (oracle/esr_oracle.c) against the reference-generated golden vectors and the SURVEY 8c KATs.
CPU only.  This is what pins the oracle; the GPU parity tests then compare the CUDA path with the oracle."""
import numpy as np
import pytest

from oracle import events as oe


def test_kat_cnt2event_survey_8c():
    # known answer produced by the reference's own Cython
    c = np.zeros((1, 2, 2, 3), np.float32)
    c[0, 0, 0, 1], c[0, 0, 1, 2], c[0, 1, 0, 0], c[0, 1, 1, 1] = 2.5, 3.0, 1.0, 3.5
    got = oe.cnt2event(c, 0)[0]
    third, two3 = np.float32(1 / 3), np.float32(2 / 3)
    want = np.array([(1, 0, 0, 1), (2, 1, 0, 1), (0, 0, 0, -1), (1, 1, 0, -1), (1, 1, third, -1), (2, 1, .5, 1),
                     (1, 1, two3, -1), (1, 0, 1, 1), (2, 1, 1, 1), (1, 1, 1, -1)], np.float32)
    assert np.array_equal(got, want)


def test_kat_redistribute_survey_8c():
    st = np.zeros((1, 3, 2, 2), np.float32)
    st[0, 0, 0, 0], st[0, 1, 1, 0], st[0, 2, 0, 1], st[0, 0, 1, 1] = 2, -3, 1, -1.5
    got = oe.event_redistribute(st, 0)[0]
    assert got.shape == (8, 4)
    assert np.array_equal(got[:, [0, 1, 3]], np.array(
        [(0, 0, 1), (1, 1, -1), (0, 0, 1), (1, 1, -1), (0, 1, -1), (0, 1, -1), (0, 1, -1), (1, 0, 1)], np.float32))
    np.testing.assert_allclose(got[:, 2], [.0033333, .0033333, .33333334, .33333334, .33666667, .50166667,
                                           .6666667, .67], rtol=0, atol=1e-7)


def test_kat_events_to_channels_survey_8c():
    xs = np.array([0, 5, 5, 2.9, 6, -1, 3], np.float32)
    ys = np.array([0, 3, 3, 1.2, 1, 2, 4], np.float32)
    ps = np.array([1, -1, -1, 1, 1, 1, -1], np.float32)
    out = oe.events_to_channels(xs, ys, ps, (4, 6))
    pos, neg = np.zeros((4, 6), np.float32), np.zeros((4, 6), np.float32)
    pos[0, 0], pos[1, 2] = 1, 1
    neg[0, 0], neg[3, 5] = 1, 2          # neg[0,0] is the out-of-range quirk
    assert np.array_equal(out[0], pos) and np.array_equal(out[1], neg)
    assert xs[4] == 0 and xs[5] == 0 and xs[6] == 0 and ys[6] == 0   # caller's arrays are mutated


def test_events_to_channels_golden(golden_events):
    g = golden_events
    for i in range(int(g["n_e2c"])):
        xs, ys, ps = g[f"e2c{i}_xs"].copy(), g[f"e2c{i}_ys"].copy(), g[f"e2c{i}_ps"].copy()
        H, W = g[f"e2c{i}_hw"]
        out = oe.events_to_channels(xs, ys, ps, (H, W))
        assert np.array_equal(out, g[f"e2c{i}_out"]), i
        assert np.array_equal(xs, g[f"e2c{i}_xs_after"]) and np.array_equal(ys, g[f"e2c{i}_ys_after"])


def test_lift_golden(golden_events):
    g = golden_events
    for i in range(int(g["n_lift"])):
        H, W, k = g[f"lift{i}_dims"]
        xs = oe.lift_coords(g[f"lift{i}_xs"], W, W * k)
        ys = oe.lift_coords(g[f"lift{i}_ys"], H, H * k)
        out = oe.events_to_channels(xs, ys, g[f"lift{i}_ps"], (H * k, W * k))
        assert np.array_equal(out, g[f"lift{i}_out"]), i


@pytest.mark.parametrize("mode", [0, 1])
def test_cnt2event_golden(golden_events, mode):
    g = golden_events
    for i in range(int(g["n_c2e"])):
        assert np.array_equal(oe.cnt2event(g[f"c2e{i}_in"], mode), g[f"c2e{i}_out{mode}"]), i


@pytest.mark.parametrize("mode", [0, 1])
def test_event_redistribute_golden(golden_events, mode):
    g = golden_events
    for i in range(int(g["n_er"])):
        assert np.array_equal(oe.event_redistribute(g[f"er{i}_in"], mode), g[f"er{i}_out{mode}"]), i


def test_cnt2event_negative_raises():
    c = np.zeros((1, 2, 2, 2), np.float32)
    c[0, 0, 0, 0], c[0, 1, 1, 1] = 3, -1
    with pytest.raises(ValueError):
        oe.cnt2event(c, 0)


def test_stack_and_voxel_golden(golden_events):
    """events_to_stack_no_polarity (+ the reference's own binary search) and events_to_voxel, incl. in-place effects."""
    g = golden_events
    for i in range(int(g["n_stack"])):
        H, W, TB = (int(v) for v in g[f"stk{i}_dims"])
        xs, ys, ts, ps = (g[f"stk{i}_{k}"].copy() for k in ("xs", "ys", "ts", "ps"))
        out = oe.events_to_stack_no_polarity(xs, ys, ts, ps, TB, (H, W))
        assert np.array_equal(out, g[f"stk{i}_out"]), i
        assert np.array_equal(xs, g[f"stk{i}_xs_after"]) and np.array_equal(ps, g[f"stk{i}_ps_after"])
        xs, ys, ts, ps = (g[f"stk{i}_{k}"].copy() for k in ("xs", "ys", "ts", "ps"))
        vox = oe.events_to_voxel(xs, ys, ts, ps, max(TB, 2), (H, W))
        assert np.array_equal(vox, g[f"stk{i}_voxel"]), i
        assert np.array_equal(xs, g[f"stk{i}_voxel_xs_after"])
