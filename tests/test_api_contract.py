

def test_fused_cnt2event_sizing_rules_without_a_gpu():
    """The host half of the fused redistribution (esr_b200.expand.FusedCnt2Event.result): the reference's sizing rules of
    cnt2event.pyx:33-60 applied to the statistics the kernels leave behind -- an all-empty call yields [B, 1, 4] zeros, a sample whose
    rounded counts sum to zero contributes length 1, an active sample with a negative count raises like np.zeros([-n, 4]), and a call
    outside the recorded capacity / count limit returns None so that the caller takes the general chain."""
    import numpy as np
    import torch
    from esr_b200.expand import FusedCnt2Event
    f = FusedCnt2Event.__new__(FusedCnt2Event)
    f.B, f.H, f.W, f.dev = 3, 4, 4, torch.device("cpu")
    f.cap, f.mcap = 30, 8
    f.out = torch.arange(f.cap * 4, dtype=torch.float32)

    def stats(rows):
        f.stats_host = torch.tensor(rows, dtype=torch.int64)
    stats([[0, 0, 0, 0]] * 3)
    ev, rows, mx = f.result()
    assert tuple(ev.shape) == (3, 1, 4) and not ev.any() and rows == 3
    stats([[5, 5, 0, 2], [0, 7, 1, 3], [9, 9, 0, 4]])          # sample 1: +/- cancel -> inactive, its events and its negative flag ignored
    ev, rows, mx = f.result()
    assert tuple(ev.shape) == (3, 9, 4) and rows == 27 and mx == 4 and ev.data_ptr() == f.out.data_ptr()
    stats([[5, 5, 0, 2], [3, 4, 1, 3], [0, 0, 0, 0]])          # active sample with a negative count
    import pytest
    with pytest.raises(ValueError):
        f.result()
    stats([[11, 11, 0, 2], [1, 1, 0, 1], [2, 2, 0, 1]])        # 3 * 11 rows > capacity 30
    assert f.result()[0] is None
    stats([[5, 5, 0, 9], [1, 1, 0, 1], [2, 2, 0, 1]])          # a count above the recorded limit
    assert f.result() == (None, 15, 9)
    assert np.random.get_state()[1][0] == np.random.RandomState(123).get_state()[1][0]     # the reference's reseeding side effect


def test_fused_cnt2event_key_tables():
    """The tables the fused redistribution sorts by (esr_b200.expand._xf_tables_host): for every count n <= m and event j < n the key
    indexes exactly the timestamp the reference assigns, float32(np.linspace(0, 1, n)[j]) (cnt2event.pyx:74); keys are strictly
    increasing in j (so a pixel emits a key at most once -- what lets the kernels derive key histograms from count histograms) and
    ascend with the timestamp (so the counting sort over keys is the reference's stable sort by time)."""
    import numpy as np
    from esr_b200.expand import _xf_tables_host
    blob, desc = _xf_tables_host()
    ks = []
    for i in range(7):
        m = 1 << i
        ro, uo, K = (int(v) for v in desc[i])
        assert ro % 16 == 0 and uo % 16 == 0
        rank = np.frombuffer(blob, dtype=np.uint16, count=(m + 1) * m, offset=ro).reshape(m + 1, m)
        uniq = np.frombuffer(blob, dtype=np.float32, count=K, offset=uo)
        assert np.all(np.diff(uniq) > 0) and uniq[0] == 0.0 and (m == 1 or uniq[-1] == 1.0)
        for n in range(1, m + 1):
            want = np.linspace(0, 1, n).astype(np.float32)
            assert np.array_equal(uniq[rank[n, :n]], want), (m, n)
            assert np.all(np.diff(rank[n, :n].astype(np.int64)) > 0), (m, n)
        ks.append(K)
    assert ks == [1, 2, 5, 19, 73, 309, 1229]          # the kernels size their per-key counters by these (XF_MAXK = 1232)
