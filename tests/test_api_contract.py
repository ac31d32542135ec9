

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
