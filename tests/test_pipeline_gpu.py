"""GPU tests of the end-to-end pipeline object (events -> counts -> network -> event lists): the software-pipelined
asynchronous API returns, per batch, exactly what the synchronous calls return, with and without the CUDA graph."""
import numpy as np
import pytest
import torch

from oracle import model_ref

pytestmark = pytest.mark.gpu


def _events(B, L, lr, n, seed):
    rng = np.random.default_rng(seed)
    F = B * L
    xs = torch.from_numpy(rng.integers(0, lr[1], F * n).astype(np.float32))
    ys = torch.from_numpy(rng.integers(0, lr[0], F * n).astype(np.float32))
    ps = torch.from_numpy(rng.choice(np.array([-1, 1], np.float32), F * n))
    off = torch.arange(0, F * n + 1, n, dtype=torch.int64)
    return [t.pin_memory() for t in (xs, ys, ps, off)]


@pytest.mark.parametrize("graph", [False, True])
def test_async_pipeline_equals_synchronous(graph):
    from esr_b200.model import DeepRecurrNet
    from esr_b200.pipeline import EventSRPipeline
    dev = torch.device("cuda:0")
    B, L, lr, scale, n = 2, 5, (32, 40), 2, 300
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    net.load_state_dict(model_ref.seeded_state_dict(2))
    net = net.to(dev).eval()
    pipe = EventSRPipeline(net, B, L, lr, scale, dev)
    g = torch.Generator().manual_seed(0)
    pipe.sr_bias = torch.poisson(torch.full(((L - 2) * B, 2, lr[0] * scale, lr[1] * scale), 0.4), generator=g).to(dev)
    batches = [_events(B, L, lr, n, s) for s in range(4)]
    want = [pipe.run_host(*b, n).clone() for b in batches]          # synchronous reference, one batch at a time
    if graph:
        pipe.run_device(*[t.to(dev) for t in batches[0]], n)
        pipe.capture()
        for b, w in zip(batches, want):
            assert torch.equal(pipe.run_host(*b, n), w)
    # software-pipelined: submit(i+1) before finish(i), collect(i-1) last
    got, pend_a, pend_b = [], None, None
    for b in batches:
        h = pipe.submit_host(*b, n)
        if pend_a is not None:
            hb = pipe.finish(pend_a)
            if pend_b is not None:
                got.append(pipe.collect(pend_b).clone())
            pend_b = hb
        pend_a = h
    for hdl in (pend_b, pend_a):
        got.append(pipe.collect(hdl).clone())
    assert len(got) == len(want)
    for i, (a, w) in enumerate(zip(got, want)):
        assert a.shape == w.shape and torch.equal(a, w), i
    # simple two-in-flight form (finish implied by collect)
    h0 = pipe.submit_host(*batches[1], n)
    assert torch.equal(pipe.collect(h0), want[1])


def test_graph_capacity_miss_falls_back_and_regrows():
    """The step graph records the fused redistribution with a row capacity and a largest-count guess taken from the data in the
    frame bank at capture time.  A batch that exceeds either must still return the right events (general chain on the same SR
    counts) and make the pipeline record larger graphs, after which the same batch is served by the graphs -- in the synchronous
    and in the software-pipelined form."""
    from esr_b200.model import DeepRecurrNet
    from esr_b200.pipeline import EventSRPipeline
    dev = torch.device("cuda:0")
    B, L, lr, scale, n = 2, 5, (32, 40), 2, 300
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    net.load_state_dict(model_ref.seeded_state_dict(2))
    net = net.to(dev).eval()
    pipe = EventSRPipeline(net, B, L, lr, scale, dev)
    g = torch.Generator().manual_seed(1)
    shape = ((L - 2) * B, 2, lr[0] * scale, lr[1] * scale)
    small = torch.poisson(torch.full(shape, 0.2), generator=g).to(dev)
    big = torch.poisson(torch.full(shape, 3.0), generator=g).to(dev)
    big[0, 0, 3, 5] += 40                                               # also beyond twice the largest count seen at capture time
    batch = _events(B, L, lr, n, 7)
    pipe.sr_bias = big
    want_big = pipe.run_host(*batch, n).clone()                          # eager reference
    pipe.sr_bias = small
    want_small = pipe.run_host(*batch, n).clone()
    pipe.capture(fused_rows=int(want_small.shape[0] * want_small.shape[1]) + 8, fused_max_count=4)
    assert torch.equal(pipe.run_host(*batch, n), want_small)
    cap0 = pipe._graphs[0]["fused"].cap
    pipe.sr_bias.copy_(big)                                              # same tensor object: the graphs read it
    assert torch.equal(pipe.run_host(*batch, n), want_big)               # capacity miss -> general chain, graphs re-recorded
    assert pipe._graphs[0]["fused"].cap > cap0 and pipe._graphs[0]["fused"].mcap >= 64
    assert torch.equal(pipe.run_host(*batch, n), want_big)               # now from the graphs
    # the pipelined form across a miss
    pipe.capture(fused_rows=int(want_small.shape[0] * want_small.shape[1]) + 8, fused_max_count=4)
    h1 = pipe.submit_host(*batch, n)
    h2 = pipe.submit_host(*batch, n)
    assert torch.equal(pipe.collect(h1), want_big)
    assert torch.equal(pipe.collect(h2), want_big)
    assert torch.equal(pipe.collect(pipe.submit_host(*batch, n)), want_big)
