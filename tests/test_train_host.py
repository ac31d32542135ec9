"""CPU tests of the training path's HOST logic (no GPU, no kernels): the window / pair / slot bookkeeping of
esr_b200.train.forward_sequence, the deferred ConvGRU weight-gradient batching and the N>1 gradient exchange, with the
CUDA operators swapped for plain torch ops of the same maths.  (The operators themselves are covered by the `-m gpu`
parity tests in tests/test_train_gpu.py.)"""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import model_ref

torch.set_num_threads(min(8, torch.get_num_threads()))
_ACTS = {0: lambda v: v, 1: torch.relu, 2: torch.sigmoid, 3: torch.tanh}


def _patch_ops(train):
    """Route esr_b200.train's operator entry points to torch CPU equivalents; returns an undo function."""
    saved = {k: getattr(train, k) for k in ("conv2d", "dcn_v2", "upsample2x", "mse_loss", "_GruHRFn", "_GruBlendFn")}

    def conv2d(x, w, b, stride=1, act=None, defer=None):
        return _ACTS[train._ACT[act]](F.conv2d(x, w, b, stride=stride, padding=w.shape[-1] // 2))

    class HR:
        @staticmethod
        def apply(h, zr):
            return h * zr[:, h.shape[1]:]

    class Blend:
        @staticmethod
        def apply(h, zr, o):
            z = zr[:, :h.shape[1]]
            return h * (1 - z) + o * z

    train.conv2d = conv2d
    train.dcn_v2 = lambda inp, off, msk, w, b, dg=8: model_ref.dcn_v2_forward(inp, w, b, off, msk, dg)
    train.upsample2x = lambda x: F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    train.mse_loss = F.mse_loss
    train._GruHRFn, train._GruBlendFn = HR, Blend

    def undo():
        for k, v in saved.items():
            setattr(train, k, v)
    return undo


def _net(sd):
    from esr_b200.model import DeepRecurrNet
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    net.load_state_dict(sd)
    return net


def _data(B, L, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    return (torch.poisson(torch.full((B, L, 2, H, W), 0.3), generator=g), torch.poisson(torch.full((B, L, 2, H, W), 0.3), generator=g))


def _oracle_loss_and_grads(sd, frames, gt):
    ref = {k: v.clone().requires_grad_() for k, v in sd.items()}
    states, loss, preds = None, 0, []
    for w in range(frames.shape[1] - 2):
        pred, states = model_ref.forward(ref, frames[:, w:w + 3], states)
        preds.append(pred)
        loss = loss + F.mse_loss(pred, gt[:, w + 1])
    loss.backward()
    return loss.item(), {k: v.grad for k, v in ref.items()}, torch.cat(preds, 0).detach(), [s.detach() for s in states]


@pytest.mark.parametrize("B,L,H,W", [(2, 5, 16, 24), (1, 3, 20, 12), (1, 6, 18, 16)])
def test_forward_sequence_bookkeeping_vs_oracle(B, L, H, W):
    """All windows as one graph (pairs, slots, reversed chain, window-major output, carried state) == the reference's loop of
    windows, for predictions, final states, loss and all 68 gradients -- incl. a non-multiple-of-8 size (CropSize)."""
    from esr_b200 import train
    sd = model_ref.seeded_state_dict(3 + L)
    frames, gt = _data(B, L, H, W, 5)
    want_loss, want_grads, want_pred, want_states = _oracle_loss_and_grads(sd, frames, gt)
    undo = _patch_ops(train)
    try:
        net = _net(sd)
        pred, states = train.forward_sequence(net, frames, None)
        Wn = L - 2
        loss = Wn * F.mse_loss(pred, gt[:, 1:1 + Wn].transpose(0, 1).reshape(pred.shape))
        loss.backward()
    finally:
        undo()
    assert pred.shape == want_pred.shape and torch.allclose(pred, want_pred, rtol=1e-4, atol=1e-6)
    for a, b in zip(states, want_states):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-6)
    assert abs(loss.item() - want_loss) <= 1e-5 * abs(want_loss)
    for name, p in net.named_parameters():
        g = want_grads[name]
        assert (p.grad - g).abs().max() <= 1e-4 * g.abs().max() + 1e-9, name


def test_carried_state_across_calls_and_reset():
    from esr_b200 import train
    sd = model_ref.seeded_state_dict(9)
    frames, _ = _data(1, 5, 16, 16, 2)
    undo = _patch_ops(train)
    try:
        net = _net(sd)
        full, _ = train.forward_sequence(net, frames, None)
        a, st = train.forward_sequence(net, frames[:, 0:4], None)            # windows 0, 1
        b, _ = train.forward_sequence(net, frames[:, 2:5], st)               # window 2 with the carried state
    finally:
        undo()
    assert torch.allclose(torch.cat([a, b], 0), full, rtol=1e-5, atol=1e-7)


def _free_port():
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    torch.set_num_threads(2)
    from esr_b200 import dist as ed
    from esr_b200 import train
    ed.init_from_env("gloo")
    _patch_ops(train)
    sd = model_ref.seeded_state_dict(4)
    frames, gt = _data(2, 4, 16, 16, 21)                                     # global batch of 2 sequences, one per rank
    lo, hi = ed.shard_range(2, world, rank)
    net = _net(sd)
    pred, _ = train.forward_sequence(net, frames[lo:hi], None)
    loss = 2 * F.mse_loss(pred, gt[lo:hi, 1:3].transpose(0, 1).reshape(pred.shape))
    loss.backward()
    grads = [p.grad for p in net.parameters()]
    ed.flat_allreduce_(grads, average=True)                                  # the step's single exchange (DDP's role)
    q.put((rank, torch.cat([g.reshape(-1) for g in grads]).numpy()))       # numpy: pickled by value, no fd hand-back to an exited worker
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_exchange_gloo():
    """world_size 2 on CPU: per-rank gradients of a batch shard, averaged through the flat bucket, equal the gradient of the
    mean-over-ranks loss computed in one process (what DDP gives the reference)."""
    from esr_b200 import train
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: torch.from_numpy(a) for r, a in (q.get(timeout=300) for _ in range(2))}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert torch.equal(res[0], res[1])
    sd = model_ref.seeded_state_dict(4)
    frames, gt = _data(2, 4, 16, 16, 21)
    undo = _patch_ops(train)
    try:
        net = _net(sd)
        total = 0
        for b in range(2):                                                   # DDP averages per-rank losses
            pred, _ = train.forward_sequence(net, frames[b:b + 1], None)
            total = total + 0.5 * 2 * F.mse_loss(pred, gt[b:b + 1, 1:3].transpose(0, 1).reshape(pred.shape))
        total.backward()
    finally:
        undo()
    want = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    assert (res[0] - want).abs().max() <= 1e-5 * want.abs().max()
