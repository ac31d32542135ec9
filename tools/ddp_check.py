"""2-GPU check (torchrun): the reference's training recipe -- DDP-wrapped model, per-window forward loop, summed MSE,
loss.backward(), torch.optim.Adam -- runs on esr_b200.DeepRecurrNet unchanged, and the gradients DDP all-reduces are the
mean of the per-rank gradients of the same operators.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/ddp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist
import torch.nn as nn
from torch.nn.parallel import DistributedDataParallel as DDP

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from esr_b200.model import DeepRecurrNet  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = bench.synth_weights(0)

    def make():
        net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
        net.load_state_dict(sd)
        return net.to(dev)

    B, L, H, W = 2, 5, 64, 64
    g = torch.Generator().manual_seed(10 + rank)                       # different data per rank
    frames = torch.poisson(torch.full((B, L, 2, H, W), 0.2), generator=g).to(dev)
    gt = torch.poisson(torch.full((B, L, 2, H, W), 0.2), generator=g).to(dev)

    def loop(model, net):
        net.reset_states()
        loss = 0
        for w in range(L - 2):                                         # train_ours_cnt_seq.py:217-231
            loss = loss + nn.functional.mse_loss(model(frames[:, w:w + 3]), gt[:, w + 1])
        loss.backward()
        return loss

    solo = make()
    loop(solo, solo)
    local_grads = torch.cat([p.grad.reshape(-1) for p in solo.parameters()])
    dist.all_reduce(local_grads)
    local_grads /= world                                               # what DDP must produce

    net = make()
    ddp = DDP(net, device_ids=[local], output_device=local)
    opt = torch.optim.Adam(ddp.parameters(), lr=1e-3, weight_decay=1e-4, amsgrad=True)
    opt.zero_grad()
    loss = loop(ddp, net)
    ddp_grads = torch.cat([p.grad.reshape(-1) for p in net.parameters()])
    rel = ((ddp_grads - local_grads).abs().max() / local_grads.abs().max()).item()
    opt.step()
    with torch.no_grad():
        net.reset_states()
        out = net(frames[:, 0:3].contiguous())                         # inference plan repacks the updated parameters
    ok = rel < 1e-4 and torch.isfinite(out).all().item() and torch.isfinite(loss).item()
    t = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ddp_check: world={world} loss={loss.item():.6f} grad_rel_vs_manual_mean={rel:.2e} ->", "OK" if t.item() == 1.0 else "FAIL")
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
