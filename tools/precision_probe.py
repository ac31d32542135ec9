"""CPU probe (oracle only, no GPU): how much output error does rounding the conv WEIGHTS to fp16 / bf16 cause?  It is the
error floor of any 2-pass product (activations hi+lo, weights one 16-bit plane).  Result (profiles/r1_notes.md): fp16 0.6e-3 ..
1.7e-3, bf16 5e-3 .. 8e-3 of the output max-norm -- above the 1e-3 parity bar, so the 3-pass split product stays.

    python tools/precision_probe.py
"""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import bench
from oracle import model_ref
torch.set_num_threads(8)
torch.manual_seed(0)
def run(sd, frames):
    states, outs = None, []
    for w in range(frames.shape[1] - 2):
        y, states = model_ref.forward(sd, frames[:, w:w + 3], states)
        outs.append(y)
    return torch.stack(outs)
for seed in (0, 1):
    sd = bench.synth_weights(seed)
    g = torch.Generator().manual_seed(seed)
    frames = torch.poisson(torch.full((1, 6, 2, 128, 128), 0.3), generator=g)
    with torch.no_grad():
        ref = run({k: v.double() for k, v in sd.items()}, frames.double())
        f32 = run(sd, frames)
        def rnd(dt):
            return {k: (v.to(dt).float() if v.dim() == 4 else v) for k, v in sd.items()}
        h16 = run(rnd(torch.float16), frames)
        b16 = run(rnd(torch.bfloat16), frames)
    den = ref.abs().max()
    print(seed, 'max|ref|', den.item(), 'fp32 err', ((f32 - ref).abs().max() / den).item(), 'w->fp16 err', ((h16 - ref).abs().max() / den).item(),
          'w->bf16 err', ((b16 - ref).abs().max() / den).item())
    # per-window
    print(' per-window fp16:', [round(((h16[i] - ref[i]).abs().max() / ref[i].abs().max()).item(), 6) for i in range(ref.shape[0])])
