"""Generate tests/golden/model_golden.npz by running the REFERENCE's own models/model.py (build container only).

The reference model is imported from /root/reference with the two stubs of SURVEY.md 8c:
  - sys.modules['myutils.vis_events.matplotlib_plot_events'] = empty module   (models/model.py:16 star-import)
  - sys.modules['_ext'] = shim whose dcn_v2_forward is torchvision.ops.deform_conv2d   (models/DCNv2/dcn_v2.py:13)
Weights: oracle.model_ref.seeded_state_dict(seed) loaded into the reference module (same key names), so the
fixtures only need to store seeds, input and output tensors.  Inputs: Poisson(0.1)/(0.4) count tensors.
Cases cover: multiple-of-8 and padded/cropped sizes, B>1, recurrent state carried over several windows,
reset_states(), and the zero-initialised conv_offset_mask the reference ships with.
"""
import os
import sys
import types

import numpy as np
import torch
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(1, "/root/reference")

sys.modules["myutils.vis_events.matplotlib_plot_events"] = types.ModuleType("stub")
ext = types.ModuleType("_ext")
ext.dcn_v2_forward = lambda inp, w, b, off, m, kh, kw, sh, sw, ph, pw, dh, dw, dg: \
    torchvision.ops.deform_conv2d(inp, off, w, b, stride=(sh, sw), padding=(ph, pw), dilation=(dh, dw), mask=m)
sys.modules["_ext"] = ext

from models.model import DeepRecurrNet  # noqa: E402  (the reference)
from oracle import model_ref  # noqa: E402

CASES = [
    # name, seed, B, H, W, lam, n_windows, zero_offset_init
    ("a", 0, 1, 32, 32, 0.1, 3, False),
    ("b", 1, 2, 36, 44, 0.4, 2, False),     # padded to 40x48 and cropped back
    ("c", 2, 1, 64, 48, 0.1, 1, True),      # conv_offset_mask zero-initialised like the shipped model
    ("d", 3, 2, 24, 40, 0.2, 2, False),
]


def make_input(seed, B, H, W, lam, n_windows):
    g = torch.Generator().manual_seed(1000 + seed)
    frames = torch.poisson(torch.full((B, n_windows + 2, 2, H, W), lam), generator=g)
    return frames


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    out = {"cases": np.array([c[0] for c in CASES])}
    for name, seed, B, H, W, lam, nwin, zero_off in CASES:
        sd = model_ref.seeded_state_dict(seed)
        if zero_off:
            for k in sd:
                if "conv_offset_mask" in k:
                    sd[k] = torch.zeros_like(sd[k])
        net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
        assert list(net.state_dict().keys()) == list(sd.keys()), "state_dict key order differs from the reference"
        net.load_state_dict(sd)
        net.eval()
        frames = make_input(seed, B, H, W, lam, nwin)
        outs = []
        with torch.no_grad():
            net.reset_states()
            for wdx in range(nwin):
                outs.append(net(frames[:, wdx:wdx + 3].contiguous()).clone())
            state_fwd = net.time_propagate.states[0].clone()
            net.reset_states()
            again = net(frames[:, 0:3].contiguous()).clone()
        assert torch.equal(again, outs[0])
        out[f"{name}_meta"] = np.array([seed, B, H, W, nwin, int(zero_off)])
        out[f"{name}_lam"] = np.array(lam)
        out[f"{name}_out"] = torch.stack(outs).numpy()
        out[f"{name}_state_fwd"] = state_fwd.numpy()[:, :4]   # a slice of the carried state
        # cross-check the restatement right here
        o = model_ref.OracleNet(sd)
        for wdx in range(nwin):
            got = o(frames[:, wdx:wdx + 3])
            err = (got - outs[wdx]).abs().max().item()
            print(name, wdx, "oracle vs reference max abs err", err, "ref max", outs[wdx].abs().max().item())
    path = os.path.join(HERE, "model_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path))


if __name__ == "__main__":
    main()
