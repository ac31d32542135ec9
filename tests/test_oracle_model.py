"""CPU: the torch fp32 restatement (oracle/model_ref.py) against fixtures produced by the reference's own
models/model.py (tests/golden/make_golden_model.py).  This pins the model oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import model_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))


def golden_case(g, name):
    seed, B, H, W, nwin, zero_off = (int(v) for v in g[f"{name}_meta"])
    lam = float(g[f"{name}_lam"])
    sd = model_ref.seeded_state_dict(seed)
    if zero_off:
        sd = {k: (torch.zeros_like(v) if "conv_offset_mask" in k else v) for k, v in sd.items()}
    gen = torch.Generator().manual_seed(1000 + seed)
    frames = torch.poisson(torch.full((B, nwin + 2, 2, H, W), lam), generator=gen)
    return sd, frames, nwin


def test_param_inventory_matches_reference_contract():
    shapes = model_ref.param_shapes()
    assert len(shapes) == 68
    assert sum(int(np.prod(s)) for s in shapes.values()) == 1813120


@pytest.mark.parametrize("name", ["a", "b", "c", "d"])
def test_oracle_matches_reference_fixtures(golden, name):
    torch.set_num_threads(8)
    sd, frames, nwin = golden_case(golden, name)
    net = model_ref.OracleNet(sd)
    want = golden[f"{name}_out"]
    for w in range(nwin):
        got = net(frames[:, w:w + 3]).numpy()
        assert got.shape == want[w].shape
        np.testing.assert_allclose(got, want[w], rtol=0, atol=2e-6 + 1e-5 * np.abs(want[w]).max())
    np.testing.assert_allclose(net.states[0].numpy()[:, :4], golden[f"{name}_state_fwd"], rtol=0, atol=1e-5)


def test_dcn_zero_offset_identity():
    """The reference's own known-answer test (models/DCNv2/testcuda.py:32-67): zero offsets, mask 0.5,
    identity 3x3 weights => 2*out == in."""
    B, C, H, W = 2, 8, 7, 9
    x = torch.randn(B, C, H, W)
    w = torch.zeros(C, C, 3, 3)
    for c in range(C):
        w[c, c, 1, 1] = 1.0
    out = model_ref.dcn_v2_forward(x, w, torch.zeros(C), torch.zeros(B, 2 * 18, H, W),
                                   torch.full((B, 2 * 9, H, W), 0.5), 2)
    assert (2 * out - x).abs().max().item() < 1e-10


def test_dcn_matches_torchvision():
    import torchvision
    torch.manual_seed(3)
    B, C, H, W, G = 2, 64, 12, 10, 8
    x, w, b = torch.randn(B, C, H, W), torch.randn(C, C, 3, 3) * 0.05, torch.randn(C) * 0.1
    off, m = torch.randn(B, G * 18, H, W) * 2.5, torch.rand(B, G * 9, H, W)
    want = torchvision.ops.deform_conv2d(x, off, w, b, stride=1, padding=1, dilation=1, mask=m)
    got = model_ref.dcn_v2_forward(x, w, b, off, m, G)
    assert (got - want).abs().max().item() < 2e-5
