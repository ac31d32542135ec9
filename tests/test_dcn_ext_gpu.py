"""GPU: the `_ext.dcn_v2_forward` drop-in against the oracle restatement and the reference's own known-answer test."""
import pytest
import torch

from oracle import model_ref

pytestmark = pytest.mark.gpu


def test_zero_offset_identity_kat():
    """models/DCNv2/testcuda.py:32-67 (check_zero_offset): zero offsets, mask 0.5, identity weights => 2*out == in."""
    from esr_b200 import dcn_v2_ext as ext
    dev = torch.device("cuda:0")
    B, C, H, W, G = 2, 64, 16, 24, 8
    x = torch.randn(B, C, H, W, device=dev)
    w = torch.zeros(C, C, 3, 3, device=dev)
    w[torch.arange(C), torch.arange(C), 1, 1] = 1.0
    out = ext.dcn_v2_forward(x, w, torch.zeros(C, device=dev), torch.zeros(B, G * 18, H, W, device=dev),
                             torch.full((B, G * 9, H, W), 0.5, device=dev), 3, 3, 1, 1, 1, 1, 1, 1, G)
    # split-bf16 storage carries ~2^-17 relative error; the reference KAT tolerance of 1e-10 assumes exact fp32
    assert ((2 * out - x).abs().max() / x.abs().max()).item() < 2e-5


@pytest.mark.parametrize("H,W,scale", [(16, 16, 0.5), (32, 32, 3.0), (12, 20, 8.0)])
def test_against_oracle(H, W, scale):
    from esr_b200 import dcn_v2_ext as ext
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H + W)
    B, C, G = 2, 64, 8
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / 24
    b = torch.randn(C, generator=g) * 0.1
    off = torch.randn(B, G * 18, H, W, generator=g) * scale        # large offsets exercise the image borders
    m = torch.rand(B, G * 9, H, W, generator=g)
    want = model_ref.dcn_v2_forward(x, w, b, off, m, G)
    got = ext.dcn_v2_forward(*(t.to(dev) for t in (x, w, b, off, m)), 3, 3, 1, 1, 1, 1, 1, 1, G).cpu()
    assert ((got - want).abs().max() / want.abs().max()).item() < 1e-4


def test_invalid_config_raises():
    """every configuration the reference operator accepts is served (tests/test_dropin.py); what it rejects raises RuntimeError:
    channels not divisible by the groups (dcn_v2_cuda.cu:38-62 AT_ASSERTM), kernel / weight shape mismatch, CPU tensors"""
    from esr_b200 import dcn_v2_ext as ext
    dev = torch.device("cuda:0")
    x = torch.randn(1, 32, 8, 8, device=dev)
    with pytest.raises(RuntimeError):
        ext.dcn_v2_forward(x, torch.randn(32, 32, 3, 3, device=dev), torch.zeros(32, device=dev),
                           torch.zeros(1, 54, 8, 8, device=dev), torch.zeros(1, 27, 8, 8, device=dev), 3, 3, 1, 1, 1, 1, 1, 1, 3)
    with pytest.raises(RuntimeError):
        ext.dcn_v2_forward(x, torch.randn(32, 32, 3, 3, device=dev), torch.zeros(32, device=dev),
                           torch.zeros(1, 18, 8, 8, device=dev), torch.zeros(1, 9, 8, 8, device=dev), 5, 5, 1, 1, 2, 2, 1, 1, 1)
    with pytest.raises(RuntimeError):
        ext.dcn_v2_forward(x.cpu(), torch.randn(32, 32, 3, 3), torch.zeros(32), torch.zeros(1, 18, 8, 8), torch.zeros(1, 9, 8, 8),
                           3, 3, 1, 1, 1, 1, 1, 1, 1)


@pytest.mark.parametrize("H,W,scale", [(16, 16, 0.5), (12, 20, 3.0)])
def test_backward_against_autograd_of_oracle(H, W, scale):
    """`_ext.dcn_v2_backward` (models/DCNv2/dcn_v2.py:46-68) vs torch.autograd through the oracle restatement."""
    from esr_b200 import dcn_v2_ext as ext
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(H * 7 + W)
    B, C, G = 2, 64, 8
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / 24
    b = torch.randn(C, generator=g) * 0.1
    off = torch.randn(B, G * 18, H, W, generator=g) * scale
    m = torch.rand(B, G * 9, H, W, generator=g)
    go = torch.randn(B, C, H, W, generator=g)
    leaves = [t.clone().requires_grad_(True) for t in (x, off, m, w, b)]
    out = model_ref.dcn_v2_forward(leaves[0], leaves[3], leaves[4], leaves[1], leaves[2], G)
    want = torch.autograd.grad(out, leaves, go)                        # d/d(input, offset, mask, weight, bias)
    got = ext.dcn_v2_backward(*(t.to(dev) for t in (x, w, b, off, m, go)), 3, 3, 1, 1, 1, 1, 1, 1, G)
    names = ["grad_input", "grad_offset", "grad_mask", "grad_weight", "grad_bias"]
    for name, gt, wt in zip(names, got, want):
        rel = ((gt.cpu() - wt).abs().max() / wt.abs().max().clamp_min(1e-12)).item()
        assert rel < 2e-4, (name, rel)
