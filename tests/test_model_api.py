"""CPU: the product's DeepRecurrNet keeps the reference's constructor / state_dict contract."""
import pytest
import torch

from oracle import model_ref


def test_state_dict_keys_shapes_and_count():
    from esr_b200.model import DeepRecurrNet
    net = DeepRecurrNet(inch=2, basech=8, num_frame=3)
    sd = net.state_dict()
    want = model_ref.param_shapes(2, 8, 3)
    assert list(sd.keys()) == list(want.keys())
    for k, v in sd.items():
        assert tuple(v.shape) == tuple(want[k]), k
    assert sum(v.numel() for v in sd.values()) == 1813120
    net.load_state_dict(model_ref.seeded_state_dict(0))      # a reference-format checkpoint loads


def test_default_constructor_matches_reference_default_width():
    from esr_b200.model import DeepRecurrNet
    net = DeepRecurrNet()                                     # basech=16 default (models/model.py:295)
    assert sum(p.numel() for p in net.parameters()) == 6995488


def test_reference_init_conventions():
    from esr_b200.model import DeepRecurrNet
    net = DeepRecurrNet(basech=8)
    assert net.spacetime_fuse.dcn.conv_offset_mask.weight.abs().max().item() == 0.0      # dcn_v2.py:210-212
    assert net.spacetime_fuse.dcn.bias.abs().max().item() == 0.0
    g = net.time_propagate.lstm.recurrent_block.update_gate
    w = g.weight.view(64, -1)
    assert torch.allclose(w @ w.t(), torch.eye(64), atol=1e-4)                            # orthogonal init
    assert g.bias.abs().max().item() == 0.0


def test_no_cpu_fallback_and_unsupported_config():
    from esr_b200 import _lib
    from esr_b200.model import DeepRecurrNet
    net = DeepRecurrNet(basech=8)
    with torch.no_grad(), pytest.raises(_lib.ESRError):
        net(torch.zeros(1, 3, 2, 16, 16))                     # CPU tensor
    with torch.no_grad(), pytest.raises(_lib.ESRError):
        DeepRecurrNet(basech=16)(torch.zeros(1, 3, 2, 16, 16))
