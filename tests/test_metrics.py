"""Evaluation metrics: the CPU restatement's known answers (no GPU) and GPU parity against it.
Tolerance: the GPU path accumulates in fp64 like the restatement; 1e-9 relative on every scalar (summation order only)."""
import math

import numpy as np
import pytest
import torch

from oracle import metrics as om


def _pair(C, H, W, seed, lam=0.3):
    rng = np.random.default_rng(seed)
    tgt = rng.poisson(lam, (1, C, H, W)).astype(np.float32)
    pred = np.maximum(tgt + rng.normal(0, 0.4, tgt.shape), 0).astype(np.float32)
    return pred, tgt


def test_oracle_known_answers():
    pred, tgt = _pair(2, 40, 48, 0)
    assert om.ssim_loss(tgt, tgt) == pytest.approx(1.0, abs=1e-12)
    assert om.psnr_loss(tgt, tgt) == float("inf")
    # constant offset d on a constant image c: ux = c + d, uy = c, all (co)variances 0
    c, d = 3.0, 0.5
    a, b = np.full((32, 32), c + d, np.float32), np.full((32, 32), c, np.float32)
    C1, C2 = (0.01 * 2) ** 2, (0.03 * 2) ** 2
    want = ((2 * (c + d) * c + C1) * C2) / (((c + d) ** 2 + c ** 2 + C1) * C2)
    assert om.structural_similarity(a, b) == pytest.approx(want, rel=1e-12)
    assert om.peak_signal_noise_ratio(b, a, data_range=4.0) == pytest.approx(10 * math.log10(16 / d ** 2), rel=1e-12)
    # the reference's channel logic: range = tgt[c].max() - tgt.min() with the minimum over ALL channels (restore.py:80)
    t = np.zeros((1, 2, 16, 16), np.float32); t[0, 0] += 1.0; t[0, 0, 0, 0] = 5.0; t[0, 1, 0, 0] = 2.0
    p = t + 0.25
    want = (10 * math.log10((5.0 - 0.0) ** 2 / 0.0625) + 10 * math.log10((2.0 - 0.0) ** 2 / 0.0625)) / 2
    assert om.psnr_loss(p, t) == pytest.approx(want, rel=1e-9)
    with pytest.raises(ValueError):
        om.structural_similarity(np.zeros((5, 5), np.float32), np.zeros((5, 5), np.float32))


@pytest.mark.gpu
@pytest.mark.parametrize("C,H,W", [(2, 256, 256), (2, 90, 160), (1, 64, 64), (3, 37, 53), (2, 7, 7), (2, 1024, 1024)])
def test_gpu_metrics_vs_oracle(C, H, W):
    from esr_b200 import metrics as gm
    dev = torch.device("cuda:0")
    pred, tgt = _pair(C, H, W, C * 1000 + H)
    p, t = torch.from_numpy(pred).to(dev), torch.from_numpy(tgt).to(dev)
    assert gm.ssim_loss()(p, t) == pytest.approx(om.ssim_loss(pred, tgt), rel=1e-9)
    assert gm.psnr_loss()(p, t) == pytest.approx(om.psnr_loss(pred, tgt), rel=1e-9)
    assert float(gm.l1(p, t)) == pytest.approx(om.l1(pred, tgt), rel=1e-6)
    assert float(gm.mse(p, t)) == pytest.approx(om.mse(pred, tgt), rel=1e-6)
    ev = gm.evaluate(p, t)
    assert ev["ssim"] == pytest.approx(om.ssim_loss(pred, tgt), rel=1e-9) and ev["l1"] == pytest.approx(om.l1(pred, tgt), rel=1e-9)
    assert ev["psnr"] == pytest.approx(om.psnr_loss(pred, tgt), rel=1e-9) and ev["mse"] == pytest.approx(om.mse(pred, tgt), rel=1e-9)
    assert gm.ssim_loss()(t, t) == pytest.approx(1.0, abs=1e-12) and gm.psnr_loss()(t, t) == float("inf")


@pytest.mark.gpu
def test_gpu_metrics_errors_and_batches():
    from esr_b200 import _lib
    from esr_b200 import metrics as gm
    dev = torch.device("cuda:0")
    with pytest.raises(_lib.ESRError):
        gm.ssim_loss()(torch.zeros(1, 2, 16, 16), torch.zeros(1, 2, 16, 16))          # CPU tensors: no fallback
    with pytest.raises(_lib.ESRError):
        gm.ssim_loss()(torch.zeros(1, 2, 5, 5, device=dev), torch.zeros(1, 2, 5, 5, device=dev))   # win_size exceeds image extent
    # a batch: mean over samples of the per-sample reference value
    preds, tgts = zip(*[_pair(2, 48, 64, 50 + i) for i in range(3)])
    P, T = np.concatenate(preds), np.concatenate(tgts)
    ev = gm.evaluate(torch.from_numpy(P).to(dev), torch.from_numpy(T).to(dev))
    assert ev["ssim"] == pytest.approx(np.mean([om.ssim_loss(p, t) for p, t in zip(preds, tgts)]), rel=1e-9)
    assert ev["psnr"] == pytest.approx(np.mean([om.psnr_loss(p, t) for p, t in zip(preds, tgts)]), rel=1e-9)
