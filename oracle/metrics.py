"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the evaluation metrics of infer_ours_cnt.py:81-100.

Follows loss/restore.py:42-61 (ssim_loss) and :64-90 (psnr_loss) line by line; the two skimage functions they call are a
THIRD-PARTY dependency absent from /root/reference and from this image (scikit-image; the reference pins no version, its
shipped binaries are cp37 / cp38, i.e. scikit-image 0.16 - 0.19).  Their published algorithms are restated here with numpy +
scipy.ndimage.uniform_filter (the primitive skimage itself calls):

  skimage.metrics.structural_similarity(im1, im2)      [float images, defaults]
      win_size = 7, uniform window, use_sample_covariance = True, K1 = 0.01, K2 = 0.03,
      data_range = dtype range of float images = 2, float64 arithmetic,
      S = ((2 ux uy + C1)(2 vxy + C2)) / ((ux^2 + uy^2 + C1)(vx + vy + C2)), mean over S cropped by (win_size - 1) // 2
  skimage.metrics.peak_signal_noise_ratio(image_true, image_test, data_range)
      10 log10(data_range^2 / mean((true - test)^2, float64)); data_range=None with a float image: 1 if true.min() >= 0 else 2

PARITY UNPINNED: neither scikit-image nor any stored metric value exists in the reference, so this restatement cannot be
checked against the reference's own output here; it is anchored on the call sites above and on known answers
(identical images -> SSIM 1, PSNR inf; constant offset -> closed forms) in tests/test_metrics.py.
"""
import numpy as np
from scipy.ndimage import uniform_filter


def structural_similarity(im1, im2, win_size=7, data_range=2.0, K1=0.01, K2=0.03):
    X, Y = im1.astype(np.float64), im2.astype(np.float64)
    if min(X.shape) < win_size:
        raise ValueError("win_size exceeds image extent.")
    NP = win_size ** X.ndim
    cov_norm = NP / (NP - 1)
    ux, uy = uniform_filter(X, size=win_size), uniform_filter(Y, size=win_size)
    uxx, uyy, uxy = uniform_filter(X * X, size=win_size), uniform_filter(Y * Y, size=win_size), uniform_filter(X * Y, size=win_size)
    vx, vy, vxy = cov_norm * (uxx - ux * ux), cov_norm * (uyy - uy * uy), cov_norm * (uxy - ux * uy)
    C1, C2 = (K1 * data_range) ** 2, (K2 * data_range) ** 2
    S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
    pad = (win_size - 1) // 2
    return float(S[tuple(slice(pad, -pad) for _ in range(X.ndim))].mean(dtype=np.float64))


def peak_signal_noise_ratio(image_true, image_test, data_range=None):
    if data_range is None:
        data_range = 1.0 if image_true.min() >= 0 else 2.0
    err = np.mean((image_true.astype(np.float64) - image_test.astype(np.float64)) ** 2, dtype=np.float64)
    return float(10 * np.log10((data_range ** 2) / err)) if err > 0 else float("inf")


def ssim_loss(pred, tgt):
    """loss/restore.py:46-61; pred, tgt numpy, 1xNxHxW"""
    pred, tgt = np.squeeze(pred), np.squeeze(tgt)
    if pred.ndim == 3:
        return sum(structural_similarity(pred[i], tgt[i]) for i in range(pred.shape[0])) / pred.shape[0]
    return structural_similarity(pred, tgt)


def psnr_loss(pred, tgt):
    """loss/restore.py:68-90"""
    pred, tgt = np.squeeze(pred), np.squeeze(tgt)
    if pred.ndim == 3:
        loss = 0.0
        for i in range(pred.shape[0]):
            loss += peak_signal_noise_ratio(tgt[i], pred[i], data_range=tgt[i].max() - tgt.min())
        return loss / pred.shape[0]
    return peak_signal_noise_ratio(pred.clip(0, 1), tgt.clip(0, 1))


def l1(pred, tgt):
    return float(np.abs(pred.astype(np.float64) - tgt.astype(np.float64)).mean())


def mse(pred, tgt):
    return float(((pred.astype(np.float64) - tgt.astype(np.float64)) ** 2).mean())
