"""TEST INFRASTRUCTURE (oracle): CPU restatement of the per-frame tensor factory of the reference's dataset,
H5Dataset.__getitem__ (dataloader/h5dataset.py:276-406), on numpy + the C oracle encodings (oracle/events.py) + torch's CPU
F.interpolate (the implementation the reference itself calls).  Pinned against tests/golden/items_golden.npz, which the
reference's own unmodified __getitem__ produced (tests/golden/make_golden_items.py).  Only tests/ and the cpu_baseline leg of
tools/bench_items.py may import this module; the product path is esr_b200.dataset.create_item.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import events as oe


def event_formatting(events):
    """dataloader/base_dataset.py:26-33: float32, t normalised by (t - t0) / (t_last - t0 + 1e-6) in fp32."""
    xs, ys, ts, ps = (np.asarray(events[c]).astype(np.float32) for c in range(4))
    ts = ((ts - ts[0]) / (ts[-1] - ts[0] + np.float32(1e-6))).astype(np.float32)
    return np.stack([xs, ys, ts, ps])


def create_normalized_events(events, res):
    """dataloader/h5dataset.py:508-518 (fp32 divisions)."""
    return np.stack([events[0] / np.float32(res[1]), events[1] / np.float32(res[0]), events[2], events[3]]).astype(np.float32)


def _scaled_cnt(norm, res):
    """create_scaled_encoding(..., 'cnt') (h5dataset.py:526-528): fresh products, so the caller's events are not modified."""
    return oe.events_to_channels((norm[0] * np.float32(res[1])).astype(np.float32), (norm[1] * np.float32(res[0])).astype(np.float32),
                                 norm[3].copy(), res)


def _interp(x, size, mode):
    kw = {"align_corners": False} if mode == "bicubic" else {}
    return F.interpolate(torch.from_numpy(x).unsqueeze(0), size=tuple(size), mode=mode, **kw).squeeze(0).numpy()


def create_item(inp_events, gt_events, inp_res, scale, time_bins=1):
    """The event-derived entries of the item dict, in the reference's call order (its encodings zero out-of-range events in
    place, encodings.py:251-256, and later encodings of the same tensor see that)."""
    inp_res = [int(v) for v in inp_res]
    gt_res = [round(i * scale) for i in inp_res]
    down = [round(i / scale) for i in inp_res]
    ev = event_formatting(inp_events)                                   # rows are views: in-place effects persist
    gt = event_formatting(gt_events)
    item = {}
    item["inp_stack"] = oe.events_to_stack_no_polarity(ev[0], ev[1], ev[2], ev[3], time_bins, inp_res)      # :339
    item["inp_cnt"] = oe.events_to_channels(ev[0], ev[1], ev[3], inp_res)                                   # :340
    item["inp_bicubic_cnt"] = _interp(item["inp_cnt"], gt_res, "bicubic")                                   # :341
    item["inp_bicubic_stack"] = _interp(item["inp_stack"], gt_res, "bicubic")
    item["inp_near_cnt"] = _interp(item["inp_cnt"], gt_res, "nearest")
    item["inp_near_stack"] = _interp(item["inp_stack"], gt_res, "nearest")
    norm = create_normalized_events(ev, inp_res)                                                            # :345
    item["inp_scaled_cnt"] = _scaled_cnt(norm, gt_res)                                                      # :349
    item["inp_scaled_stack"] = oe.events_to_stack_no_polarity((norm[0] * np.float32(gt_res[1])).astype(np.float32),
                                                              (norm[1] * np.float32(gt_res[0])).astype(np.float32),
                                                              norm[2], norm[3].copy(), time_bins, gt_res)
    # create_unsupervised_data (:538-550): .long() truncation, renormalise, count on two grids, floor-divide by scale^2
    dx = np.trunc(norm[0] * np.float32(down[1])).astype(np.float32)
    dy = np.trunc(norm[1] * np.float32(down[0])).astype(np.float32)
    dn = create_normalized_events(np.stack([dx, dy, norm[2], norm[3]]), down)
    item["inp_down_cnt"] = np.floor(_scaled_cnt(dn, down) / np.float32(scale ** 2)).astype(np.float32)
    item["inp_down_scaled_cnt"] = np.floor(_scaled_cnt(dn, inp_res) / np.float32(scale ** 2)).astype(np.float32)
    item["gt_stack"] = oe.events_to_stack_no_polarity(gt[0], gt[1], gt[2], gt[3], time_bins, gt_res)        # :353
    item["gt_cnt"] = oe.events_to_channels(gt[0], gt[1], gt[3], gt_res)                                     # :354
    for k in ("inp_custom_cnt", "inp_custom_scaled_cnt", "inp_custom_down_cnt", "inp_custom_down_scaled_cnt", "gt_custom_cnt"):
        item[k] = np.zeros_like(item["inp_cnt"])                                                            # :364-365
    item["gt_img"] = np.zeros([1] + gt_res, np.float32)
    item["gt_inp_size_img"] = np.zeros([1] + inp_res, np.float32)
    item["frame"] = np.zeros([1] + gt_res, np.float32)
    return item
