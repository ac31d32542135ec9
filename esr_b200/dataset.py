"""Host-side mirrors of the dataset glue on the hot path (no HDF5 I/O here: h5py is absent and no data ships; the
reader itself is SURVEY.md 8f "next").  Same names, argument meaning and results as the reference methods, but the
tensors live on the GPU and the encodings are the sm_100a kernels of esr_b200.encodings.

  event_formatting(events)                                   dataloader/base_dataset.py:26-33
  create_normalized_events(events, sensor_resolution)        dataloader/h5dataset.py:508-518
  create_scaled_encoding(norm_events, sensor_resolution, mode, time_bins)   dataloader/h5dataset.py:520-536
  create_cnt_encoding(events, sensor_resolution)             dataloader/h5dataset.py:611-619
  sliding_windows(frames, num_frame)                         dataloader/h5dataloader.py:210-246 (custom_collate / concat_dict)
"""
import numpy as np
import torch

from . import encodings


def _dev():
    return torch.device("cuda", torch.cuda.current_device())


def event_formatting(events, device=None):
    """events: numpy [4, n] (x, y, t, p) -> float32 tensor [4, n] with t normalised to [0, 1)."""
    device = device or _dev()
    xs = torch.from_numpy(np.asarray(events[0]).astype(np.float32)).to(device)
    ys = torch.from_numpy(np.asarray(events[1]).astype(np.float32)).to(device)
    ts = torch.from_numpy(np.asarray(events[2]).astype(np.float32)).to(device)
    ps = torch.from_numpy(np.asarray(events[3]).astype(np.float32)).to(device)
    ts = (ts - ts[0]) / (ts[-1] - ts[0] + 1e-6)      # tensor / tensor: IEEE division on the device too
    return torch.stack([xs, ys, ts, ps])


def create_normalized_events(events, sensor_resolution):
    xs, ys, ts, ps = events[0], events[1], events[2], events[3]
    # divisors as device tensors: torch's CUDA `tensor / python_scalar` multiplies by the reciprocal, which is not the
    # IEEE division the reference's CPU tensors get (differs for non-power-of-two sensor sizes)
    dw = torch.tensor(float(sensor_resolution[1]), dtype=torch.float32, device=xs.device)
    dh = torch.tensor(float(sensor_resolution[0]), dtype=torch.float32, device=xs.device)
    xs, ys = xs / dw, ys / dh
    return torch.stack([xs, ys, ts, ps]).float()


def create_scaled_encoding(normalized_events, sensor_resolution, mode, time_bins=1):
    xs, ys, ts, ps = normalized_events[0], normalized_events[1], normalized_events[2], normalized_events[3]
    if mode == 'cnt':
        return encodings.events_to_channels(xs * sensor_resolution[1], ys * sensor_resolution[0], ps,
                                            sensor_size=sensor_resolution)
    if mode == 'stack':
        return encodings.events_to_stack_no_polarity(xs * sensor_resolution[1], ys * sensor_resolution[0], ts, ps,
                                                     B=time_bins, sensor_size=sensor_resolution)
    if mode == 'events':
        return torch.stack([(xs * sensor_resolution[1]).long(), (ys * sensor_resolution[0]).long(), ts, ps], dim=0)
    raise Exception(f'mode: {mode} is NOT supported!')


def create_cnt_encoding(events, sensor_resolution):
    xs, ys, ts, ps = events[0], events[1], events[2], events[3]
    return encodings.events_to_channels(xs, ys, ps, sensor_size=sensor_resolution)


def sliding_windows(frames, num_frame=3):
    """frames: [B, L, ...] -> list of L-num_frame+1 windows [B, num_frame, ...] sliding by one frame, the layout
    HDF5DataLoaderSequence.custom_collate builds.  (esr_b200.DeepRecurrNet.forward_sequence takes [B, L, ...] directly and
    never materialises these copies.)"""
    L = frames.shape[1]
    return [frames[:, w:w + num_frame].contiguous() for w in range(L - num_frame + 1)]
