"""Generate tests/golden/items_golden.npz from the REFERENCE's own H5Dataset.__getitem__ (run in the build container only).

Needs /root/reference and oracle/_ref (python oracle/build_ref.py).  The reference class is imported unmodified; only its
HDF5 access is replaced: a subclass skips __init__ (h5py is absent and no data file ships) and returns seeded synthetic events
from get_events / get_gt_events, so the tensor factory of dataloader/h5dataset.py:276-406 -- count / stack encodings, bicubic
and nearest up-samplings, the normalise-and-lift encodings and create_unsupervised_data -- runs exactly as shipped.
Modules the file imports for plotting / image I/O only (h5py, cv2, matplotlib, the visualisation helpers) are stubbed.
"""
import os
import sys
import types
from unittest import mock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(1, "/root/reference")
for name in ("h5py", "cv2", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1",
             "myutils", "myutils.vis_events", "myutils.vis_events.visualization"):
    m = types.ModuleType(name)
    m.__dict__.setdefault("__path__", [])
    if name == "matplotlib.pyplot":
        m.style = mock.MagicMock()
    if name == "mpl_toolkits.axes_grid1":
        m.ImageGrid = mock.MagicMock()
    sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from dataloader.h5dataset import H5Dataset  # noqa: E402

CASES = [  # (inp H, W), scale, time_bins, n_inp events, seed
    ((32, 48), 2, 1, 3000, 1),
    ((45, 80), 4, 5, 2500, 2),
    ((30, 42), 2, 3, 500, 3),           # odd down-scaled size (15 x 21), few events
    ((40, 40), 4, 1, 3000, 4),
]


class Synthetic(H5Dataset):
    def __init__(self, inp_res, scale, time_bins, n, seed):                 # no HDF5: attributes the factory reads
        self.config = {"data_augment": {"enabled": False}}
        self.need_gt_events, self.need_gt_frame, self.data_mode = True, False, "events"
        self.add_noise, self.custom_resolution = {"enabled": False}, None
        self.scale, self.time_bins = scale, time_bins
        self.inp_sensor_resolution = list(inp_res)
        self.gt_sensor_resolution = [round(i * scale) for i in inp_res]
        self.inp_down_sensor_resolution = [round(i / scale) for i in inp_res]
        rng = np.random.default_rng(seed)

        def ev(n, H, W):
            ts = np.sort(rng.random(n)) * 0.05 + 3.0                         # float64 seconds, as stored on disk
            return np.stack([rng.integers(0, W, n).astype(np.float64), rng.integers(0, H, n).astype(np.float64), ts,
                             rng.choice([-1.0, 1.0], n)])
        self.inp = ev(n, *self.inp_sensor_resolution)
        self.gt = ev(n * scale * scale, *self.gt_sensor_resolution)

    def get_event_indices(self, index): return 0, self.inp.shape[1]
    def get_gt_event_indices(self, index): return 0, self.gt.shape[1]
    def get_events(self, i0, i1): return self.inp[:, i0:i1].copy()
    def get_gt_events(self, i0, i1): return self.gt[:, i0:i1].copy()


def main():
    out = {"n_cases": len(CASES)}
    for i, (res, scale, tb, n, seed) in enumerate(CASES):
        ds = Synthetic(res, scale, tb, n, seed)
        item = ds.__getitem__(0, seed=0)
        out[f"c{i}_cfg"] = np.array([res[0], res[1], scale, tb])
        out[f"c{i}_inp_events"], out[f"c{i}_gt_events"] = ds.inp, ds.gt
        for k, v in item.items():
            out[f"c{i}_{k}"] = v.numpy()
    np.savez_compressed(os.path.join(HERE, "items_golden.npz"), **out)
    print("wrote items_golden.npz:", {k: tuple(v.shape) for k, v in item.items()})


if __name__ == "__main__":
    main()
