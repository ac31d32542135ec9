"""Generate tests/golden/index_golden.npz from the REFERENCE's own H5Dataset indexing code (run in the build container only).

The reference class (dataloader/h5dataset.py) is imported unmodified; `h5py.File` is replaced by an in-memory object with the
group / dataset / attribute access pattern the class uses (h5py is absent and no data file ships), filled with seeded synthetic
event columns in the on-disk dtypes of generate_dataset/tools/event_packagers.py:129-132.  The window tables come from the
reference's compute_k_indices / compute_timeblock_indices / compute_frame_indices + get_gt_event_indices_num + find_ts_index
(h5dataset.py:196-270, 451-475) and its Python bisection (base_dataset.py:78-91), duplicate timestamps included; the first
frames' formatted events come from get_events + event_formatting (h5dataset.py:492-498, base_dataset.py:26-33).
"""
import os
import sys
import types
from unittest import mock

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
sys.path.insert(1, "/root/reference")


class _Node:
    def __init__(self, value=None, attrs=None):
        self.value, self.attrs, self.children = value, attrs or {}, {}

    def __getitem__(self, key):
        if isinstance(key, str):
            node = self
            for part in key.split("/"):
                node = node.children[part]
            return node
        return self.value[key]

    def __iter__(self):
        return iter(sorted(self.children))

    def __len__(self):
        return len(self.children) if self.value is None else len(self.value)

    def keys(self):
        return self.children.keys()

    @property
    def shape(self):
        return self.value.shape


_FILES = {}


def _open(path, mode="r"):
    return _FILES[path]


h5 = types.ModuleType("h5py")
h5.File = _open
sys.modules["h5py"] = h5
for name in ("cv2", "matplotlib", "matplotlib.pyplot", "mpl_toolkits", "mpl_toolkits.axes_grid1", "myutils", "myutils.vis_events",
             "myutils.vis_events.visualization"):
    m = types.ModuleType(name)
    m.__dict__.setdefault("__path__", [])
    if name == "matplotlib.pyplot":
        m.style = mock.MagicMock()
    if name == "mpl_toolkits.axes_grid1":
        m.ImageGrid = mock.MagicMock()
    sys.modules[name] = m
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]

from dataloader.h5dataset import H5Dataset  # noqa: E402


def synth_columns(seed, sensor, n_ori, scales, dup_every=7):
    """event columns per scale prefix: finer scales hold 4x the events per octave (as the simulator produces), timestamps sorted
    float64 seconds with repeated values (the exact-hit branch of the reference's bisection)."""
    rng = np.random.default_rng(seed)
    cols = {}
    for prex, div in scales.items():
        n = max(16, n_ori // (div * div))
        H, W = round(sensor[0] / div), round(sensor[1] / div)
        ts = np.sort(rng.random(n)) * 2.0 + 10.0
        ts[dup_every::dup_every] = ts[dup_every - 1:-1:dup_every][:len(ts[dup_every::dup_every])]   # duplicates
        ts = np.sort(ts)
        cols[prex] = {"xs": rng.integers(0, W, n).astype(np.int16), "ys": rng.integers(0, H, n).astype(np.int16),
                      "ts": ts.astype(np.float64), "ps": rng.choice([-1.0, 1.0], n).astype(np.float64)}
    return cols


def fake_file(path, cols, sensor, image_ts):
    root = _Node(attrs={"sensor_resolution": np.array(sensor)})
    for prex, c in cols.items():
        g = _Node()
        for k, v in c.items():
            g.children[k] = _Node(v)
        root.children[f"{prex}_events"] = g
    imgs = _Node()
    for i, t in enumerate(image_ts):
        imgs.children["image{:09d}".format(i)] = _Node(np.zeros((2, 2), np.uint8), {"timestamp": float(t)})
    root.children["ori_images"] = imgs
    _FILES[path] = root


CASES = [
    # name, seed, sensor, n_ori, config
    ("k2", 1, (64, 96), 48000, dict(scale=2, ori_scale="down4", need_gt_events=True, mode="events", window=512, sliding_window=128)),
    ("k4", 2, (64, 64), 64000, dict(scale=4, ori_scale="down4", need_gt_events=True, mode="events", window=300, sliding_window=0)),
    ("t2", 3, (48, 80), 40000, dict(scale=2, ori_scale="down2", need_gt_events=True, mode="time", window=0.05, sliding_window=0.01)),
    ("f2", 4, (64, 96), 30000, dict(scale=2, ori_scale="down4", need_gt_events=True, mode="frame", window=0, sliding_window=0, need_gt_frame=True)),
    ("k2n", 5, (64, 96), 20000, dict(scale=2, ori_scale="down8", need_gt_events=False, mode="events", window=100, sliding_window=50, dataset_length=40)),
]


def main():
    out = {"names": np.array([c[0] for c in CASES])}
    for name, seed, sensor, n_ori, cfg in CASES:
        cols = synth_columns(seed, sensor, n_ori, {"ori": 1, "down2": 2, "down4": 4, "down8": 8})
        rng = np.random.default_rng(seed + 100)
        inp_ts = cols[cfg["ori_scale"]]["ts"]
        image_ts = np.sort(rng.uniform(inp_ts[0], inp_ts[-1], 24))
        image_ts[5] = inp_ts[len(inp_ts) // 3]                               # an exact hit
        path = f"/fake/{name}.h5"
        fake_file(path, cols, sensor, image_ts)
        config = dict(time_bins=1, data_augment={"enabled": False}, **cfg)
        ds = H5Dataset(path, config)
        out[f"{name}_sensor"] = np.array(sensor)
        out[f"{name}_image_ts"] = image_ts
        for prex, c in cols.items():
            if prex in (ds.inp_prex, ds.gt_prex):                            # only the streams this configuration reads
                for k, v in c.items():
                    out[f"{name}_{prex}_{k}"] = v
        out[f"{name}_cfg"] = np.array([repr(cfg)])
        out[f"{name}_length"] = np.array([ds.length])
        out[f"{name}_event_indices"] = np.array(ds.event_indices, dtype=np.int64)
        if cfg["need_gt_events"]:
            out[f"{name}_gt_event_indices"] = np.array(ds.gt_event_indices, dtype=np.int64)
        out[f"{name}_res"] = np.array([ds.inp_sensor_resolution, ds.gt_sensor_resolution])
        for fr in (0, ds.length // 2):
            i0, i1 = ds.get_event_indices(fr)
            out[f"{name}_events_{fr}"] = ds.event_formatting(ds.get_events(i0, i1)).numpy()
        print(name, "length", ds.length, "first", ds.event_indices[:2], ds.gt_event_indices[:2] if cfg["need_gt_events"] else None)
    np.savez_compressed(os.path.join(HERE, "index_golden.npz"), **out)
    print("wrote index_golden.npz", os.path.getsize(os.path.join(HERE, "index_golden.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
