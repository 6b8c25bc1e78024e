"""Dataset container for the features the hot path produces (SURVEY section 8f "next" #2).

The reference appends float64 arrays to resizable HDF5 datasets (preprocessing/utils/io_manager.py:39-76:
``pianoroll``, ``onoff`` (N,860,128) and ``spec_{style}`` (N,1025,860)) and reads them back whole in
model/train.py:45-104.  h5py is not available here; the same logical layout is kept as a directory of ``.npy``
shards plus ``index.json`` -- one shard per append, so ranks can write disjoint song ranges independently.
``ShardManager`` mirrors ``h5pyManager`` (same method names / arguments), ``ShardDataset`` mirrors ``Dataseth5py``
(same item: X (256,860) = roll||onoff transposed, X_cond = a random chunk's spectrogram of a random style, y = the
matching spectrogram).
"""
import json
import os
import random

import numpy as np
import torch


class ShardManager():
    """Drop-in for io_manager.h5pyManager: indexes line up so that pianoroll[i] / onoff[i] / spec_{style}[i] match.

    ``dtype='native'`` stores rolls as int8 and spectrograms as float32 (bit-identical to the reference after the
    FloatTensor conversion of train.py:93-99, 8x / 2x smaller); ``dtype='float64'`` stores what the reference stores.
    """

    def __init__(self, root, dtype=None, mode="a"):
        """mode 'w' truncates like the reference's ``h5py.File(outfile, 'w')`` (io_manager.py:41), 'a' appends to an
        existing container, 'r' opens an existing one read-only.  ``dtype`` None = whatever the container holds
        ('native' for a new one); an explicit dtype that contradicts an existing container raises."""
        if mode not in ("w", "a", "r"):
            raise ValueError(f"mode={mode!r} (w | a | r)")
        self.root, self.mode = root, mode
        self.index_path = os.path.join(root, "index.json")
        if mode == "r" and not os.path.exists(self.index_path):
            raise FileNotFoundError(self.index_path)
        os.makedirs(root, exist_ok=True)
        if mode == "w" and os.path.exists(self.index_path):
            with open(self.index_path) as f:
                old = json.load(f)
            for entries in old.get("keys", {}).values():      # drop the shards the old index owns, nothing else
                for e in entries:
                    try:
                        os.remove(os.path.join(root, e["file"]))
                    except FileNotFoundError:
                        pass
            os.remove(self.index_path)
        if os.path.exists(self.index_path):
            with open(self.index_path) as f:
                self.index = json.load(f)
            if dtype is not None and dtype != self.index["dtype"]:
                raise ValueError(f"container at {root} holds dtype={self.index['dtype']!r}, asked for {dtype!r}")
        else:
            self.index = {"dtype": dtype or "native", "keys": {}}
        self.dtype = self.index["dtype"]
        if self.dtype not in ("native", "float64"):
            raise ValueError(f"dtype={self.dtype!r} (native | float64)")

    def _append(self, key, arr, native):
        if self.mode == "r":
            raise IOError("container opened read-only")
        arr = np.asarray(arr)
        arr = arr.astype(np.float64 if self.dtype == "float64" else native, copy=False)
        entries = self.index["keys"].setdefault(key, [])
        if entries and list(arr.shape[1:]) != entries[0]["shape"][1:]:
            raise ValueError(f"{key}: chunk shape {arr.shape[1:]} does not match earlier shards {entries[0]['shape'][1:]}")
        os.makedirs(os.path.join(self.root, key), exist_ok=True)
        name = os.path.join(key, f"{len(entries):05d}.npy")
        np.save(os.path.join(self.root, name), np.ascontiguousarray(arr))
        entries.append({"file": name, "shape": list(arr.shape)})
        with open(self.index_path, "w") as f:
            json.dump(self.index, f, indent=1)

    def write_pianoroll(self, pianoroll_list, onoff_list):
        """io_manager.py:46-62."""
        self._append("pianoroll", pianoroll_list, np.int8)
        self._append("onoff", onoff_list, np.int8)

    def write_spectrum(self, spec_list, style):
        """io_manager.py:64-76."""
        self._append(f"spec_{style}", spec_list, np.float32)

    def keys(self):
        return list(self.index["keys"].keys())

    def n_rows(self, key):
        return sum(e["shape"][0] for e in self.index["keys"][key])

    def read(self, key, n_read=None):
        parts, have = [], 0
        for e in self.index["keys"][key]:
            if n_read is not None and have >= n_read:
                break
            a = np.load(os.path.join(self.root, e["file"]), mmap_mode="r")
            parts.append(a if n_read is None else a[:max(0, n_read - have)])
            have += parts[-1].shape[0]
        return np.concatenate(parts, axis=0) if parts else np.zeros((0,))


class ShardDataset(torch.utils.data.Dataset):
    """Drop-in for train.py::Dataseth5py (train.py:45-104) over a ShardManager directory."""

    def __init__(self, in_dir, seed=42, n_read=None, device=None, resident=None):
        """``resident`` (default: True when ``device`` is a CUDA device): upload every array ONCE and serve items
        straight from GPU memory -- int8 rolls and float32 spectrograms stay in HBM (a 100-chunk song is 22 MB of rolls
        and 353 MB per style), an item is two gathers and a transposed cast on the device, nothing crosses PCIe per
        item (the reference builds ``torch.cuda.FloatTensor(item)`` from host arrays on every access, train.py:93-99)."""
        super(ShardDataset, self).__init__()
        self.manager = ShardManager(in_dir, mode="r")
        self.styles = [name for name in self.manager.keys() if 'spec_' in name]
        self.pianoroll = self.manager.read('pianoroll', n_read)
        self.onoff = self.manager.read('onoff', n_read)
        self.specs = {style: self.manager.read(style, n_read) for style in self.styles}
        self.n_data = self.pianoroll.shape[0]
        self.device = None if device is None else torch.device(device)
        self.resident = (self.device is not None and self.device.type == "cuda") if resident is None else bool(resident)
        if self.resident:
            if self.device is None or self.device.type != "cuda":
                raise ValueError("resident=True needs a CUDA device")
            up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(self.device)
            self.pianoroll, self.onoff = up(self.pianoroll), up(self.onoff)
            self.specs = {k: up(v) for k, v in self.specs.items()}
        random.seed(seed)

    def __getitem__(self, index):
        if self.resident:
            # same RNG call order as the host path below (train.py:76-101): style first, then the conditioning index
            style = random.choice(self.styles)
            rand_index = random.randint(0, self.n_data - 1)
            X = torch.cat((self.pianoroll[index], self.onoff[index]), dim=-1).t().to(torch.float32).contiguous()
            return X, self.specs[style][rand_index].to(torch.float32), self.specs[style][index].to(torch.float32)
        pianoroll = np.concatenate((self.pianoroll[index], self.onoff[index]), axis=-1)
        pianoroll = np.transpose(pianoroll, (1, 0))
        style = random.choice(self.styles)
        spec = self.specs[style][index]
        rand_index = random.randint(0, self.n_data - 1)
        spec_rand = self.specs[style][rand_index]
        to = dict(dtype=torch.float32, device=self.device)
        X = torch.as_tensor(np.ascontiguousarray(pianoroll), **to)
        X_cond = torch.as_tensor(np.ascontiguousarray(spec_rand), **to)
        y = torch.as_tensor(np.ascontiguousarray(spec), **to)
        return X, X_cond, y

    def __len__(self):
        return self.n_data
