"""``PointCloud`` container of the reference's IO layer (util/point_cloud.py:17-174): same fields,
methods and .npz / .ply formats, so objects round-trip with the reference's tools.  The O(N M) methods
(farthest-point sampling, nearest points) run on the GPU through the C ABI (``pcd_farthest_point_sample``,
``pcd_nearest_points``); like the rest of the package there is no CPU fallback for them."""
import contextlib
import os
import random
from dataclasses import dataclass
from typing import BinaryIO, Dict, List, Optional, Union

import numpy as np
import torch

from . import ops
from .ply_util import write_ply

COLORS = frozenset(["R", "G", "B", "A"])


def preprocess(data, channel):
    # colour channels are stored in [0, 1] and fed to the models in [0, 255] (util/point_cloud.py:12-15)
    return np.round(data * 255.0) if channel in COLORS else data


@contextlib.contextmanager
def _opened(f, mode):
    """A path is opened (and closed) here; an already open binary stream is used as is."""
    if isinstance(f, (str, os.PathLike)):
        with open(f, mode) as stream:
            yield stream
    else:
        yield f


def _device() -> torch.device:
    if not torch.cuda.is_available():
        from ._lib import PcdError
        raise PcdError("PointCloud sampling / neighbour queries run on the GPU: no CUDA device (there is no CPU fallback)")
    return torch.device("cuda")


@dataclass
class PointCloud:
    """coords: [N, 3] array; channels: name -> [N] array."""

    coords: np.ndarray
    channels: Dict[str, np.ndarray]

    # ---- .npz / .ply ----------------------------------------------------------------
    # .npz layout shared with the reference: one array "coords" plus one array per channel name.
    @classmethod
    def load(cls, f: Union[str, BinaryIO]) -> "PointCloud":
        with _opened(f, "rb") as stream:
            arrays = dict(np.load(stream))
        xyz = arrays.pop("coords")
        return cls(coords=xyz, channels=arrays)

    def save(self, f: Union[str, BinaryIO]) -> None:
        with _opened(f, "wb") as stream:
            np.savez(stream, coords=self.coords, **self.channels)

    def write_ply(self, raw_f: BinaryIO) -> None:
        colours = None
        if COLORS.issuperset("RGB") and all(c in self.channels for c in "RGB"):
            colours = np.column_stack([self.channels["R"], self.channels["G"], self.channels["B"]])
        write_ply(raw_f, coords=self.coords, rgb=colours)

    # ---- subsampling ----------------------------------------------------------------
    def random_sample(self, num_points: int, **subsample_kwargs) -> "PointCloud":
        if len(self.coords) <= num_points:
            return self
        indices = np.random.choice(len(self.coords), size=(num_points,), replace=False)
        return self.subsample(indices, **subsample_kwargs)

    def farthest_point_sample(self, num_points: int, init_idx: Optional[int] = None,
                              **subsample_kwargs) -> "PointCloud":
        """Evenly spread subset: greedy farthest-point sampling from ``init_idx`` (random if None)."""
        if len(self.coords) <= num_points:
            return self
        init_idx = random.randrange(len(self.coords)) if init_idx is None else init_idx
        pts = torch.from_numpy(np.ascontiguousarray(self.coords, dtype=np.float32)).to(_device())[None]
        indices = ops.farthest_point_sample(pts, num_points, init_idx)[0].cpu().numpy()
        return self.subsample(indices, **subsample_kwargs)

    def subsample(self, indices: np.ndarray, average_neighbors: bool = False) -> "PointCloud":
        if not average_neighbors:
            return PointCloud(coords=self.coords[indices],
                              channels={k: v[indices] for k, v in self.channels.items()})
        new_coords = self.coords[indices]
        owner = PointCloud(coords=new_coords, channels={}).nearest_points(self.coords)
        owner[indices] = np.arange(len(indices))  # every kept point owns itself (duplicates / rounding)
        counts = np.bincount(owner, minlength=len(indices)).astype(np.float64)
        new_channels = {}
        for k, v in self.channels.items():
            sums = np.bincount(owner, weights=v.astype(np.float64), minlength=len(indices))
            new_channels[k] = (sums / counts).astype(v.dtype)
        return PointCloud(coords=new_coords, channels=new_channels)

    def select_channels(self, channel_names: List[str]) -> np.ndarray:
        return np.stack([preprocess(self.channels[name], name) for name in channel_names], axis=-1)

    def nearest_points(self, points: np.ndarray, batch_size: int = 16384) -> np.ndarray:
        """For every row of ``points`` [M, 3] the index of the closest point of this cloud."""
        dev = _device()
        cloud = torch.from_numpy(np.ascontiguousarray(self.coords, dtype=np.float32)).to(dev)[None]
        out = []
        for i in range(0, len(points), batch_size):
            q = torch.from_numpy(np.ascontiguousarray(points[i:i + batch_size], dtype=np.float32)).to(dev)[None]
            out.append(ops.nearest_points(q, cloud, form=1)[0][0].cpu().numpy())
        return np.concatenate(out, axis=0)

    def combine(self, other: "PointCloud") -> "PointCloud":
        """Concatenate two clouds with the same channel set (low-res + upsampled stage outputs)."""
        if set(self.channels) != set(other.channels):
            raise AssertionError(f"channel sets differ: {sorted(self.channels)} vs {sorted(other.channels)}")
        merged = {name: np.concatenate((mine, other.channels[name])) for name, mine in self.channels.items()}
        return PointCloud(coords=np.vstack((self.coords, other.coords)), channels=merged)
