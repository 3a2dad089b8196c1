"""Binary PLY writer for point clouds and meshes (mirrors reference util/ply_util.py:9-52: same header
lines, little-endian float32 xyz, colours quantised as round(rgb * 255.499), faces as uchar count +
3 uint32)."""
from typing import BinaryIO, Optional

import numpy as np


def write_ply(raw_f: BinaryIO, coords: np.ndarray, rgb: Optional[np.ndarray] = None,
              faces: Optional[np.ndarray] = None) -> None:
    coords = np.asarray(coords)
    n = len(coords)
    props = ["property float x", "property float y", "property float z"]
    if rgb is not None:
        props += ["property uchar red", "property uchar green", "property uchar blue"]
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {n}", *props]
    if faces is not None:
        header += [f"element face {len(faces)}", "property list uchar int vertex_index"]
    header.append("end_header")
    raw_f.write(("\n".join(header) + "\n").encode("ascii"))

    # vertex records assembled as one structured array instead of per-vertex struct.pack calls
    if rgb is not None:
        rec = np.empty(n, dtype=[("xyz", "<f4", (3,)), ("rgb", "u1", (3,))])
        rec["rgb"] = (np.asarray(rgb) * 255.499).round().astype(np.int64).astype(np.uint8)
    else:
        rec = np.empty(n, dtype=[("xyz", "<f4", (3,))])
    rec["xyz"] = coords.astype(np.float32)
    raw_f.write(rec.tobytes())
    if faces is not None:
        faces = np.asarray(faces)
        frec = np.empty(len(faces), dtype=[("n", "u1"), ("v", "<u4", (3,))])
        frec["n"] = faces.shape[1] if faces.ndim == 2 else 3
        frec["v"] = faces.astype(np.uint32)
        raw_f.write(frec.tobytes())
