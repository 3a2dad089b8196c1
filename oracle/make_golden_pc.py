"""Golden vectors for the point-cloud container / PLY writer / evaluation metrics, produced by the
UNMODIFIED reference (util/point_cloud.py, util/ply_util.py, models/util.py) in the authoring container:

    python -m oracle.make_golden_pc        # -> tests/golden/point_cloud.npz
"""
import importlib
import io
import os

import numpy as np
import torch

from oracle import det
from oracle.ref_harness import load_reference

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def inputs():
    coords = det.uniform((700, 3), 901, 0.5).numpy().astype(np.float32)
    rgb = (det.uniform((700, 3), 902, 0.5).numpy().astype(np.float32) + 0.5).clip(0, 1)
    pred = det.uniform((3, 500, 3), 903, 0.5)
    gt = pred[:, :400] + 0.02 * det.normal((3, 400, 3), 904)
    return coords, rgb, pred, gt


def main():
    load_reference()
    pc_mod = importlib.import_module("point_e.util.point_cloud")
    util = importlib.import_module("point_e.models.util")
    coords, rgb, pred, gt = inputs()
    pc = pc_mod.PointCloud(coords=coords, channels={k: rgb[:, i] for i, k in enumerate("RGB")})
    out = {}
    f = io.BytesIO(); pc.write_ply(f); out["ply_rgb"] = np.frombuffer(f.getvalue(), dtype=np.uint8)
    f = io.BytesIO(); pc_mod.PointCloud(coords=coords, channels={}).write_ply(f)
    out["ply_xyz"] = np.frombuffer(f.getvalue(), dtype=np.uint8)
    fps = pc.farthest_point_sample(128, init_idx=17)
    out["fps_coords"] = fps.coords
    out["fps_R"] = fps.channels["R"]
    q = det.uniform((300, 3), 905, 0.5).numpy().astype(np.float32)
    out["nearest"] = pc.nearest_points(q)
    sub = pc.subsample(np.arange(0, 700, 7), average_neighbors=True)
    out["sub_avg_coords"], out["sub_avg_G"] = sub.coords, sub.channels["G"]
    out["select"] = pc.select_channels(["R", "B"])
    comb = pc.combine(fps)
    out["combine_n"] = np.array([len(comb.coords)])
    f = io.BytesIO(); pc.save(f); f.seek(0)
    back = pc_mod.PointCloud.load(f)
    out["npz_keys"] = np.array(sorted(["coords"] + list(back.channels.keys())))
    for name, thr, sq in (("fscore", 0.03, False), ("fscore_sq", 1e-3, True)):
        fn = util.fscore_point_cloud_batch_squared if sq else util.fscore_point_cloud_batch
        fs, p, r = fn(pred, gt, threshold=thr)
        out[name] = torch.stack([fs, p, r]).numpy()
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "point_cloud.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
