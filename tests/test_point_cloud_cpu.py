"""Point-cloud container / PLY writer / evaluation metrics: oracle vs goldens of the unmodified
reference, and the host-side (non-GPU) parts of the product."""
import io

import numpy as np
import torch

import pcd_b200 as P
from conftest import load_golden
from oracle import det
from oracle import point_cloud as OP
from oracle.make_golden_pc import inputs


def test_oracle_matches_reference_goldens():
    g = load_golden("point_cloud")
    coords, rgb, pred, gt = inputs()
    assert OP.ply_bytes(coords, rgb) == g["ply_rgb"].tobytes()
    assert OP.ply_bytes(coords) == g["ply_xyz"].tobytes()
    idx = OP.farthest_point_sample_indices(coords, 128, 17)
    assert np.array_equal(coords[idx], g["fps_coords"]) and np.array_equal(rgb[idx, 0], g["fps_R"])
    q = det.uniform((300, 3), 905, 0.5).numpy().astype(np.float32)
    assert np.array_equal(OP.nearest_points(coords, q), g["nearest"])
    sc, sch = OP.subsample_average(coords, {"G": rgb[:, 1]}, np.arange(0, 700, 7))
    assert np.array_equal(sc, g["sub_avg_coords"]) and np.allclose(sch["G"], g["sub_avg_G"], rtol=1e-6)
    for name, thr, sq in (("fscore", 0.03, False), ("fscore_sq", 1e-3, True)):
        got = torch.stack(OP.fscore(pred, gt, thr, sq)).numpy()
        assert np.allclose(got, g[name], atol=1e-7), name


def test_product_ply_and_npz_formats():
    """Byte-exact PLY output and .npz round trip of the product's PointCloud (host-side IO)."""
    g = load_golden("point_cloud")
    coords, rgb, _, _ = inputs()
    pc = P.PointCloud(coords=coords, channels={k: rgb[:, i] for i, k in enumerate("RGB")})
    f = io.BytesIO(); pc.write_ply(f)
    assert f.getvalue() == g["ply_rgb"].tobytes()
    f = io.BytesIO(); P.PointCloud(coords=coords, channels={}).write_ply(f)
    assert f.getvalue() == g["ply_xyz"].tobytes()
    f = io.BytesIO(); pc.save(f); f.seek(0)
    back = P.PointCloud.load(f)
    assert sorted(["coords"] + list(back.channels)) == list(g["npz_keys"])
    assert np.array_equal(back.coords, coords) and np.array_equal(back.channels["B"], rgb[:, 2])
    assert np.array_equal(pc.select_channels(["R", "B"]), g["select"])
    assert len(pc.combine(pc.subsample(np.arange(128))).coords) == int(g["combine_n"][0])
    # mesh form of the writer (faces): header + records
    f = io.BytesIO()
    P.ply_util.write_ply(f, coords[:4], faces=np.array([[0, 1, 2], [1, 2, 3]]))
    raw = f.getvalue()
    assert raw.startswith(b"ply\nformat binary_little_endian 1.0\nelement vertex 4\n") and b"element face 2\n" in raw
    assert len(raw) == raw.index(b"end_header\n") + 11 + 4 * 12 + 2 * 13


def test_point_cloud_gpu_methods_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        return
    coords, rgb, _, _ = inputs()
    pc = P.PointCloud(coords=coords, channels={})
    for fn in (lambda: pc.farthest_point_sample(16, init_idx=0), lambda: pc.nearest_points(coords[:4])):
        try:
            fn()
        except P._lib.PcdError:
            continue
        raise AssertionError("expected PcdError without a CUDA device")
