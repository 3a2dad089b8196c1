"""GPU parity of the point-cloud utilities (farthest-point sampling, nearest points, F-score) through the
C ABI against the oracle and the reference goldens."""
import numpy as np
import pytest
import torch

import pcd_b200 as P
from conftest import load_golden
from gpu_util import DEV
from oracle import det
from oracle import point_cloud as OP
from oracle.make_golden_pc import inputs

pytestmark = pytest.mark.gpu
ops = P.ops


def test_point_cloud_methods_match_reference_goldens():
    g = load_golden("point_cloud")
    coords, rgb, _, _ = inputs()
    pc = P.PointCloud(coords=coords, channels={k: rgb[:, i] for i, k in enumerate("RGB")})
    fps = pc.farthest_point_sample(128, init_idx=17)
    assert np.array_equal(fps.coords, g["fps_coords"]) and np.array_equal(fps.channels["R"], g["fps_R"])
    q = det.uniform((300, 3), 905, 0.5).numpy().astype(np.float32)
    assert np.array_equal(pc.nearest_points(q), g["nearest"])
    sub = pc.subsample(np.arange(0, 700, 7), average_neighbors=True)
    assert np.array_equal(sub.coords, g["sub_avg_coords"])
    assert np.allclose(sub.channels["G"], g["sub_avg_G"], rtol=1e-6)


@pytest.mark.parametrize("B,N,n", [(1, 700, 128), (3, 4096, 1024), (2, 8192, 1024), (2, 33, 33),
                                   (2, 16384, 512)])   # MVP-size complete clouds: distances in a global workspace
def test_farthest_point_sample_indices(B, N, n):
    """Index parity (bit-exact) with the oracle's numpy restatement at the evaluation sizes."""
    pts = det.uniform((B, N, 3), 930 + N, 0.5)
    init = torch.tensor([(7 * b + 3) % N for b in range(B)], dtype=torch.int32)
    got = ops.farthest_point_sample(pts.to(DEV), n, init.to(DEV)).cpu().numpy()
    for b in range(B):
        want = OP.farthest_point_sample_indices(pts[b].numpy(), n, int(init[b]))
        if not np.array_equal(got[b], want):
            # numpy's float32 matmul (BLAS sgemv) may fuse / reorder the 3-term dot product, so the reference itself
            # is only defined up to fp32 rounding of |a|^2 + |b|^2 - 2 a.b (terms of magnitude ~1 here): accept a
            # different winner only where the two candidates' exact distances agree to a few ulps of those terms
            k = int(np.argmax(got[b] != want))
            c = pts[b].numpy().astype(np.float64)
            sel = c[want[:k]]
            d = lambda i: ((sel - c[i]) ** 2).sum(1).min()
            assert abs(d(got[b][k]) - d(want[k])) <= 1e-6, (b, k, got[b][k], want[k], d(got[b][k]), d(want[k]))
            assert np.array_equal(got[b][:k], want[:k])
    assert len(set(got[0].tolist())) == n  # no point twice


def test_farthest_point_sample_rejects_and_clamps_bad_start_indices():
    pts = det.uniform((2, 100, 3), 951, 0.5).to(DEV)
    with pytest.raises(IndexError):
        ops.farthest_point_sample(pts, 10, 100)
    with pytest.raises(IndexError):
        ops.farthest_point_sample(pts, 10, -1)
    # a device tensor cannot be checked without a sync: the kernel clamps it into the cloud
    got = ops.farthest_point_sample(pts, 10, torch.tensor([1000, -5], dtype=torch.int32, device=DEV)).cpu()
    assert got[0, 0] == 99 and got[1, 0] == 0
    want = ops.farthest_point_sample(pts, 10, torch.tensor([99, 0], dtype=torch.int32, device=DEV)).cpu()
    assert torch.equal(got, want)


def test_fscore_and_nearest_points():
    g = load_golden("point_cloud")
    _, _, pred, gt = inputs()
    for name, thr, sq in (("fscore", 0.03, False), ("fscore_sq", 1e-3, True)):
        got = torch.stack(ops.fscore_point_cloud_batch(pred.to(DEV), gt.to(DEV), thr, squared=sq)).cpu().numpy()
        assert np.allclose(got, g[name], atol=1e-6), (name, got, g[name])
    # evaluation size: 8192 predicted vs 8192 target points (the reference materialises [B, N, M, 3])
    a, b = det.uniform((2, 8192, 3), 940, 0.5), det.uniform((2, 8192, 3), 941, 0.5)
    got = torch.stack(ops.fscore_point_cloud_batch(a.to(DEV), b.to(DEV), 0.01)).cpu()
    want = torch.stack(OP.fscore(a[:1, :2048], b[:1], 0.01))  # oracle on a slice (memory)
    idx, d2 = ops.nearest_points(a.to(DEV), b.to(DEV), form=0)
    dd = ((a[0, :512, None] - b[0, None]) ** 2).sum(-1)
    assert torch.equal(idx[0, :512].cpu(), dd.argmin(1)) and torch.allclose(d2[0, :512].cpu(), dd.min(1).values, atol=1e-7)
    assert got.shape == (3, 2) and bool((got >= 0).all()) and bool((got <= 1).all()) and want.shape == (3, 1)
