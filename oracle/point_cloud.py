"""CPU restatement (numpy) of the reference's point-cloud container, PLY writer and evaluation
metrics -- the data formats and callers either side of the sampler.
TEST INFRASTRUCTURE ONLY: imported by tests/ and oracle/make_golden_pc.py, never by the product.

Pinned against the unmodified reference through tests/golden/point_cloud.npz
(oracle/make_golden_pc.py runs util/point_cloud.py, util/ply_util.py and models/util.py of the
reference on seeded inputs)."""
import struct

import numpy as np
import torch


def ply_bytes(coords, rgb=None):
    """Binary little-endian PLY of a point cloud (reference util/ply_util.py:9-52, faces=None):
    header, then per vertex 3 float32 (+ 3 uint8 = round(rgb * 255.499))."""
    out = [b"ply\n", b"format binary_little_endian 1.0\n", b"element vertex %d\n" % len(coords),
           b"property float x\n", b"property float y\n", b"property float z\n"]
    if rgb is not None:
        out += [b"property uchar red\n", b"property uchar green\n", b"property uchar blue\n"]
    out.append(b"end_header\n")
    if rgb is not None:
        q = (np.asarray(rgb) * 255.499).round().astype(int)
        for c, k in zip(np.asarray(coords).tolist(), q.tolist()):
            out.append(struct.pack("<3f3B", *c, *k))
    else:
        for c in np.asarray(coords).tolist():
            out.append(struct.pack("<3f", *c))
    return b"".join(out)


def nearest_points(coords, points, batch_size=16384):
    """Index of the closest cloud point for every query point, |a|^2 + |b|^2 - 2 a.b form
    (reference util/point_cloud.py:148-165)."""
    norms = np.sum(coords ** 2, axis=-1)
    res = []
    for i in range(0, len(points), batch_size):
        b = points[i:i + batch_size]
        d = norms + np.sum(b ** 2, axis=-1)[:, None] - 2 * (b @ coords.T)
        res.append(np.argmin(d, axis=-1))
    return np.concatenate(res, axis=0)


def farthest_point_sample_indices(coords, num_points, init_idx):
    """Greedy farthest-point sampling (reference util/point_cloud.py:82-118): distances in the
    |a|^2 + |b|^2 - 2 a.b form, running minimum, first arg-max."""
    idx = np.zeros([num_points], dtype=np.int64)
    idx[0] = init_idx
    sq = np.sum(coords ** 2, axis=-1)

    def dists(i):
        return sq + sq[i] - 2 * (coords @ coords[i])

    cur = dists(init_idx)
    for k in range(1, num_points):
        j = np.argmax(cur)
        idx[k] = j
        cur = np.minimum(cur, dists(j))
    return idx


def subsample_average(coords, channels, indices):
    """PointCloud.subsample(average_neighbors=True) (reference util/point_cloud.py:120-142)."""
    new_coords = coords[indices]
    nb = nearest_points(new_coords, coords)
    nb[indices] = np.arange(len(indices))
    out = {}
    for k, v in channels.items():
        s = np.zeros_like(v[:len(indices)])
        c = np.zeros_like(v[:len(indices)])
        np.add.at(s, nb, v)
        np.add.at(c, nb, 1)
        out[k] = s / c
    return new_coords, out


def fscore(pred, gt, threshold, squared=False):
    """fscore_point_cloud_batch / _squared (reference models/util.py:195-262): pred [B,N,3],
    gt [B,M,3] -> (fscore, precision, recall), each [B]."""
    d2 = ((pred.unsqueeze(2) - gt.unsqueeze(1)) ** 2).sum(-1)
    a, b = d2.min(2).values, d2.min(1).values
    if not squared:
        a, b = torch.sqrt(a), torch.sqrt(b)
    p = (a < threshold).float().mean(1)
    r = (b < threshold).float().mean(1)
    return 2 * p * r / (p + r + 1e-8), p, r
