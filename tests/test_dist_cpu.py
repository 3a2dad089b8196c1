"""world_size-2 gloo test of the batch-sharding + final all-gather logic (host side only)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, batch, q):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import pcd_b200 as P
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    emb = torch.arange(batch * 4, dtype=torch.float32).reshape(batch, 4)

    def fake_sampler(b, kw):  # stands in for PointCloudSampler.sample_batch on the rank's shard
        assert kw["embeddings"].shape[0] == b and kw["flag"] == "x"
        return kw["embeddings"].sum(1)[:, None, None].expand(b, 3, 5).contiguous() + 0.0

    out = P.dist.sample_sharded(fake_sampler, batch, dict(embeddings=emb, flag="x"))
    q.put((rank, out.numpy().copy()))
    dist.destroy_process_group()


@pytest.mark.parametrize("batch", [4, 5, 1])
def test_sharded_sampling_gloo(batch):
    import pcd_b200 as P
    assert [P.dist.shard_bounds(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [P.dist.shard_bounds(1, 2, r) for r in range(2)] == [(0, 1), (1, 1)]
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, q)) for r in range(world)]
    [p.start() for p in procs]
    res = dict(q.get(timeout=120) for _ in range(world))
    [p.join(60) for p in procs]
    want = torch.arange(batch * 4, dtype=torch.float32).reshape(batch, 4).sum(1)[:, None, None].expand(batch, 3, 5)
    for r in range(world):
        assert torch.equal(torch.from_numpy(res[r]), want)
