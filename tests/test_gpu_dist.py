"""Multi-GPU path through the product API on real devices (NCCL): batch-sharded sampling must reproduce the
single-GPU result bit for bit when every rank draws the FULL-batch noise and keeps its own slice (SURVEY 7.2 RNG
rule).  Needs two GPUs; the CPU suite covers the same host logic with gloo (tests/test_dist_cpu.py)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_path):
    import torch.distributed as dist

    import pcd_b200 as P
    from gpu_util import build_model
    from oracle import cases

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import gpu_util
        gpu_util.DEV = dev
        case = "small_imagevec_guided"
        sc = cases.SAMPLER_CASES[case]
        batch = 5  # ragged over two ranks: shards of 3 and 2 clouds
        model, cfg, _ = build_model(sc["model"], torch.bfloat16)
        model = model.to(dev)
        diffusion = P.diffusion_from_config(P.DIFFUSION_CONFIGS["base40M"])
        from oracle import det
        e = det.normal((batch, 768), 31)
        kw_full = dict(embeddings=(e / e.norm(dim=1, keepdim=True)).to(dev))

        def make(lo, hi, graph):
            noise = cases.DetNoise(sc["noise_seed"])
            # every rank draws the full-batch noise and keeps its slice: a cloud's trajectory does not depend on
            # which rank samples it
            return P.PointCloudSampler(dev, [model], [diffusion], [cfg["n_ctx"]], ["R", "G", "B"], guidance_scale=[3.0],
                                       use_karras=[True], karras_steps=[8], sigma_min=[1e-3], sigma_max=[120.0], s_churn=[3.0],
                                       use_cuda_graph=graph,
                                       noise_fn=lambda shp: noise((batch,) + tuple(shp[1:]))[lo:hi].to(dev))
        lo, hi = P.dist.shard_bounds(batch, world, rank)
        for graph in (False, True):
            sharded = P.dist.sample_sharded(make(lo, hi, graph).sample_batch, batch, kw_full)
            assert sharded.shape == (batch, 6, cfg["n_ctx"])
            if rank == 0:
                whole = make(0, batch, graph).sample_batch(batch, kw_full)
                assert torch.equal(sharded, whole), float((sharded - whole).abs().max())
        if rank == 0:
            open(out_path, "w").write("ok")
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_sharded_sampling_equals_single_gpu_bitwise(tmp_path):
    import torch.multiprocessing as mp
    out = tmp_path / "result"
    mp.spawn(_worker, args=(2, 29500 + os.getpid() % 2000, str(out)), nprocs=2, join=True)
    assert out.read_text() == "ok"
