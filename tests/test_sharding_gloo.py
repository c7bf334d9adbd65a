"""The N > 1 path on CPU: two gloo ranks shard the sample loop and reduce their partial radiance sums."""
import importlib
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp_total, out):
    for p in (ROOT, os.path.join(ROOT, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rt = importlib.import_module("raytracing-1w_b200")
    shard = importlib.import_module("raytracing-1w_b200.shard")
    import oracle_binding  # the CPU checker stands in for the per-rank renderer (no GPU here)

    hs = rt.api.HostScene("cornel_box", seed=1)
    osc = oracle_binding.OracleScene(hs.desc)
    s0, s1 = shard.sample_range(rank, world, spp_total)
    img, _, st = osc.render(hs.camera(), hs.params(width=24, spp=s1, sample_begin=s0), threads=1)
    part = torch.from_numpy(img.astype(np.float32))
    mine = part.clone()
    shard.reduce_radiance(part, dst=0)
    gathered = [torch.zeros_like(mine) for _ in range(world)] if rank == 0 else None
    dist.gather(mine, gathered, dst=0)
    if rank == 0:
        total = torch.stack(gathered).sum(0)
        assert torch.allclose(part, total, rtol=1e-6, atol=1e-6)
        mean = shard.resolve(part, spp_total)
        torch.save({"mean": mean, "paths": st.paths}, out)
    dist.destroy_process_group()


def test_two_ranks_shard_and_reduce(tmp_path, built):
    out = str(tmp_path / "r0.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, 10, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["mean"].shape == (24, 24, 3) and torch.isfinite(res["mean"]).all()
    assert res["paths"] == 24 * 24 * 5  # rank 0 rendered half of the 10 samples
    assert 0.02 < float(res["mean"].mean()) < 0.4  # a plausible Cornell mean radiance


def test_sample_ranges_partition():
    shard = importlib.import_module("raytracing-1w_b200.shard")
    for world in (1, 2, 3, 4, 8):
        for spp in (1, 7, 100, 4096):
            ranges = [shard.sample_range(r, world, spp) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == spp
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    assert shard.weak_sample_range(3, 100) == (300, 400)
