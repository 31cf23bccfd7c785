"""The N>1 path on the CPU: world_size-2 gloo process group, samples split across ranks, accumulators summed
with one reduction, resolve on rank 0.  The per-rank accumulators are produced by the oracle (the GPU kernel
cannot run here); what is under test is the product's host-side plumbing in grayshift_b200/dist.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, spp, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import oracle_ffi as orc
    from grayshift_b200 import dist as gdist

    begin, end = gdist.sample_range(spp, rank, world)
    ora = orc.OracleScene.build("c3", width=24, spp=end - begin)
    rgb, counts, _, _ = ora.render(seed=9, rng=orc.OracleScene.PHILOX, threads=2, sample_begin=begin)
    # the oracle returns per-pixel means; turn them back into (sum r, sum g, sum b, count) accumulators
    acc = np.concatenate([rgb.reshape(-1, 3) * counts.reshape(-1, 1), counts.reshape(-1, 1).astype(np.float64)], 1)
    accum = torch.from_numpy(acc.astype(np.float32))
    gdist.reduce_accumulators(accum, dst=0)
    if rank == 0:
        out, n = gdist.resolve_host(accum)
        np.save(os.path.join(out_dir, "rgb.npy"), out.numpy())
        np.save(os.path.join(out_dir, "counts.npy"), n.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_sample_ranges_tile_exactly():
    from grayshift_b200.dist import sample_range

    for spp in (1, 7, 100, 4096, 10000):
        for world in (1, 2, 3, 4, 8):
            r = [sample_range(spp, g, world) for g in range(world)]
            assert r[0][0] == 0 and r[-1][1] == spp
            assert all(r[g][1] == r[g + 1][0] for g in range(world - 1))
            sizes = [e - b for b, e in r]
            assert max(sizes) - min(sizes) <= 1


def test_two_ranks_sum_to_the_single_rank_render(tmp_path):
    import oracle_ffi as orc

    spp, world = 10, 2
    mp.spawn(_worker, args=(world, 29500 + os.getpid() % 2000, spp, str(tmp_path)), nprocs=world, join=True)
    rgb = np.load(tmp_path / "rgb.npy")
    counts = np.load(tmp_path / "counts.npy")
    single, c1, _, _ = orc.OracleScene.build("c3", width=24, spp=spp).render(seed=9, rng=orc.OracleScene.PHILOX, threads=2)
    assert np.all(counts == spp) and np.all(c1 == spp)
    # identical sample sets (Philox is keyed by the global sample index); only fp32 summation order differs
    np.testing.assert_allclose(rgb, single.reshape(-1, 3), rtol=2e-5, atol=1e-6)


def test_reduce_is_a_no_op_without_a_process_group():
    from grayshift_b200.dist import reduce_accumulators, resolve_host

    a = torch.tensor([[2.0, 4.0, 6.0, 2.0], [0.0, 0.0, 0.0, 0.0]])
    assert reduce_accumulators(a) is a
    rgb, n = resolve_host(a)
    assert rgb.tolist() == [[1.0, 2.0, 3.0], [0.0, 0.0, 0.0]] and n.tolist() == [2, 0]
