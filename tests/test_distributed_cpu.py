"""world_size-2 gloo test of the sample-range sharding host logic (firework_b200/distributed.py) on CPU.
The shard renderer is a stand-in (the oracle, as the checker); the GPU path uses the same driver with
GpuShardRenderer (tests/test_gpu_parity.py covers that on the B200)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from firework_b200.distributed import render_sharded, shard_range


def test_shard_range_partitions():
    for samples in (1, 2, 7, 32, 1000, 4096):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(samples, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == samples
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from conftest import oracle_scene, params_for
    from oracle import oracle as orc
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = oracle_scene("cornell_box")
    samples, w, h = 6, 32, 32

    def render_shard(begin, count):
        p = params_for("cornell_box", w, h, samples, seed=5, sample_begin=begin, sample_count=count)
        _, s, _ = sc.render(p, want_rgb=False, threads=2)
        return torch.from_numpy(s.copy())

    def resolve(total):
        return orc.resolve(total.numpy(), samples, 2.2)

    img, total = render_sharded(render_shard, resolve, samples, rank, world)
    if rank == 0:
        np.save(os.path.join(out_dir, "img.npy"), img)
        np.save(os.path.join(out_dir, "sum.npy"), total.numpy())
    else:
        assert img is None
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    from conftest import oracle_scene, params_for
    sc = oracle_scene("cornell_box")
    rgb, s, _ = sc.render(params_for("cornell_box", 32, 32, 6, seed=5))
    got_sum = np.load(tmp_path / "sum.npy")
    got_img = np.load(tmp_path / "img.npy")
    # same samples, only the fp32 summation order differs: ((s0+s1+s2) + (s3+s4+s5)) vs sequential
    assert np.allclose(got_sum, s, rtol=1e-5, atol=1e-6)
    assert np.abs(got_img.astype(int) - rgb.astype(int)).max() <= 1
