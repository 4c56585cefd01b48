"""Multi-GPU rendering on the B200 box: the single-process C-ABI path (fw_render_multi: NCCL reduce / fused peer-memory
reduce + resolve) and the one-process-per-GPU torch.distributed path (firework_b200/distributed.py) must both reproduce
the single-GPU render — same samples, only the fp32 summation order across devices differs.  Tests that need two devices
skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

from conftest import native_scene, params_for
from firework_b200 import _native as N
from firework_b200.scenes import CONFIGS

pytestmark = pytest.mark.gpu


def _n_devices():
    return N.lib().fw_device_count()


def test_render_multi_on_one_device_is_fw_render():
    ns = native_scene("cornell_box")
    p = params_for("cornell_box", 96, 96, 6, seed=8)
    rgb, s, st = ns.render(p)
    for mode in ("nccl", "peer"):
        rgb1, s1, st1 = ns.render_multi(p, 1, reduce=mode)
        assert np.array_equal(rgb1, rgb) and np.array_equal(s1, s)
        assert st1["samples"] == st["samples"] and st1["rays"] == st["rays"]
    ns.close()


@pytest.mark.parametrize("name,w,h,spp", [("cornell_box", 128, 128, 12), ("teapot", 192, 108, 6), ("hdri_test", 200, 100, 8),
                                          ("part2_all", 192, 108, 5)])
@pytest.mark.parametrize("mode", ["nccl", "peer"])
def test_render_multi_matches_single_gpu(name, w, h, spp, mode):
    n = min(_n_devices(), 4)
    if n < 2:
        pytest.skip("needs two CUDA devices")
    ns = native_scene(name)
    p = params_for(name, w, h, spp, seed=3)
    rgb, s, st = ns.render(p)
    for g in sorted({2, n}):
        rgbm, sm, stm = ns.render_multi(p, g, reduce=mode)
        assert stm["samples"] == st["samples"] and stm["rays"] == st["rays"]          # the very same paths
        ok = np.isfinite(s)
        assert np.array_equal(ok, np.isfinite(sm))
        assert np.allclose(sm[ok], s[ok], rtol=2e-5, atol=1e-5)
        assert np.abs(rgbm.astype(int) - rgb.astype(int)).max() <= 1
        # a sample range that does not divide evenly, and one with fewer samples than devices
        q = params_for(name, w, h, spp, seed=3, sample_begin=1, sample_count=1)
        _, s1, _ = ns.render(q)
        _, s1m, _ = ns.render_multi(q, g, reduce=mode)
        assert np.array_equal(s1m, s1)     # one device renders the single sample, the others add zeros
    ns.close()


def test_texture_store_hit_on_another_device_reads_the_texels_back():
    """fw_scene_set_hdr skips the host copy when the content is resident anywhere; a commit onto a device that does not
    hold it must fetch the texels from the array that does (api.cu host_texels)."""
    if _n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    from conftest import scene_text, ASSETS
    from firework_b200.engine import NativeScene, release_cached_memory, texture_store_stats
    release_cached_memory()
    text = scene_text("hdri_test")
    p = params_for("hdri_test", 200, 100, 8, seed=5)
    a = NativeScene(text, device=0, asset_dir=ASSETS)
    _, sa, _ = a.render(p, want_rgb=False)
    before = texture_store_stats()
    b = NativeScene(text, device=1, asset_dir=ASSETS)        # set-time hit (device 0), commit on device 1
    _, sb, _ = b.render(p, want_rgb=False)
    after = texture_store_stats()
    assert after["uploads"] - before["uploads"] == 1 and after["arrays"] == 2
    assert np.array_equal(sa, sb)
    _, sm, _ = b.render_multi(p, 2, devices=[1, 0], want_rgb=False)   # its replica on device 0 shares a's array
    assert texture_store_stats()["arrays"] == 2
    assert np.allclose(sm, sa, rtol=2e-5, atol=1e-5)
    a.close(); b.close()
    release_cached_memory()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _rank_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch
    import torch.distributed as dist
    from conftest import ASSETS, scene_text
    from firework_b200.distributed import GpuShardRenderer
    from firework_b200.engine import NativeScene
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    ns = NativeScene(scene_text("cornell_box"), device=rank, asset_dir=ASSETS)
    r = CONFIGS["cornell_box"].renderer(width=300, height=300, samples=16, seed=4)
    gs = GpuShardRenderer(ns, r, rank)
    for step in range(3):        # repeated steps: a resolve that raced the reduce would show up as a dim image
        img, total = gs.render(16, rank, world)
        if rank == 0:
            np.save(os.path.join(out_dir, f"img{step}.npy"), img)
            np.save(os.path.join(out_dir, f"sum{step}.npy"), total.cpu().numpy())
        else:
            assert img is None
    ns.close()
    dist.destroy_process_group()


def test_two_rank_nccl_render_equals_single_gpu(tmp_path):
    """rank 0 must not resolve before the NCCL reduce has landed.  Two ranks over NCCL, image and
    sums compared with the one-GPU render of the same samples."""
    if _n_devices() < 2:
        pytest.skip("needs two CUDA devices")
    import torch.multiprocessing as mp
    mp.spawn(_rank_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ns = native_scene("cornell_box")
    rgb, s, _ = ns.render(params_for("cornell_box", 300, 300, 16, seed=4))
    ns.close()
    for step in range(3):
        got = np.load(tmp_path / f"sum{step}.npy").reshape(s.shape)
        img = np.load(tmp_path / f"img{step}.npy")
        assert np.allclose(got, s, rtol=2e-5, atol=1e-5)
        assert np.abs(img.astype(int) - rgb.astype(int)).max() <= 1
