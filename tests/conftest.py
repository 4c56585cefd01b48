import gzip
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from firework_b200.assets import load_asset  # noqa: E402
from firework_b200.scenes import CONFIGS, SCENE_DIR  # noqa: E402
from firework_b200.serde_yaml import loads  # noqa: E402

ASSETS = os.path.join(SCENE_DIR, "assets")
ALL_SCENES = ["random_spheres", "cornell_box", "suzanne", "teapot", "hdri_test", "earth", "part2_all", "conics",
              "conics_cli", "volume", "heightmap"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _has_gpu():
    try:
        from firework_b200 import _native
        return _native.lib().fw_device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # `-m gpu` on a box without a device must fail loudly, not skip: only skip when not explicitly selected.
    if "gpu" in (config.getoption("-m") or ""):
        return
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


_text_cache = {}


def scene_text(name):
    if name not in _text_cache:
        p = CONFIGS[name].path()
        with (gzip.open(p, "rt") if p.endswith(".gz") else open(p)) as f:
            _text_cache[name] = f.read()
    return _text_cache[name]


def asset_loader(path, kind):
    return load_asset(path, kind, ASSETS)


_doc_cache = {}


def scene_doc(name):
    if name not in _doc_cache:
        _doc_cache[name] = loads(scene_text(name))
    return _doc_cache[name]


def oracle_scene(name, use_bvh=None, fast=False):
    from oracle.oracle import OracleScene
    cfg = CONFIGS[name]
    return OracleScene(scene_doc(name), cfg.use_bvh if use_bvh is None else use_bvh, asset_loader=asset_loader, fast=fast)


def native_scene(name, commit=True):
    from firework_b200.engine import NativeScene
    return NativeScene(scene_text(name), asset_dir=ASSETS, commit=commit)


def params_for(name, width, height, samples, seed=0, sample_begin=0, sample_count=None):
    r = CONFIGS[name].renderer(width=width, height=height, samples=samples, seed=seed)
    return r.params(sample_begin=sample_begin, sample_count=sample_count)


def psnr_u8(a, b):
    mse = np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2)
    return float("inf") if mse == 0 else 10.0 * np.log10(255.0 ** 2 / mse)
