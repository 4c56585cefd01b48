"""Host-side asset decoding for ImageTexture / HdrEnvironment paths.

The reference decodes with the `image` crate (texture.rs:288, hdri_test.rs:45-67).  Here PIL decodes once on
the host and the *same* texel array is handed to whoever consumes the scene, so decoder differences cannot
affect parity.  The HDR map of examples/hdri_test.rs (urban_street_04_4k.hdr) is not in the reference
repository; `synthetic_hdr:<W>x<H>:<seed>` names a procedural stand-in of the same resolution.
"""
from __future__ import annotations

import os
from functools import lru_cache

import numpy as np


@lru_cache(maxsize=8)
def synthetic_hdr(width: int, height: int, seed: int) -> np.ndarray:
    """Smooth sky gradient + warm ground + one small bright sun disc (~50) + low-amplitude hash noise.
    Returns (H, W, 3) fp32, row 0 = top (v = 1)."""
    y = (np.arange(height, dtype=np.float32) + 0.5) / height          # 0 = top
    x = (np.arange(width, dtype=np.float32) + 0.5) / width
    elev = (0.5 - y) * np.pi                                             # +pi/2 at top
    az = (x * 2.0 - 1.0) * np.pi
    e, a = np.meshgrid(elev, az, indexing="ij")
    t = np.clip(np.sin(e) * 0.5 + 0.5, 0, 1).astype(np.float32)
    sky = np.stack([1.0 - 0.6 * t, 1.0 - 0.35 * t, np.ones_like(t)], -1) * (0.6 + 0.8 * t[..., None])
    ground = np.stack([0.35 + 0 * t, 0.3 + 0 * t, 0.25 + 0 * t], -1)
    img = np.where((e > 0)[..., None], sky, ground).astype(np.float32)
    # sun at elevation 40 deg, azimuth 60 deg, angular radius ~1.5 deg
    se, sa = np.radians(40.0), np.radians(60.0)
    cosang = np.sin(e) * np.sin(se) + np.cos(e) * np.cos(se) * np.cos(a - sa)
    img[cosang > np.cos(np.radians(1.5))] = np.array([50.0, 46.0, 40.0], np.float32)
    # hash noise (deterministic, seed-keyed)
    ii, jj = np.meshgrid(np.arange(height, dtype=np.uint32), np.arange(width, dtype=np.uint32), indexing="ij")
    hsh = (ii * np.uint32(73856093)) ^ (jj * np.uint32(19349663)) ^ np.uint32((seed * 83492791) & 0xFFFFFFFF)
    hsh = (hsh ^ (hsh >> np.uint32(13))) * np.uint32(1274126177)
    noise = ((hsh >> np.uint32(8)).astype(np.float32) / 16777216.0 - 0.5) * 0.04
    img *= (1.0 + noise[..., None])
    return np.ascontiguousarray(img, dtype=np.float32)


def load_asset(path: str, kind: str, asset_dir: str | None = None, registry: dict | None = None) -> np.ndarray:
    """kind = "image" -> (H, W, 4) u8 RGBA   |   kind = "hdr" -> (H, W, 3) fp32."""
    if registry and path in registry:
        return registry[path]
    if path.startswith("synthetic_hdr:"):
        _, dims, seed = path.split(":")
        w, h = dims.split("x")
        return synthetic_hdr(int(w), int(h), int(seed))
    cands = [path]
    if asset_dir:
        cands = [os.path.join(asset_dir, path), os.path.join(asset_dir, os.path.basename(path)),
                 os.path.join(asset_dir, "assets", os.path.basename(path)), path]
    for c in cands:
        if os.path.exists(c):
            if kind == "hdr":
                if c.endswith(".npy"):
                    return np.ascontiguousarray(np.load(c), dtype=np.float32)
                return load_hdr(c)
            from PIL import Image
            return np.ascontiguousarray(np.asarray(Image.open(c).convert("RGBA")), dtype=np.uint8)
    raise FileNotFoundError(f"asset {path!r} not found (searched {cands})")


def load_hdr(path: str) -> np.ndarray:
    """Radiance RGBE .hdr -> (H, W, 3) fp32, row 0 = top — the native decoder (fw_hdr_load), which restates
    image 0.23.9 `HdrDecoder::read_image_hdr` (examples/hdri_test.rs:45-67)."""
    import ctypes as C

    from . import _native as N
    L = N.lib()
    w, h, p = C.c_uint32(), C.c_uint32(), C.POINTER(C.c_float)()
    N.check(L.fw_hdr_load(path.encode(), C.byref(w), C.byref(h), C.byref(p)))
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 3)).copy()
    finally:
        L.fw_hdr_free(p)


def load_obj(path: str):
    """Wavefront OBJ -> list of models {name, positions (N,3), normals (N,3)|None, texcoords (N,2)|None, indices (M,)}:
    the native loader (fw_obj_load), which restates tobj 1.0.0 `load_obj` (examples/suzanne.rs:19)."""
    import ctypes as C

    from . import _native as N
    L = N.lib()
    h = C.c_void_p()
    N.check(L.fw_obj_load(path.encode(), C.byref(h)))
    try:
        models = []
        for m in range(L.fw_obj_num_models(h)):
            sz = np.zeros(4, np.uint32)
            N.check(L.fw_obj_model_sizes(h, m, N.ptr(sz)))
            pos, nrm, tex = (np.zeros(int(n), np.float32) for n in sz[:3])
            idx = np.zeros(int(sz[3]), np.uint32)
            N.check(L.fw_obj_model_copy(h, m, N.ptr(pos), N.ptr(nrm), N.ptr(tex), N.ptr(idx)))
            models.append({"name": L.fw_obj_model_name(h, m).decode(), "positions": pos.reshape(-1, 3),
                           "normals": nrm.reshape(-1, 3) if len(nrm) else None,
                           "texcoords": tex.reshape(-1, 2) if len(tex) else None, "indices": idx})
        return models
    finally:
        L.fw_obj_destroy(h)


def load_image_native(path: str) -> np.ndarray:
    """PNG / baseline JPEG -> (H, W, 4) u8 RGBA through the library's own decoder (fw_image_load, csrc/images.cpp) — what the
    native command-line driver uses; `load_asset` decodes with PIL so that the oracle and the GPU see one texel array."""
    import ctypes as C
    from . import _native as N
    L = N.lib()
    w, h = C.c_uint32(), C.c_uint32()
    p = C.POINTER(C.c_uint8)()
    N.check(L.fw_image_load(path.encode(), C.byref(w), C.byref(h), C.byref(p)))
    try:
        return np.ctypeslib.as_array(p, shape=(h.value, w.value, 4)).copy()
    finally:
        L.fw_image_free(p)


def write_png(path: str, rgb: np.ndarray) -> None:
    """(H, W, 3) u8 -> PNG through fw_png_write (window.rs `save_image`)."""
    from . import _native as N
    rgb = np.ascontiguousarray(rgb, np.uint8)
    N.check(N.lib().fw_png_write(path.encode(), rgb.shape[1], rgb.shape[0], N.ptr(rgb)))
