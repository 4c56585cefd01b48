"""Sample-range checkpoint / resume and progressive output.

The fp32 sum buffer of samples [0, k) *is* a resumable checkpoint: because the RNG is keyed by the global
sample index, rendering [0, k) now and [k, n) later gives the same sums as one render of [0, n) (up to the
fp32 add of the two partial sums per pixel).  This is the same mechanism as the multi-GPU sample sharding.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

import hashlib

from .engine import NativeScene, resolve_host


def render_fingerprint(scene, renderer) -> str:
    """Identifies WHAT is being rendered: the scene document and every renderer / camera parameter except the sample
    count (the sums are un-normalised, so a render may be resumed towards a LARGER total; a smaller one is refused).
    A checkpoint may only be resumed by the render that wrote it."""
    p = renderer.params()
    h = hashlib.sha256()
    h.update(scene.to_yaml().encode("utf-8"))
    h.update(repr((p.width, p.height, p.use_bvh, p.gamma, tuple(p.cam_pos), tuple(p.look_at), p.vfov,
                   p.aperture, p.focus_dist, p.seed)).encode())
    return h.hexdigest()


# ---- the native driver's checkpoint format (csrc/cli_main.cpp): a render begun by one driver can be finished by the other -----------
_FWCK_MAGIC = b"FWCKPT01"


def fwck_fingerprint(yaml_text: str, p) -> int:
    """crc32 << 32 | adler32 over the scene text followed by the packed render parameters (cli_main.cpp `fingerprint`)."""
    import struct
    import zlib
    data = yaml_text.encode("utf-8")
    packed = struct.pack("<3I10fQ", p.width, p.height, p.use_bvh, p.gamma, *p.cam_pos, *p.look_at, p.vfov, p.aperture, p.focus_dist, p.seed)
    c = zlib.crc32(packed, zlib.crc32(data)) & 0xFFFFFFFF
    a = zlib.adler32(packed, zlib.adler32(data)) & 0xFFFFFFFF
    return (c << 32) | a


def fwck_load(path: str):
    import struct
    with open(path, "rb") as f:
        head = f.read(40)
        if len(head) != 40 or head[:8] != _FWCK_MAGIC:
            raise ValueError(f"{path!r} is not a firework checkpoint")
        w, h, done, seed, fp = struct.unpack("<IIQQQ", head[8:])
        sums = np.fromfile(f, np.float32)
    if sums.size != w * h * 3:
        raise ValueError(f"checkpoint {path!r} is truncated")
    return sums.reshape(h, w, 3), int(done), int(seed), int(fp)


def fwck_save(path: str, sums: np.ndarray, done: int, seed: int, fp: int) -> None:
    import struct
    tmp = path + ".tmp"
    with open(tmp, "wb") as f:
        f.write(_FWCK_MAGIC + struct.pack("<IIQQQ", sums.shape[1], sums.shape[0], done, seed, fp))
        np.ascontiguousarray(sums, np.float32).tofile(f)
    os.replace(tmp, path)


def render_progressive(scene, renderer, checkpoint: Optional[str] = None, chunk: int = 64, device: int = 0,
                       on_step=None) -> Tuple[np.ndarray, dict]:
    """Renders `renderer`'s samples in chunks, optionally resuming from / saving to `checkpoint` (.npz, or the native driver's
    .fwck format).
    Returns (rgb (H, W, 3) u8 resolved over ALL samples done, info)."""
    p0 = renderer.params()
    h, w, total = p0.height, p0.width, p0.samples
    done = 0
    sums = np.zeros((h, w, 3), np.float32)
    native_format = bool(checkpoint) and checkpoint.endswith(".fwck")     # the native driver's file (cli_main.cpp)
    fp = fwck_fingerprint(scene.to_yaml(), p0) if native_format else render_fingerprint(scene, renderer)
    if checkpoint and os.path.exists(checkpoint):
        if native_format:
            ck_sums, ck_done, ck_seed, ck_fp = fwck_load(checkpoint)
            if ck_fp != fp or ck_seed != p0.seed or ck_sums.shape != sums.shape:
                raise ValueError("checkpoint does not match this render (scene, camera, renderer parameters or size differ)")
            sums, done = ck_sums.astype(np.float32), ck_done
        else:
            ck = np.load(checkpoint)
            if "fingerprint" not in ck.files or str(ck["fingerprint"]) != fp or ck["sums"].shape != sums.shape:
                raise ValueError("checkpoint does not match this render (scene, camera, renderer parameters or size differ)")
            sums, done = ck["sums"].astype(np.float32), int(ck["done"])
        if not (0 <= done <= total):
            raise ValueError(f"checkpoint holds {done} samples but this render has {total}")
    ns = NativeScene.from_scene(scene, device)
    rays = 0
    try:
        while done < total:
            n = min(max(chunk, 1), total - done)
            _, part, st = ns.render(renderer.params(sample_begin=done, sample_count=n), want_rgb=False)
            sums += part
            done += n
            rays += st["rays"]
            if checkpoint and native_format:
                fwck_save(checkpoint, sums, done, p0.seed, fp)
            elif checkpoint:
                tmp = checkpoint + ".tmp.npz"
                np.savez(tmp, sums=sums, done=done, seed=p0.seed, fingerprint=fp)
                os.replace(tmp, checkpoint)
            if on_step:
                on_step(done, total)
        # resolve (render.rs:184-189) on the device from the accumulated sums (fw_resolve_host: no torch involved)
        rgb = resolve_host(sums, done if done else 1, p0.gamma, device)
    finally:
        ns.close()
    return rgb, {"samples_done": done, "rays": rays}
