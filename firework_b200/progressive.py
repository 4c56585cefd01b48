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


def render_progressive(scene, renderer, checkpoint: Optional[str] = None, chunk: int = 64, device: int = 0,
                       on_step=None) -> Tuple[np.ndarray, dict]:
    """Renders `renderer`'s samples in chunks, optionally resuming from / saving to `checkpoint` (.npz).
    Returns (rgb (H, W, 3) u8 resolved over ALL samples done, info)."""
    p0 = renderer.params()
    h, w, total = p0.height, p0.width, p0.samples
    done = 0
    sums = np.zeros((h, w, 3), np.float32)
    fp = render_fingerprint(scene, renderer)
    if checkpoint and os.path.exists(checkpoint):
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
            if checkpoint:
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
