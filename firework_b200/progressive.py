"""Sample-range checkpoint / resume and progressive output.

The fp32 sum buffer of samples [0, k) *is* a resumable checkpoint: because the RNG is keyed by the global
sample index, rendering [0, k) now and [k, n) later gives the same sums as one render of [0, n) (up to the
fp32 add of the two partial sums per pixel).  This is the same mechanism as the multi-GPU sample sharding.
"""
from __future__ import annotations

import os
from typing import Optional, Tuple

import numpy as np

from .engine import NativeScene


def render_progressive(scene, renderer, checkpoint: Optional[str] = None, chunk: int = 64, device: int = 0,
                       on_step=None) -> Tuple[np.ndarray, dict]:
    """Renders `renderer`'s samples in chunks, optionally resuming from / saving to `checkpoint` (.npz).
    Returns (rgb (H, W, 3) u8 resolved over ALL samples done, info)."""
    p0 = renderer.params()
    h, w, total = p0.height, p0.width, p0.samples
    done = 0
    sums = np.zeros((h, w, 3), np.float32)
    if checkpoint and os.path.exists(checkpoint):
        ck = np.load(checkpoint)
        if ck["sums"].shape != sums.shape or int(ck["seed"]) != p0.seed:
            raise ValueError("checkpoint does not match this render (size or seed)")
        sums, done = ck["sums"].astype(np.float32), int(ck["done"])
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
                np.savez(tmp, sums=sums, done=done, seed=p0.seed)
                os.replace(tmp, checkpoint)
            if on_step:
                on_step(done, total)
        # resolve (render.rs:184-189) on the device from the accumulated sums
        rgb = resolve_sums(ns, sums, done if done else 1, p0.gamma)
    finally:
        ns.close()
    return rgb, {"samples_done": done, "rays": rays}


def resolve_sums(ns: NativeScene, sums: np.ndarray, samples: int, gamma: float) -> np.ndarray:
    import torch
    d = torch.from_numpy(np.ascontiguousarray(sums, np.float32)).reshape(-1).cuda(ns.device)
    out = torch.empty(d.numel(), dtype=torch.uint8, device=d.device)
    ns.resolve_device(d.data_ptr(), d.numel() // 3, samples, gamma, out.data_ptr(),
                      torch.cuda.current_stream(d.device).cuda_stream)
    torch.cuda.synchronize(d.device)
    return out.cpu().numpy().reshape(sums.shape)
