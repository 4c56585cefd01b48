"""Multi-GPU rendering: the sample range is split across ranks (one process per GPU); every rank renders ALL
pixels for its own disjoint sample indices into an fp32 sum buffer; one `reduce(SUM)` to rank 0 over
NCCL/NVLink is the only collective; rank 0 resolves (mean, gamma, quantise).  The RNG is keyed by the GLOBAL
sample index, so the union of samples is identical for any world size (results differ only by fp32 summation
order across ranks).

The reference has no multi-process path (src/render.rs:127 is a rayon loop); this is the sharding
BASELINE.json's north_star prescribes.  The host logic is backend-agnostic so that it is testable with
`gloo` on CPU (tests/test_distributed_cpu.py) with a stand-in shard renderer.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def shard_range(samples: int, rank: int, world: int) -> Tuple[int, int]:
    """[begin, end) of the sample indices rank renders: contiguous, disjoint, covering, sizes differ by <= 1."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * samples) // world, ((rank + 1) * samples) // world


def render_sharded(render_shard: Callable[[int, int], "object"], resolve: Callable[["object"], np.ndarray],
                   samples: int, rank: int, world: int, reduce_fn: Optional[Callable[["object"], None]] = None,
                   reduce_events=None):
    """Generic driver. `render_shard(begin, count)` returns this rank's fp32 sum tensor (any torch device);
    `reduce_fn(tensor)` sums it onto rank 0 in place (default: torch.distributed.reduce); `resolve(sum)` runs on
    rank 0 only.  `reduce_events`: optional (begin, end) torch.cuda.Event pair recorded on the current stream around
    the collective.  Returns (image on rank 0 | None, sum tensor)."""
    begin, end = shard_range(samples, rank, world)
    total = render_shard(begin, end - begin)
    if world > 1:
        if reduce_events:
            reduce_events[0].record()
        if reduce_fn is None:
            import torch.distributed as dist
            dist.reduce(total, dst=0, op=dist.ReduceOp.SUM)
        else:
            reduce_fn(total)
        if reduce_events:
            reduce_events[1].record()
    if rank == 0:
        return resolve(total), total
    return None, total


class GpuShardRenderer:
    """render_shard / resolve for one rank on its own GPU, through the C ABI's device-resident entry points.

    Stream ordering: every piece of device work of a step — the zero-fill of the sum buffer, the render kernels, the
    NCCL reduce, the resolve kernel and the device->host copy of the image — is issued on ONE explicit, non-default
    torch stream (`self.stream`), made current with `use_stream()` around the whole step.  Passing torch's default
    stream would not work: its handle is 0, which the C ABI reads as "use the scene's private stream", and that stream
    has no ordering against the stream torch / NCCL use (the resolve would read the buffer before the reduce wrote it).
    """

    def __init__(self, native_scene, renderer, device_index: int):
        import torch
        self.torch = torch
        self.ns = native_scene
        self.renderer = renderer
        self.device = torch.device("cuda", device_index)
        self.stream = torch.cuda.Stream(self.device)
        assert self.stream.cuda_stream != 0
        p = renderer.params()
        self.npix = p.width * p.height
        self.shape = (p.height, p.width, 3)
        self.last_stats = None

    def use_stream(self):
        return self.torch.cuda.stream(self.stream)

    def render_shard(self, begin: int, count: int):
        torch = self.torch
        with self.use_stream():
            d_sum = torch.zeros(self.npix * 3, dtype=torch.float32, device=self.device)
            if count > 0:
                p = self.renderer.params(sample_begin=begin, sample_count=count)
                self.last_stats = self.ns.render_accumulate_device(p, d_sum.data_ptr(), self.stream.cuda_stream)
        return d_sum

    def resolve(self, d_sum):
        torch = self.torch
        p = self.renderer.params()
        with self.use_stream():
            d_rgb = torch.empty(self.npix * 3, dtype=torch.uint8, device=self.device)
            self.ns.resolve_device(d_sum.data_ptr(), self.npix, p.samples, p.gamma, d_rgb.data_ptr(), self.stream.cuda_stream)
            out = d_rgb.cpu()          # same stream: ordered after the resolve kernel, synchronises it
        return out.numpy().reshape(self.shape)

    def render(self, samples: int, rank: int, world: int, reduce_events=None):
        """One sharded render step: (image on rank 0 | None, this rank's / the reduced sum tensor)."""
        with self.use_stream():    # dist.reduce orders itself against torch's CURRENT stream
            return render_sharded(self.render_shard, self.resolve, samples, rank, world, reduce_events=reduce_events)
