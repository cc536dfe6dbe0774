"""multi_gpu.py -- the host-side ORDER of a sample-partitioned render, for the CPU (gloo) test of the N > 1 logic.

On B200s the whole of it lives in the library: jpbrt_comm_init gives a context its NCCL communicator,
jpbrt_sample_partition its share of the sample indices, jpbrt_read_film reduces the raw float32 films onto rank 0 with
one ncclReduce and applies the reference's Clamp01(mean) (integrator.cc:108) AFTER the reduce; jpbrt_render_multi does
the same inside one process (include/jetpbrt_b200.h, csrc/c_api.cu).  This module restates that order over
torch.distributed so that tests/test_multi_rank_gloo.py can run it with world_size 2 on CPUs, the oracle standing in
for the per-rank renderer.
"""
from __future__ import annotations


def split_samples(spp_total: int, rank: int, world: int) -> tuple[int, int]:
    """jpbrt_sample_partition: contiguous, as even as possible; returns (sample_begin, sample_count)."""
    base, rem = divmod(spp_total, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def render_partitioned(render_fn, spp_total: int, dist=None):
    """`render_fn(sample_begin, sample_count) -> torch.Tensor` of raw radiance sums.  Returns rank 0's (sum, spp_total)."""
    on = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    rank, world = (dist.get_rank(), dist.get_world_size()) if on else (0, 1)
    film = render_fn(*split_samples(spp_total, rank, world))
    if on:
        dist.reduce(film, dst=0, op=dist.ReduceOp.SUM)
    return film, spp_total


def finalize(film_sum, spp_total: int):
    """Clamp01(sum / spp) -- the value FFilmView::AddColor receives (integrator.cc:102-108); after the reduce."""
    return (film_sum * (1.0 / spp_total)).clamp_(0.0, 1.0)
