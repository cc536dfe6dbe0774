"""multi_gpu.py -- sample-partitioned rendering across ranks (one process per GPU).

The reference data-parallelises over 20-row bands inside one process (integrator.cc:53-71).  Across
GPUs the scene is replicated and the SAMPLE axis is split instead: rank r of N renders the sample
indices [r*spp, (r+1)*spp) of every pixel.  The sampler is keyed by (pixel, sample index), so the N
partial films are exactly the terms of the N*spp-sample image; they are summed with ONE collective
(reduce to rank 0, NCCL over NVLink on GPUs) and the reference's Clamp01(mean) (integrator.cc:108)
is applied AFTER the reduce, on rank 0.

`render_fn(sample_begin, sample_count) -> torch.Tensor` abstracts the per-rank renderer so the same
logic is exercised on CPU (gloo) in tests and on B200s (NCCL) in bench.py.
"""
from __future__ import annotations


def sample_range(spp_per_rank: int, rank: int) -> tuple[int, int]:
    """Weak scaling: every rank renders spp_per_rank samples; returns (sample_begin, sample_count)."""
    return rank * spp_per_rank, spp_per_rank


def split_samples(spp_total: int, rank: int, world: int) -> tuple[int, int]:
    """Strong scaling: spp_total samples split as evenly as possible; returns (sample_begin, sample_count)."""
    base, rem = divmod(spp_total, world)
    begin = rank * base + min(rank, rem)
    return begin, base + (1 if rank < rem else 0)


def reduce_film(film_sum, dist, dst: int = 0):
    """Sum the raw radiance films onto rank `dst` (in place).  No-op for a single rank."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film_sum, dst=dst, op=dist.ReduceOp.SUM)
    return film_sum


def finalize(film_sum, spp_total: int):
    """Clamp01(sum / spp) -- the value FFilmView::AddColor receives (integrator.cc:102-108)."""
    return (film_sum * (1.0 / spp_total)).clamp_(0.0, 1.0)


def render_partitioned(render_fn, spp_per_rank: int, dist=None):
    """Render this rank's share, reduce to rank 0, and return (film_sum, spp_total)."""
    rank = dist.get_rank() if dist is not None and dist.is_initialized() else 0
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    begin, count = sample_range(spp_per_rank, rank)
    film = render_fn(begin, count)
    reduce_film(film, dist, 0)
    return film, spp_per_rank * world
