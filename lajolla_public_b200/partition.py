"""Multi-GPU partition of a render: every (pixel, sample) is independent, so the path shards by
sample.  The scene is replicated; rank r renders a contiguous block of the per-pixel sample indices with the
PCG streams that belong to them (lj_pcg.h: stream = hash(pixel * spp_total + sample)), leaves raw sums in its
film, and ONE sum-reduce of the w*h*3 fp32 film follows.  No collective touches the data path itself.

The reference's decomposition is image-space tiles over a thread pool (render.cpp:75-100, parallel.cpp); tiles
of a 512x512 image would under-fill 148 SMs per GPU, a sample split keeps every GPU on the whole image.
"""
from typing import List, Tuple


def sample_ranges(total_spp: int, world: int) -> List[Tuple[int, int]]:
    """Split [0, total_spp) into `world` contiguous blocks whose sizes differ by at most one (strong scaling:
    the image's sample count is fixed).  Ranks beyond total_spp get an empty block."""
    if total_spp < 0 or world <= 0:
        raise ValueError("total_spp >= 0 and world > 0 required")
    base, extra = divmod(total_spp, world)
    out, begin = [], 0
    for r in range(world):
        n = base + (1 if r < extra else 0)
        out.append((begin, begin + n))
        begin += n
    return out


def weak_range(rank: int, world: int, spp_per_rank: int) -> Tuple[int, int, int]:
    """Weak scaling (bench.py): every rank renders spp_per_rank samples of a (spp_per_rank * world)-sample image.
    Returns (total_spp, sample_begin, sample_end)."""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    return spp_per_rank * world, rank * spp_per_rank, (rank + 1) * spp_per_rank


def reduce_film(film, dist=None, dst: int = 0):
    """Sum the per-rank raw films into rank `dst` (the one collective of the system).  `film` is a torch tensor
    on the backend's device (CUDA for nccl, CPU for gloo); with dist None (single process) it is returned as is."""
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.reduce(film, dst=dst, op=dist.ReduceOp.SUM)
    return film


def render_partitioned(scene, total_spp: int, rank: int, world: int, film, stream_ptr=None, dist=None, **kw):
    """Render this rank's block of a total_spp image into `film` (torch tensor, h x w x 3 fp32, on the device for the
    CUDA library) and reduce.  Rank 0 holds the normalised image afterwards; other ranks hold their raw sums."""
    begin, end = sample_ranges(total_spp, world)[rank]
    stats = None
    if end > begin:
        stats = scene.render_device(film.data_ptr(), stream_ptr, spp=total_spp, sample_begin=begin, sample_end=end,
                                    normalize=False, **kw)
    else:
        film.zero_()
    reduce_film(film, dist, 0)
    if rank == 0 and total_spp > 0:
        film.mul_(1.0 / total_spp)
    return stats
