"""Image-sharded data parallelism for the hot path: one process per GPU, contiguous image
ranges per rank (the reference's own manual-sharding idea, `--images_range`,
inference.py:120-123), codebook and code table replicated, and exactly one collective at the
end -- an all-reduce of {sum of stream bytes, sum of pixels, sum of squared error}, 3 x fp64 =
24 bytes per rank (NCCL over NVLink on GPUs, gloo in the CPU tests).  The path has no exchange
step, so there is no data-path collective.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) of `n_items` for `rank`; sizes differ by at most one."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def reduce_rate_distortion(total_bytes: float, total_pixels: float, total_sqerr: float, device=None):
    """All-reduce(SUM) of the three scalars -> (bytes, pixels, sqerr, bpp) over all ranks."""
    t = torch.tensor([total_bytes, total_pixels, total_sqerr], dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    b, p, s = t.tolist()
    return b, p, s, (8.0 * b / p if p else 0.0)


def gather_sizes(sizes: torch.Tensor) -> torch.Tensor:
    """All-gather of the per-image stream sizes [B_local,5] -> [B_total,5] in rank order, so rank 0
    can print per-image bpp like the reference's bpp.txt (inference.py:169-171).  Equal B_local."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return sizes
    out = [torch.empty_like(sizes) for _ in range(dist.get_world_size())]
    dist.all_gather(out, sizes.contiguous())
    return torch.cat(out, 0)
