"""Image sharding and the final metric gather of the multi-GPU evaluation.

The hot path has no data-path collective: images are split into contiguous blocks, one process per GPU, replicated
weights.  The only communication is one gather of a fixed-width float32 record per image at the end, replacing
`accelerator.gather_for_metrics(..., use_gather_object=True)` (lib/engine/train_diff_hand_obj.py:333-335)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist

RECORD_WIDTH = 63 + 9   # fused wrist-relative joints (21 x 3) + fused object pose (rot6d + translation)


def shard_range(n_images: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous block of images owned by `rank` (the first n_images % world ranks get one extra)."""
    base, extra = divmod(n_images, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def image_record(agg_hand_joint: torch.Tensor, agg_obj_6d: torch.Tensor) -> torch.Tensor:
    """(n, 21, 3) f32, (n, 9) f64 -> (n, 72) f32 record that the gather moves."""
    n = agg_hand_joint.shape[0]
    return torch.cat([agg_hand_joint.reshape(n, 63).float(), agg_obj_6d.float()], dim=1).contiguous()


def gather_records(local: torch.Tensor, n_images: int) -> torch.Tensor:
    """All ranks receive the (n_images, W) table in image order.  Shards may differ by one row: they are padded to the
    largest shard for the fixed-size collective and trimmed afterwards."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(n_images, world, r) for r in range(world)]
    width = local.shape[1]
    nmax = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros((nmax, width), dtype=local.dtype, device=local.device)
    buf[: local.shape[0]] = local
    out = torch.empty((world, nmax, width), dtype=local.dtype, device=local.device)
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out, buf)
    else:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        out = torch.stack(parts, 0)
    return torch.cat([out[r, : hi - lo] for r, (lo, hi) in enumerate(sizes)], dim=0)
