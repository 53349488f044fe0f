"""Multi-GPU sharding of signature batches: contiguous index ranges per rank, one
process per GPU, no data-path collective except the all-gather of the per-shard validity bitmaps.

The reference has no inter-process parallelism (rayon only, consensus/src/pipeline/virtual_processor/
utxo_validation.rs:269-277); signatures are independent units, so shards need no exchange step.
Shard sizes are padded to a multiple of 8 items so that every shard's bitmap is byte aligned and the
gathered buffer is exactly the global bitmap (bit i, LSB first in each byte, = item i valid).
"""
import numpy as np


def shard_bounds(n_items, world_size):
    """Contiguous, byte-aligned shards: returns a list of (lo, hi) with hi-lo a multiple of 8 except the last."""
    per = -(-n_items // world_size)  # ceil
    per = (per + 7) // 8 * 8
    return [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world_size)]


def shard_range(n_items, rank, world_size):
    return shard_bounds(n_items, world_size)[rank]


def all_gather_bitmaps(local_bitmap, n_items, group=None):
    """local_bitmap: torch uint8 tensor holding this rank's packed shard (on the device of the backend).
    Returns the global bitmap ((n_items+7)//8 bytes), identical on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    bounds = shard_bounds(n_items, world)
    per_bytes = (bounds[0][1] - bounds[0][0] + 7) // 8
    buf = torch.zeros(per_bytes, dtype=torch.uint8, device=local_bitmap.device)
    buf[:local_bitmap.numel()] = local_bitmap
    out = torch.empty(per_bytes * world, dtype=torch.uint8, device=local_bitmap.device)
    dist.all_gather_into_tensor(out, buf, group=group)
    return out[:(n_items + 7) // 8]
