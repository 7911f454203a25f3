"""One process per GPU: every rank builds one key-range shard of the suffix array.

The text is replicated (every compare needs random access to all of it), every rank derives the same
splitters from the same key histogram, so the data path needs NO collective.  The only exchange is
metadata: each rank needs the last suffix of the previous non-empty shard to repair one seam LCP
(reference rule: sufr_builder.rs:893-902).  That is one tiny all_gather.
"""
from __future__ import annotations

from typing import List, Optional, Tuple


def previous_last_suffix(meta: List[Tuple[int, int, int]], rank: int) -> Optional[int]:
    """meta[r] = (num_suffixes, first_suffix, last_suffix) of rank r.  Returns the last suffix of the
    nearest lower rank that has suffixes, or None when this shard starts the suffix array."""
    for r in range(rank - 1, -1, -1):
        if meta[r][0] > 0:
            return meta[r][2]
    return None


def shard_layout(meta: List[Tuple[int, int, int]]) -> Tuple[List[int], int]:
    """Offsets of every shard in the concatenated suffix array, and the total."""
    offs, acc = [], 0
    for cnt, _, _ in meta:
        offs.append(acc)
        acc += cnt
    return offs, acc


def gather_meta(num_suffixes: int, first: int, last: int, group=None) -> List[Tuple[int, int, int]]:
    """all_gather of (count, first, last) over the default (or given) torch.distributed group."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor([num_suffixes, first, last], dtype=torch.int64, device=dev)
    out = [torch.zeros(3, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    return [tuple(int(x) for x in t.tolist()) for t in out]


def finish_shard(result, group=None):
    """Seam repair + consistency check for a shard built with (rank, world_size).  Returns the meta list."""
    import torch.distributed as dist

    rank = dist.get_rank(group)
    meta = gather_meta(result.num_suffixes, result.first_suffix, result.last_suffix, group)
    offs, total = shard_layout(meta)
    result.set_shard_layout(offs[rank], total)
    prev = previous_last_suffix(meta, rank)
    if prev is not None and result.num_suffixes:
        result.patch_seam(prev)
    return meta
