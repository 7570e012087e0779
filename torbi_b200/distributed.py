"""One-process-per-GPU sharding of the decode path (torchrun style).

The path is embarrassingly parallel over sequences / files, so ranks never
exchange data while decoding: each rank takes a longest-first (LPT) share of
the items and, if the caller wants the whole result everywhere, the int32 index
rows are gathered at the end.  Works with any torch.distributed backend
(`nccl` on the GPU box, `gloo` in the CPU tests).
"""
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist

from .core import lpt_shards


def rank_shard(costs: Sequence[int], rank: Optional[int] = None,
               world_size: Optional[int] = None) -> List[int]:
    """Indices of the items this rank decodes (ascending).  Every rank computes
    the same LPT partition from the same `costs` (frame counts or file sizes),
    so no communication is needed to agree on it."""
    if rank is None:
        rank = dist.get_rank()
    if world_size is None:
        world_size = dist.get_world_size()
    shards = lpt_shards(costs, world_size)
    return shards[rank] if rank < len(shards) else []


def gather_indices(local_indices: torch.Tensor, items: Sequence[int], batch: int,
                   group=None) -> torch.Tensor:
    """All-gather the decoded rows: every rank passes its `(len(items), T)` int32
    rows and receives the full `(batch, T)` tensor in original item order.  This
    is the only collective of the path (north star: "no collective beyond
    gathering the index arrays")."""
    world_size = dist.get_world_size(group)
    frames = local_indices.shape[1]
    counts = [None] * world_size
    dist.all_gather_object(counts, list(items), group=group)
    longest = max(len(c) for c in counts)
    padded = local_indices.new_zeros((longest, frames))
    padded[:local_indices.shape[0]] = local_indices
    parts = [torch.empty_like(padded) for _ in range(world_size)]
    dist.all_gather(parts, padded, group=group)
    full = local_indices.new_empty((batch, frames))
    for part, rank_items in zip(parts, counts):
        if rank_items:
            index = torch.tensor(rank_items, dtype=torch.long, device=full.device)
            full.index_copy_(0, index, part[:len(rank_items)])
    return full


def shard_files(input_files: Sequence, output_files: Sequence, rank: Optional[int] = None,
                world_size: Optional[int] = None) -> Tuple[list, list]:
    """This rank's share of a corpus for from_files_to_files (core.py:310-368): the files are dealt by
    size -- a proxy for the frame count -- longest first, the partition from_files_to_files itself uses
    across `gpu=[...]` inside one process.  Every rank computes the same partition from the same file
    list; nothing is communicated."""
    sizes = [os.path.getsize(file) for file in input_files]
    mine = rank_shard(sizes, rank, world_size)
    return [input_files[i] for i in mine], [output_files[i] for i in mine]


def from_files_to_files(input_files, output_files, transition_file=None, initial_file=None, log_probs=False,
                        gpu: Optional[int] = None, num_threads: Optional[int] = None) -> None:
    """torbi.from_files_to_files for one process per GPU (torchrun): every rank decodes its share of the
    files on its own GPU (LOCAL_RANK unless `gpu` is given) and writes its own outputs; there is no
    collective on the path -- callers that need "all files are written" add a barrier."""
    import torbi_b200
    inputs, outputs = shard_files(input_files, output_files)
    if gpu is None:
        gpu = int(os.environ.get('LOCAL_RANK', dist.get_rank() if dist.is_initialized() else 0))
    if inputs:
        torbi_b200.from_files_to_files(inputs, outputs, transition_file, initial_file, log_probs, gpu=gpu,
                                       num_threads=num_threads)
