"""Approximate long-sequence mode: cut a sequence where two adjacent frames have
low normalised entropy, so the pieces decode independently and batch well.
Same results as /root/reference/torbi/chunk.py:12-85 (host logic; the pieces
feed the ragged batch kernel through data.collate)."""
import bisect
import math
from typing import List, Optional

import torch

import torbi_b200


def chunk(
    observation: torch.Tensor,
    min_chunk_size: Optional[int] = None,
    entropy_threshold: Optional[float] = None
) -> List[torch.Tensor]:
    """Chunk observations based on points of low entropy

    Arguments
        observation
            Time-varying categorical distribution in LOG space
            shape=(frames, states)
        min_chunk_size
            Minimum chunk size in frames; defaults to torbi_b200.MIN_CHUNK_SIZE
        entropy_threshold
            Threshold for entropy to allow splitting; defaults to
            torbi_b200.ENTROPY_THRESHOLD

    Returns
        chunks
            List of views into `observation`, in order (chunk.py:32-49)
    """
    # the reference binds these defaults at import (chunk.py:14-15); reading them
    # at call time lets torbi_b200.configure() take effect afterwards
    if min_chunk_size is None:
        min_chunk_size = torbi_b200.MIN_CHUNK_SIZE
    if entropy_threshold is None:
        entropy_threshold = torbi_b200.ENTROPY_THRESHOLD
    bounds = [0] + split(observation, min_chunk_size, entropy_threshold) + [observation.shape[0]]
    return [observation[start:stop] for start, stop in zip(bounds[:-1], bounds[1:])]


def split(observation, min_chunk_size=None, entropy_threshold=None) -> List[int]:
    """Find split points of minimum entropy (chunk.py:57-78): scanning from
    `min_chunk_size`, cut at the first i with frames i and i-1 both below the
    threshold, then skip `min_chunk_size` frames."""
    if min_chunk_size is None:
        min_chunk_size = torbi_b200.MIN_CHUNK_SIZE
    if entropy_threshold is None:
        entropy_threshold = torbi_b200.ENTROPY_THRESHOLD
    frames = observation.shape[0]
    low = entropy(observation.T) < entropy_threshold
    # positions where this frame and the previous one are both low-entropy
    both = (low[1:] & low[:-1]).nonzero().flatten() + 1
    candidates = both.tolist()
    points = []
    i = min_chunk_size
    while i < frames:
        k = bisect.bisect_left(candidates, i)
        if k == len(candidates):
            break
        points.append(candidates[k])
        i = candidates[k] + min_chunk_size
    return points


def entropy(observation: torch.Tensor) -> torch.Tensor:
    """Framewise entropy of (states, frames) log-distributions, normalised by
    log(states) (chunk.py:81-85)"""
    return -(
        (torch.exp(observation) * observation).sum(dim=0) /
        torch.log(torch.tensor(observation.shape[0])))
