"""Approximate long-sequence mode: cut a sequence where two adjacent frames have
low normalised entropy, so the pieces decode independently and batch well.
Same split points as /root/reference/torbi/chunk.py:12-85.

CPU tensors: the reference's own torch ops (host logic).  CUDA tensors: the framewise
entropy comes from the device kernel behind torbi_b200_entropy (csrc/entropy.cu, one pass
over HBM) and only the frames whose entropy lies within the rounding margin of the
threshold are re-evaluated with the reference's ops, so the split DECISIONS -- all that
leaves this module -- are the reference's; the pieces are row ranges of the same buffer
and feed torbi_b200.decode_packed without a copy."""
import bisect
import math
from typing import List, Optional

import torch

import torbi_b200
from . import _ffi


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
    if observation.is_cuda:
        low = low_entropy_frames_device(observation, entropy_threshold)
    else:
        low = entropy(observation.T) < entropy_threshold
    return split_points_from_candidates(low, frames, min_chunk_size)


def split_points_from_candidates(low: torch.Tensor, frames: int, min_chunk_size: int) -> List[int]:
    """The scan of chunk.py:66-76 over a (frames,) bool tensor of low-entropy frames."""
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


###############################################################################
# Device route
###############################################################################


def entropy_device(observation: torch.Tensor) -> torch.Tensor:
    """(frames, 2) float32 on the device of the CUDA tensor `observation` (frames, states):
    column 0 the framewise normalised entropy (chunk.py:81-85), column 1 the sum of the terms'
    magnitudes / log(states), which bounds the rounding difference to any other summation order."""
    if observation.dim() != 2 or not observation.is_cuda:
        raise ValueError('entropy_device takes a CUDA tensor of shape (frames, states)')
    flags = {torch.float32: 0, torch.float16: _ffi.OBS_F16, torch.bfloat16: _ffi.OBS_BF16}.get(observation.dtype)
    if flags is None:
        observation, flags = observation.to(torch.float32), 0
    observation = observation.contiguous()
    rows, states = observation.shape
    with torch.cuda.device(observation.device):
        out = torch.empty((rows, 2), dtype=torch.float32, device=observation.device)
        if rows:
            _ffi.check(_ffi.library().torbi_b200_entropy(
                observation.data_ptr(), rows, states, flags, out.data_ptr(),
                torch.cuda.current_stream(observation.device).cuda_stream))
    return out


def rounding_margin(states: int, dtype: torch.dtype) -> float:
    """Largest |H_device - H_reference| per unit of term magnitude.  fp32: two different summation
    orders of `states` fp32 terms differ by at most ~2 states 2^-24 of the terms' magnitude (plus a
    few ulp from exp / the division).  fp16 / bf16 files: the reference evaluates exp, the product and
    the result in the file's own dtype (unit roundoff 2^-11 / 2^-8), the device in fp32 on the exactly
    upcast values, so the reference's own rounding dominates."""
    margin = 4.0 * states * 2.0 ** -24 + 1e-6
    if dtype == torch.float16:
        margin += 8 * 2.0 ** -11
    elif dtype == torch.bfloat16:
        margin += 8 * 2.0 ** -8
    return margin


def low_entropy_frames_device(observation: torch.Tensor, entropy_threshold: float,
                              stats: Optional[dict] = None) -> torch.Tensor:
    """(frames,) bool CPU tensor: entropy(frame) < threshold, as the reference decides it.  The
    device entropy settles every frame that is further from the threshold than the rounding margin;
    the others (a handful) are re-evaluated with the reference's torch ops on the CPU."""
    values = entropy_device(observation).cpu()
    h, magnitude = values[:, 0], values[:, 1]
    margin = rounding_margin(observation.shape[1], observation.dtype) * torch.clamp(magnitude, min=1.0)
    low = h < entropy_threshold
    # NaN / inf entropies (inputs outside the intended domain) are always left to the reference's ops
    unsure = ~((h - entropy_threshold).abs() > margin)
    recheck = unsure.nonzero().flatten()
    if stats is not None:
        stats['frames'] = stats.get('frames', 0) + h.numel()
        stats['rechecked'] = stats.get('rechecked', 0) + recheck.numel()
    if recheck.numel():
        frames = observation.index_select(0, recheck.to(observation.device)).cpu()
        low[recheck] = entropy(frames.T) < entropy_threshold
    return low
