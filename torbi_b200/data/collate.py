"""Batch assembly (restates /root/reference/torbi/data/collate.py:9-45)."""
import torch


def collate(batch, pin_memory=False):
    """[(observation | [chunks], file), ...] -> zero-padded (B, Tmax, S) batch,
    per-item frame counts, chunks per file, and the files (collate.py:9-33).
    `pin_memory` builds the batch directly in page-locked memory (one host copy
    less than pinning afterwards)."""
    observations, input_files = zip(*batch)

    # a chunked file contributes one batch item per chunk (collate.py:13-17)
    if isinstance(observations[0], list):
        batch_chunks = [len(obs) for obs in observations]
        observations = [piece for obs in observations for piece in obs]
    else:
        batch_chunks = [1] * len(observations)

    if len(observations) == 0:
        raise ValueError('batch must contain at least 1 item')

    batch_frames = torch.tensor([obs.shape[0] for obs in observations])
    max_frames = int(batch_frames.max())

    # pad with zeros in the file dtype; padded frames are never decoded (collate.py:26-31).
    # Only the tails are zero-filled: the rest is overwritten by the copy anyway.
    observation = torch.empty(
        (len(observations), max_frames, observations[0].shape[-1]),
        dtype=observations[0].dtype, pin_memory=pin_memory)
    for i, obs in enumerate(observations):
        observation[i, :obs.shape[0]] = obs
        observation[i, obs.shape[0]:] = 0

    return observation, batch_frames, batch_chunks, input_files


def separate(indices, batch_chunks, batch_frames):
    """Re-join the decoded chunks of each file (collate.py:36-45)."""
    separated = []
    start = 0
    for chunks in batch_chunks:
        separated.append(torch.cat([
            indices[start + i, :int(batch_frames[start + i])] for i in range(chunks)]))
        start += chunks
    return separated
