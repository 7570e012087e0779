"""Data loader (restates /root/reference/torbi/data/loader.py:10-25)."""
import torch

import torbi_b200
from .collate import collate
from .dataset import Dataset


def loader(input_files, num_workers=None, collate_fn=collate, batch_sampler=None):
    """Batches of torbi_b200.BATCH_SIZE files in the given order, no shuffling.
    `batch_sampler` (a list of index lists) overrides the fixed-size batching."""
    if num_workers is None:
        num_workers = torbi_b200.NUM_WORKERS
    if batch_sampler is not None:
        return torch.utils.data.DataLoader(
            Dataset(input_files),
            num_workers=num_workers,
            batch_sampler=batch_sampler,
            collate_fn=collate_fn)
    return torch.utils.data.DataLoader(
        Dataset(input_files),
        num_workers=num_workers,
        batch_size=torbi_b200.BATCH_SIZE,
        shuffle=False,
        collate_fn=collate_fn)
