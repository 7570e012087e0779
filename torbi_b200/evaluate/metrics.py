"""Metrics of /root/reference/torbi/evaluate/metrics.py (raw pitch accuracy at several thresholds) and
the pitch transition matrix recipe of evaluate/core.py:23-34, with penn's constants written out (penn is
not importable offline)."""
from typing import Dict, Iterable, Sequence

import torch

# penn's constants used by evaluate/core.py:27-31 (not importable offline)
PITCH_BINS = 1440
CENTS_PER_BIN = 5.
OCTAVE = 1200.
MAX_OCTAVES_PER_SECOND = 35.92
HOPSIZE = 80
SAMPLE_RATE = 8000

# Thresholds (in number of 5 cent bins) for raw pitch accuracy (config/defaults.py:62)
PITCH_ERROR_THRESHOLDS = [0, 1, 2]


def pitch_transition(bins: int = PITCH_BINS) -> torch.Tensor:
    """Triangular transition PROBABILITIES of evaluate/core.py:23-34: a pitch can
    move at most `max_bins_per_frame` bins per frame, linearly less likely the
    further it moves; rows normalised.  Pass it with log_probs=False, or as
    log(A + tiny) with log_probs=True like from_files_to_files (core.py:343-345)."""
    xx, yy = torch.meshgrid(torch.arange(bins), torch.arange(bins), indexing='ij')
    bins_per_octave = OCTAVE / CENTS_PER_BIN
    max_octaves_per_frame = MAX_OCTAVES_PER_SECOND * HOPSIZE / SAMPLE_RATE
    max_bins_per_frame = max_octaves_per_frame * bins_per_octave + 1
    transition = torch.clip(max_bins_per_frame - (xx - yy).abs(), 0)
    return transition / transition.sum(dim=1, keepdims=True)


class RPA:
    """Raw pitch accuracy within `threshold` bins (evaluate/metrics.py:34-43)."""

    def __init__(self, threshold: int):
        self.threshold = threshold
        self.reset()

    def __call__(self) -> float:
        return self.total / max(self.count, 1)

    def update(self, predicted: torch.Tensor, target: torch.Tensor) -> None:
        difference = (predicted.long() - target.long()).abs()
        self.total += int((difference <= self.threshold).sum())
        self.count += predicted.numel()

    def reset(self) -> None:
        self.total = 0
        self.count = 0


class Metrics:
    """Aggregate of RPA at PITCH_ERROR_THRESHOLDS (evaluate/metrics.py:12-27)."""

    def __init__(self, thresholds: Sequence[int] = tuple(PITCH_ERROR_THRESHOLDS)):
        self.rpas = [RPA(k) for k in thresholds]

    def __call__(self) -> Dict[str, Dict[str, float]]:
        return {'rpa': {str(rpa.threshold): rpa() for rpa in self.rpas}}

    def update(self, predicted: torch.Tensor, target: torch.Tensor) -> None:
        for rpa in self.rpas:
            rpa.update(predicted, target)

    def reset(self) -> None:
        for rpa in self.rpas:
            rpa.reset()


def agreement(predicted: Iterable[torch.Tensor], target: Iterable[torch.Tensor]) -> Dict:
    """RPA of decoded bins against reference bins over a set of sequences, the
    comparison evaluate/core.py:106-112 reports between torbi and librosa."""
    metrics = Metrics()
    for p, t in zip(predicted, target):
        metrics.update(p.cpu(), t.cpu())
    return metrics()
