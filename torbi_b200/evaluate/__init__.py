"""Evaluation of pitch posteriorgram decoding: the runner of
/root/reference/torbi/evaluate/core.py:14-123 (decode a corpus of posteriorgram files with the pitch
transition matrix, report raw pitch accuracy against a reference decode and the real-time factor) and
its metrics (evaluate/metrics.py).  The reference needs `penn`, `librosa` and downloaded datasets,
none of which exist offline; here the corpus is synthetic (evaluate.synthetic_dataset) and the
comparison is the reference's COMPARE_WITH_REFERENCE = False branch (against the un-chunked decode).

    python -m torbi_b200.evaluate --datasets synthetic --gpu 0 [--config chunked.py]
"""
from .metrics import *   # noqa: F401,F403
from .metrics import Metrics, RPA, agreement, pitch_transition   # noqa: F401
from .core import datasets, synthetic_dataset, frames_to_seconds   # noqa: F401
