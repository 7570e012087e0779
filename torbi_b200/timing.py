"""Named decode timers: the part of `torchutil.time` the reference's hot path uses
(/root/reference/torbi/core.py:200 wraps torbi.decode in torchutil.time.context('torbi'), and
evaluate/core.py:40,114-118 turns the totals into real-time factors).  The reference reads the host
clock around an un-synchronised CUDA launch; here a context on a CUDA device records two events on the
current stream and results() resolves them, so the totals are device time and nothing synchronises
inside the pipeline."""
import contextlib
import threading
import time
from typing import Dict

import torch

_lock = threading.Lock()
_host: Dict[str, float] = {}
_events: Dict[str, list] = {}


def reset() -> None:
    with _lock:
        _host.clear()
        _events.clear()


@contextlib.contextmanager
def context(name: str, device=None):
    if device is not None and torch.device(device).type == 'cuda':
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        stream = torch.cuda.current_stream(device)
        start.record(stream)
        try:
            yield
        finally:
            end.record(stream)
            with _lock:
                _events.setdefault(name, []).append((start, end))
    else:
        t0 = time.perf_counter()
        try:
            yield
        finally:
            with _lock:
                _host[name] = _host.get(name, 0.0) + time.perf_counter() - t0


def results() -> Dict[str, float]:
    """Seconds per timer name (torchutil.time.results)"""
    with _lock:
        out = dict(_host)
        for name, pairs in _events.items():
            total = 0.0
            for start, end in pairs:
                end.synchronize()
                total += start.elapsed_time(end) * 1e-3
            out[name] = out.get(name, 0.0) + total
    return out
