"""Generate tests/golden/golden.npz from the UNMODIFIED reference.

Run once in the build container (needs /root/reference and oracle/_ref):

    python tests/golden/make_golden.py

For every case in cases.py it records
  * `<case>/digest`       sha256 of the regenerated inputs (tests verify it)
  * `<case>/api_indices`  torbi.from_probabilities(**case) from the reference's OWN
                          Python package (a scratch copy under /tmp with its compiled
                          CPU op dropped in as torbi/_C.so and two stub modules for
                          the uninstallable yapecs/torchutil) -- end-to-end truth
  * `<case>/prep_sha`     sha256 of the fp32 observation after the reference's
                          log-space conversion (core.py:189-197, torch CPU ops)
  * `<case>/indices`, `<case>/posterior`  reference op + debug_trellis on the
                          converted inputs
  * `<case>/trellis`      int32 backpointers, rows 1..frames-1 of every item packed
                          end to end (rows outside that range are not part of the
                          contract), or `<case>/trellis_sha` when large
Nothing here runs on the GPU box; the .npz is the committed fixture.
"""
import hashlib
import importlib
import pathlib
import shutil
import sys
import tempfile
import textwrap

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(HERE))

import cases  # noqa: E402
from oracle import build_ref, oracle  # noqa: E402

REFERENCE = pathlib.Path('/root/reference')

YAPECS_STUB = '''
import argparse
def configure(name, defaults, *a, **k):
    return None
class ArgumentParser(argparse.ArgumentParser):
    pass
'''

TORCHUTIL_STUB = '''
import contextlib, time as _time
class _Time:
    def __init__(self): self.history = {}
    @contextlib.contextmanager
    def context(self, name):
        start = _time.perf_counter()
        try:
            yield
        finally:
            self.history[name] = self.history.get(name, 0.) + _time.perf_counter() - start
    def reset(self): self.history = {}
    def results(self): return dict(self.history)
time = _Time()
def iterator(iterable, message=None, initial=0, total=None):
    class _Bar:
        def __init__(self, it): self.it = it
        def __iter__(self): return iter(self.it)
        def update(self, n=1): pass
        def close(self): pass
    return _Bar(iterable)
def notify(task):
    def wrap(fn): return fn
    return wrap
class metrics:
    class Average:
        def __init__(self): self.reset()
        def __call__(self): return self.total / max(self.count, 1)
        def update(self, values, count): self.total += values; self.count += count
        def reset(self): self.total = 0.; self.count = 0
class download:
    pass
class paths:
    pass
'''


def import_reference_package():
    """Scratch copy of the reference package made importable (nothing is written
    to /root/reference or to the repo)."""
    if not build_ref.available():
        build_ref.build()
    scratch = pathlib.Path(tempfile.mkdtemp(prefix='torbi_ref_pkg_'))
    shutil.copytree(REFERENCE / 'torbi', scratch / 'torbi',
                    ignore=shutil.ignore_patterns('mps', '__pycache__'))
    # a symlink, not a copy: the same inode dlopen()s once, so the reference's
    # TORCH_LIBRARY block is not registered twice when oracle/_ref is loaded too
    (scratch / 'torbi' / '_C.so').symlink_to(build_ref.REF_SO)
    (scratch / 'yapecs.py').write_text(textwrap.dedent(YAPECS_STUB))
    (scratch / 'torchutil.py').write_text(textwrap.dedent(TORCHUTIL_STUB))
    sys.path.insert(0, str(scratch))
    return importlib.import_module('torbi'), scratch


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def pack_trellis(trellis, frames):
    return np.concatenate([trellis[b, 1:f].reshape(-1) for b, f in enumerate(frames)]
                          + [np.zeros(0, np.int32)])


def tensor(x):
    return None if x is None else torch.from_numpy(np.array(x))


def main():
    torbi_ref, scratch = import_reference_package()
    out = {}
    for name, make in cases.CASES.items():
        case = make()
        out[f'{name}/digest'] = np.array(cases.digest(case))

        # (1) the reference's own Python entry point, CPU path
        api = torbi_ref.from_probabilities(
            observation=tensor(case['observation']).clone(),
            batch_frames=tensor(case['batch_frames']),
            transition=tensor(case['transition']),
            initial=tensor(case['initial']),
            log_probs=case['log_probs'], gpu=None, num_threads=4)
        out[f'{name}/api_indices'] = api.numpy().astype(np.int32)

        # (2) the op and its intermediates on the converted inputs
        obs, frames, A, init = oracle.ref_prepare(
            case['observation'], case['batch_frames'], case['transition'], case['initial'],
            case['log_probs'])
        out[f'{name}/prep_sha'] = np.array(sha(obs.numpy()))
        indices, posterior, trellis = oracle.ref_decode(
            obs.numpy(), frames.numpy(), A.numpy().astype(np.float32),
            init.numpy().astype(np.float32), want_debug=True)
        assert (indices == out[f'{name}/api_indices']).all(), name
        out[f'{name}/indices'] = indices.astype(np.int32)
        out[f'{name}/posterior'] = posterior
        packed = pack_trellis(trellis, frames.numpy())
        if packed.size <= 40000:
            out[f'{name}/trellis'] = packed.astype(np.int32)
        else:
            out[f'{name}/trellis_sha'] = np.array(sha(packed.astype(np.int32)))
        print(f'{name:24s} obs{tuple(obs.shape)} indices[0,:8]={indices[0, :8].tolist()}')

    assert out['kat/indices'].tolist() == [[1, 2, 2]]      # tests/test_core.py:25

    # (3) entropy chunking: the reference's own torbi.chunk.entropy / split / chunk
    # (chunk.py:12-85) on log-probability sequences, several (min_chunk_size, threshold) each
    # (the package re-exports the function `chunk` over its module, torbi/__init__.py:72-78)
    ref_chunk = sys.modules['torbi.chunk']
    for name, make in cases.CHUNK_CASES.items():
        observation, settings = make()
        out[f'{name}/digest'] = np.array(cases.chunk_digest(observation, settings))
        x = torch.from_numpy(observation)
        out[f'{name}/entropy'] = ref_chunk.entropy(x.T).numpy()
        for k, (min_chunk, threshold) in enumerate(settings):
            points = ref_chunk.split(x, min_chunk_size=min_chunk, entropy_threshold=threshold)
            pieces = ref_chunk.chunk(x, min_chunk_size=min_chunk, entropy_threshold=threshold)
            assert [int(p.shape[0]) for p in pieces] == \
                list(np.diff([0] + [int(p) for p in points] + [observation.shape[0]]))
            out[f'{name}/split{k}'] = np.array(points, dtype=np.int64)
            print(f'{name:24s} min_chunk={min_chunk:4d} thr={threshold} -> {len(points)} split points')
    np.savez_compressed(HERE / 'golden.npz', **out)
    shutil.rmtree(scratch, ignore_errors=True)
    print('wrote', HERE / 'golden.npz', (HERE / 'golden.npz').stat().st_size, 'bytes')


if __name__ == '__main__':
    main()
