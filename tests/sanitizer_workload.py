"""Small decodes through every kernel for compute-sanitizer (memcheck / racecheck /
initcheck / synccheck):  compute-sanitizer --tool memcheck python tests/sanitizer_workload.py
(test infrastructure: the only reason it lives under tests/ is that it checks against the oracle)"""
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))
import cases  # noqa: E402
import torbi_b200  # noqa: E402
from torbi_b200 import _ffi  # noqa: E402
from oracle import oracle  # noqa: E402

rng = np.random.default_rng(0)
for (B, T, S) in [(5, 6, 200), (70, 4, 96), (3, 5, 333)]:
    obs = cases.log_softmax(rng.standard_normal((B, T, S)))
    A = cases.log_softmax(rng.standard_normal((S, S)))
    init = cases.log_softmax(rng.standard_normal((S,)))
    frames = rng.integers(1, T + 1, size=B).astype(np.int32)
    frames[0] = T
    want = oracle.c_decode(obs, frames, A, init)
    for name, flag in [('generic', _ffi.FORCE_GENERIC), ('tiled64', _ffi.TILE_ROWS_64), ('tiled32', _ffi.TILE_ROWS_32),
                       ('tiled16', _ffi.TILE_ROWS_16), ('wide', _ffi.FORCE_WIDE),
                       ('tiled64 backpointers', _ffi.TILE_ROWS_64 | _ffi.BACKPOINTERS),
                       ('tiled16 backpointers', _ffi.TILE_ROWS_16 | _ffi.BACKPOINTERS),
                       ('banded', _ffi.BANDED | _ffi.TILE_ROWS_64), ('banded16', _ffi.BANDED | _ffi.TILE_ROWS_16),
                       ('banded backpointers', _ffi.BANDED | _ffi.BACKPOINTERS), ('split', _ffi.FORCE_SPLIT), ('split streamed', _ffi.FORCE_SPLIT | _ffi.SPLIT_STREAMED),
                       ('cluster', _ffi.FORCE_CLUSTER), ('flow', _ffi.FORCE_FLOW)]:
        got = torbi_b200.decode(torch.from_numpy(obs).cuda(), torch.from_numpy(frames), torch.from_numpy(A).cuda(),
                                torch.from_numpy(init).cuda(), flags=flag | _ffi.EPS_FLOOR)
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy(), want), (name, B, T, S)
        print('ok', name, (B, T, S), flush=True)

# packed layout + device entropy
B, S = 6, 96
lengths = np.array([9, 1, 4, 7, 2, 5], np.int32)
rows = np.concatenate([[0], np.cumsum(lengths)[:-1]]).astype(np.int64)
packed = cases.log_softmax(rng.standard_normal((int(lengths.sum()), S)))
A = cases.log_softmax(rng.standard_normal((S, S)))
init = cases.log_softmax(rng.standard_normal((S,)))
for name, flag in [('packed tiled', _ffi.TILE_ROWS_16), ('packed split', _ffi.FORCE_SPLIT), ('packed flow', _ffi.FORCE_FLOW)]:
    got = torbi_b200.decode_packed(torch.from_numpy(packed).cuda(), torch.from_numpy(rows), torch.from_numpy(lengths),
                                   torch.from_numpy(A).cuda(), torch.from_numpy(init).cuda(), flags=flag).cpu().numpy()
    for b in range(B):
        want = oracle.c_decode(packed[None, rows[b]:rows[b] + lengths[b]], lengths[b:b + 1], A, init)
        assert np.array_equal(got[rows[b]:rows[b] + lengths[b]], want[0]), (name, b)
    print('ok', name, flush=True)
import importlib
chunk_module = importlib.import_module('torbi_b200.chunk')
values = chunk_module.entropy_device(torch.from_numpy(packed).cuda()).cpu()
assert torch.allclose(values[:, 0], chunk_module.entropy(torch.from_numpy(packed).T), atol=1e-5)
print('ok entropy', flush=True)
