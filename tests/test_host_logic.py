"""CPU: host-side logic of the drop-in (no GPU compute): library loading and
exported symbols, argument validation, sharding, collate / separate, chunking,
CLI parsing, configuration."""
import ctypes
import pathlib
import re

import numpy as np
import pytest
import torch

import cases
import torbi_b200
from torbi_b200 import _ffi

ROOT = pathlib.Path(__file__).resolve().parent.parent


def test_library_exports_every_declared_symbol():
    header = (ROOT / 'include' / 'torbi_b200.h').read_text()
    declared = re.findall(r'TORBI_B200_API\s+[\w\s\*]+?(torbi_b200_\w+)\s*\(', header)
    assert len(declared) >= 11
    lib = ctypes.CDLL(str(torbi_b200.LIBRARY))
    for name in declared:
        assert hasattr(lib, name), name
    # and the Python binding covers all of them
    assert set(declared) == set(_ffi.SIGNATURES)
    assert _ffi.library().torbi_b200_abi_version() == _ffi.ABI_VERSION


def test_flag_values_match_header():
    header = (ROOT / 'include' / 'torbi_b200.h').read_text()
    for name in ('OBS_IS_PROBABILITY', 'EPS_FLOOR', 'FORCE_GENERIC', 'FORCE_TILED', 'FORCE_WIDE', 'TIME_KERNELS',
                 'BANDED', 'BACKPOINTERS', 'FORCE_SPLIT', 'SPLIT_STREAMED'):
        shift = int(re.search(rf'#define TORBI_B200_{name} \(1u << (\d+)\)', header).group(1))
        assert getattr(_ffi, name) == 1 << shift
    for name, value in (('TILE_ROWS_16', 1), ('TILE_ROWS_32', 2), ('TILE_ROWS_64', 3)):
        assert re.search(rf'#define TORBI_B200_{name} \({value}u << 12\)', header)
        assert getattr(_ffi, name) == value << 12
    for name in ('INVALID_ARGUMENT', 'WORKSPACE', 'CUDA', 'OUT_OF_MEMORY'):
        code = int(re.search(rf'#define TORBI_B200_ERR_{name} \((-\d+)\)', header).group(1))
        assert getattr(_ffi, f'ERR_{name}') == code


def test_workspace_bytes_and_shape_validation():
    lib = _ffi.library()
    # the fp32 score history dominates (B*T*S*4), or uint16 backpointers (B*T*S*2) in backpointer mode
    assert lib.torbi_b200_workspace_bytes(512, 2000, 1440, 0) >= 512 * 2000 * 1440 * 4
    assert lib.torbi_b200_workspace_bytes(512, 2000, 1440, 0) < 512 * 2000 * 1440 * 4 * 1.01
    assert lib.torbi_b200_workspace_bytes(512, 2000, 1440, _ffi.BACKPOINTERS) >= 512 * 2000 * 1440 * 2
    assert lib.torbi_b200_workspace_bytes(512, 2000, 1440, _ffi.BACKPOINTERS) < 512 * 2000 * 1440 * 2 * 1.02
    assert lib.torbi_b200_workspace_bytes(0, 10, 10, 0) == 0
    assert lib.torbi_b200_workspace_bytes(1, 10, 65536, 0) == 0
    assert 'exceeds 65535' in _ffi.last_error()


def plan(frames, T, S, flags=0, sm_count=148):
    frames = np.ascontiguousarray(frames, dtype=np.int32)
    name = _ffi.library().torbi_b200_plan(frames.ctypes.data, len(frames), T, S, flags, sm_count)
    return name.decode() if name else None


def test_kernel_plan_host_logic():
    """Kernel selection is pure host logic (measured cost models): checked without a GPU."""
    full = [2000] * 512
    assert plan(full, 2000, 1440) == 'forward_tiled'                       # the headline shape: 64-row tiles
    assert plan([2000] * 256, 2000, 1440) == 'forward_tiled_32'            # 8 groups need 8 batches
    assert plan([2000] * 128, 2000, 1440) == 'forward_split'               # too few for 64 x 80 tiles: split-K 32 x 20
    assert plan([2000] * 64, 2000, 1440) == 'forward_split'                # the per-GPU share of 512 over 8 GPUs
    assert plan([2000] * 128, 2000, 1440, _ffi.BACKPOINTERS) == 'forward_tiled_16'   # (no split kernel in that mode)
    assert plan([2000], 2000, 1440) == 'forward_flow'                      # config 1
    assert plan([1000000], 1000000, 4096) == 'forward_wide'                # config 5
    assert plan([2000], 2000, 4096, _ffi.FORCE_FLOW) == 'forward_flow'
    assert plan([2000] * 8, 2000, 1440) == 'forward_wide'                  # a few sequences in lockstep
    assert plan([2000] * 32, 2000, 1440) == 'forward_split'
    assert plan(full, 2000, 1440, _ffi.FORCE_SPLIT) == 'forward_split'
    assert plan(full, 2000, 1440, _ffi.FORCE_SPLIT | _ffi.BACKPOINTERS) is None
    assert plan(full, 2000, 1440, _ffi.BANDED) == 'forward_tiled_banded'
    assert plan(full, 2000, 1440, _ffi.FORCE_GENERIC) == 'forward_generic'
    assert plan(full, 2000, 1440, _ffi.TILE_ROWS_16) == 'forward_tiled_16'
    # ragged, heavy tailed: smaller tiles shorten the critical path of the longest items
    rng = np.random.default_rng(1234)
    ragged = np.exp(rng.uniform(np.log(200), np.log(20000), size=256)).round().astype(np.int32)
    assert plan(ragged, 20000, 1440) == 'forward_split'
    assert plan(ragged, 20000, 1440, _ffi.BACKPOINTERS) == 'forward_tiled_16'
    assert plan(ragged, 20000, 1440, _ffi.BANDED) == 'forward_tiled_banded_16'
    # more destination tiles than SMs: no tiled kernel; 20000 states still fit the wide one
    assert plan([4, 4], 4, 20000) == 'forward_wide'
    assert plan([4, 4], 4, 20000, _ffi.FORCE_TILED) is None and 'tiled kernel' in _ffi.last_error()
    assert plan([3], 3, 65535) == 'forward_generic'
    assert plan([3], 3, 70000) is None


def test_no_cpu_fallback():
    observation = torch.rand(1, 3, 3)
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        torbi_b200.from_probabilities(observation)
    with pytest.raises(RuntimeError, match='CUDA'):
        torbi_b200.decode(observation, torch.tensor([3]), torch.rand(3, 3), torch.rand(3))
    with pytest.raises(RuntimeError, match='MPS'):
        torbi_b200.from_probabilities(observation, gpu='mps')
    with pytest.raises(ValueError):
        torbi_b200.from_probabilities(torch.rand(3, 3), gpu=0)      # 2-D, like the reference


def test_decode_validates_before_touching_the_gpu():
    obs = torch.rand(2, 4, 5)
    with pytest.raises(ValueError, match='transition'):
        torbi_b200.decode(obs, torch.tensor([4, 4]), torch.rand(4, 5), torch.rand(5))
    with pytest.raises(ValueError, match='float32'):
        torbi_b200.decode(obs.double(), torch.tensor([4, 4]), torch.rand(5, 5), torch.rand(5))
    # host entry: frames outside [1, T] rejected by the C side before any CUDA call
    with pytest.raises(ValueError, match='batch_frames'):
        torbi_b200.decode_host(obs, torch.tensor([4, 0]), torch.rand(5, 5), torch.rand(5), gpu=0)
    with pytest.raises(ValueError, match='batch_frames'):
        torbi_b200.decode_host(obs, torch.tensor([5, 1]), torch.rand(5, 5), torch.rand(5), gpu=[0, 1])


def test_lpt_shards():
    costs = [5, 100, 7, 90, 30, 30, 1, 60]
    shards = torbi_b200.lpt_shards(costs, 3)
    assert sorted(sum(shards, [])) == list(range(len(costs)))
    loads = [sum(costs[i] for i in s) for s in shards]
    assert max(loads) <= 1.25 * sum(costs) / 3
    assert all(s == sorted(s) for s in shards)
    assert torbi_b200.lpt_shards([3, 1], 8) == [[0], [1]]


def test_size_bucketed_batches():
    sizes = [10, 500, 20, 400, 30, 30, 5, 300]
    batches = torbi_b200.size_bucketed_batches(sizes, batch_size=3, max_padded_bytes=1000)
    assert sorted(sum(batches, [])) == list(range(len(sizes)))
    assert all(len(b) <= 3 for b in batches)
    assert all(len(b) * sizes[b[0]] <= 1000 or len(b) == 1 for b in batches)
    assert batches[0] == [1, 3]                      # 3 x 500 would exceed the padded budget
    flat = sum(batches, [])
    assert [sizes[i] for i in flat] == sorted(sizes, reverse=True)


def test_collate_and_separate():
    a = torch.arange(12, dtype=torch.float32).reshape(4, 3)
    b = torch.arange(6, dtype=torch.float32).reshape(2, 3) + 100
    observation, frames, chunks, files = torbi_b200.data.collate([(a, 'a.pt'), (b, 'b.pt')])
    assert observation.shape == (2, 4, 3) and frames.tolist() == [4, 2] and chunks == [1, 1]
    assert files == ('a.pt', 'b.pt')
    assert torch.equal(observation[1, :2], b) and (observation[1, 2:] == 0).all()
    # chunked files: one batch item per chunk, re-joined by separate()
    observation, frames, chunks, _ = torbi_b200.data.collate(
        [([a[:1], a[1:]], 'a.pt'), ([b], 'b.pt')])
    assert observation.shape == (3, 3, 3) and frames.tolist() == [1, 3, 2] and chunks == [2, 1]
    indices = torch.tensor([[7, 0, 0], [1, 2, 3], [4, 5, 0]])
    joined = torbi_b200.data.separate(indices, chunks, frames)
    assert [j.tolist() for j in joined] == [[7, 1, 2, 3], [4, 5]]
    with pytest.raises(ValueError):
        torbi_b200.data.collate([])


@pytest.mark.parametrize('name', list(cases.CHUNK_CASES))
def test_chunk_matches_reference_golden(name, golden):
    """Split points, framewise entropy and chunk views against what the reference's OWN
    torbi.chunk.split / entropy / chunk returned (tests/golden/make_golden.py imports the
    reference package; /root/reference/torbi/chunk.py:12-85)."""
    import importlib
    chunk_module = importlib.import_module('torbi_b200.chunk')   # (the package re-exports the function over the module)
    observation, settings = cases.CHUNK_CASES[name]()
    assert cases.chunk_digest(observation, settings) == str(golden[f'{name}/digest'])
    x = torch.from_numpy(observation)
    entropy = chunk_module.entropy(x.T).numpy()
    assert np.array_equal(entropy.view(np.int32), golden[f'{name}/entropy'].view(np.int32))
    for k, (min_chunk, threshold) in enumerate(settings):
        expected = golden[f'{name}/split{k}'].tolist()
        assert chunk_module.split(x, min_chunk, threshold) == expected
        chunks = torbi_b200.chunk(x, min_chunk, threshold)
        assert [int(c.shape[0]) for c in chunks] == np.diff([0] + expected + [x.shape[0]]).tolist()
        assert torch.equal(torch.cat(chunks), x)
        assert all(c.data_ptr() == x[b].data_ptr() for c, b in zip(chunks, [0] + expected))   # views


def test_packed_batch_is_cut_into_the_reference_chunks(monkeypatch):
    """Host logic of the packed route with MIN_CHUNK_SIZE: the files of a batch lie back to back in one buffer, the
    low-entropy decision is taken for all frames at once, and every FILE is cut exactly where the reference's
    chunk() would cut it alone (chunk.py:12-49) -- never across a file boundary.  The device pieces (entropy
    kernel, decode) are stood in for on the CPU here; the GPU tests cover them."""
    import importlib
    from torbi_b200 import core
    chunk_module = importlib.import_module('torbi_b200.chunk')
    rng = np.random.default_rng(5)
    files = []
    for frames in (120, 7, 64, 1, 90):
        logits = rng.standard_normal((frames, 24))
        logits[rng.random(frames) < 0.6] *= 12.0
        files.append(torch.from_numpy(cases.log_softmax(logits)))
    frames = torch.tensor([f.shape[0] for f in files], dtype=torch.int32)
    item_row = torch.zeros(len(files), dtype=torch.int64)
    item_row[1:] = torch.cumsum(frames.to(torch.int64), 0)[:-1]
    packed = torch.cat(files)
    seen = {}

    def fake_decode_packed(observation, rows, lengths, transition, initial, *, flags=0, **kwargs):
        seen['rows'], seen['frames'], seen['flags'] = rows.tolist(), lengths.tolist(), flags
        return torch.zeros(observation.shape[0], dtype=torch.int32)
    monkeypatch.setattr(core, 'decode_packed', fake_decode_packed)
    monkeypatch.setattr(chunk_module, 'low_entropy_frames_device',
                        lambda observation, threshold, stats=None: chunk_module.entropy(observation.T) < threshold)
    stats = {}
    out = core._from_probabilities_packed(packed, item_row, frames, None, None, True, torch.device('cpu'),
                                          min_chunk_size=8, entropy_threshold=0.5, stats=stats)
    assert out.shape == (packed.shape[0],)
    want_rows, want_frames = [], []
    for row0, observation in zip(item_row.tolist(), files):
        bounds = [0] + chunk_module.split(observation, 8, 0.5) + [observation.shape[0]]
        want_rows += [row0 + b for b in bounds[:-1]]
        want_frames += list(np.diff(bounds))
    assert seen['rows'] == want_rows and seen['frames'] == want_frames
    assert len(want_rows) > len(files) and stats['pieces'] == len(want_rows)
    assert seen['flags'] & _ffi.EPS_FLOOR and not seen['flags'] & _ffi.OBS_IS_PROBABILITY
    # without chunking the items go through unchanged; probabilities are flagged for the fused log
    core._from_probabilities_packed(packed.exp(), item_row, frames, None, None, False, torch.device('cpu'))
    assert seen['rows'] == item_row.tolist() and seen['frames'] == frames.tolist()
    assert seen['flags'] & _ffi.OBS_IS_PROBABILITY


def test_evaluation_runner_host_logic(tmp_path, monkeypatch):
    """evaluate.datasets (evaluate/core.py:14-123) with the decode stubbed out: corpus synthesis,
    partition bookkeeping, the un-chunked reference decode, RPA and real-time-factor reporting."""
    import json
    calls = []

    def fake_from_files_to_files(input_files, output_files, transition_file=None, initial_file=None,
                                 log_probs=False, gpu=None, num_threads=None):
        assert log_probs and torch.load(transition_file).shape == (1440, 1440)
        calls.append(torbi_b200.MIN_CHUNK_SIZE)
        with torbi_b200.timing.context('torbi'):
            for source, target in zip(input_files, output_files):
                bins = torch.load(source).argmax(dim=-1).to(torch.int32)
                if torbi_b200.MIN_CHUNK_SIZE is not None:
                    bins[::7] += 1                                  # the "chunked" decode differs a little
                torch.save(bins, target)
    monkeypatch.setattr(torbi_b200, 'from_files_to_files', fake_from_files_to_files)
    stems = torbi_b200.evaluate.synthetic_dataset('tiny', tmp_path, files=3, states=1440, median_frames=60)
    assert json.loads((tmp_path / 'assets' / 'partitions' / 'tiny.json').read_text()) == stems
    x = torch.load(tmp_path / 'data' / 'cache' / 'tiny' / f'{stems[0]}.pt')
    assert x.dim() == 2 and x.shape[1] == 1440 and torch.allclose(x.exp().sum(-1), torch.ones(x.shape[0]), atol=1e-4)
    result = torbi_b200.evaluate.datasets(['tiny'], gpu=0, root=tmp_path)
    assert calls == [None] and result['tiny']['rpa'] == {'0': 1.0, '1': 1.0, '2': 1.0}      # CONFIG = torbi: against itself
    assert set(result['tiny']['rtf']) == {'torbi', 'end_to_end'} and result['tiny']['files'] == 3
    assert result['tiny']['rpa_vs_truth']['2'] > 0.5
    monkeypatch.setattr(torbi_b200, 'CONFIG', 'chunked')
    monkeypatch.setattr(torbi_b200, 'MIN_CHUNK_SIZE', 32)
    result = torbi_b200.evaluate.datasets(['tiny'], gpu=0, root=tmp_path)
    assert calls == [None, 32]                                                                  # reference outputs reused
    assert 0.8 < result['tiny']['rpa']['0'] < 0.9 and result['tiny']['rpa']['1'] == 1.0
    assert json.loads((tmp_path / 'eval' / 'chunked.json').read_text())['tiny']['frames'] == result['tiny']['frames']
    from torbi_b200.evaluate.__main__ import parse_args
    assert parse_args(['--datasets', 'a', 'b', '--gpu', '0'])['datasets'] == ['a', 'b']


def test_cli_and_configure(tmp_path):
    from torbi_b200.__main__ import parse_args
    args = parse_args(['--input_files', 'a.pt', 'b.pt', '--output_files', 'c.pt', 'd.pt',
                       '--log_probs', '--gpu', '0', '--unknown_flag', '1'])
    assert args['gpu'] == 0 and args['log_probs'] and len(args['input_files']) == 2
    assert set(args) == {'input_files', 'output_files', 'transition_file', 'initial_file',
                         'log_probs', 'gpu', 'num_threads'}
    args = parse_args(['--input_files', 'a.pt', '--output_files', 'c.pt', '--gpu', '0', '1'])
    assert args['gpu'] == [0, 1]
    config = tmp_path / 'config.py'
    config.write_text('BATCH_SIZE = 64\nMIN_CHUNK_SIZE = 32\nNOT_A_CONSTANT = 1\n')
    try:
        torbi_b200.configure(config)
        assert torbi_b200.BATCH_SIZE == 64 and torbi_b200.MIN_CHUNK_SIZE == 32
        assert not hasattr(torbi_b200, 'NOT_A_CONSTANT')
        loader = torbi_b200.data.loader(['x.pt'])
        assert loader.batch_size == 64
    finally:
        torbi_b200.BATCH_SIZE, torbi_b200.MIN_CHUNK_SIZE = 512, None


def test_dispatcher_op_schema_and_meta_kernel():
    # same schema as the reference operator (ops.cpp:16-18), traceable through the meta kernel
    op = torch.ops.torbi_b200.viterbi_decode
    schema = str(op.default._schema)
    assert 'Tensor observation, Tensor batch_frames, Tensor transition, Tensor initial' in schema
    out = op(torch.empty(3, 7, 5, device='meta'), torch.empty(3, device='meta', dtype=torch.int32),
             torch.empty(5, 5, device='meta'), torch.empty(5, device='meta'))
    assert out.shape == (3, 7) and out.dtype == torch.int32
    with pytest.raises((NotImplementedError, RuntimeError)):
        op(torch.rand(1, 3, 3), torch.tensor([3], dtype=torch.int32), torch.rand(3, 3), torch.rand(3))


def test_evaluate_transition_and_rpa():
    import cases
    A = torbi_b200.evaluate.pitch_transition(96)
    assert torch.allclose(A.sum(dim=1), torch.ones(96))
    assert (A[0, 90:] == 0).all() and A[40, 40] == A[40].max()
    # same recipe as the fixture generator used for the golden vectors
    width = torbi_b200.evaluate.MAX_OCTAVES_PER_SECOND * torbi_b200.evaluate.HOPSIZE / \
        torbi_b200.evaluate.SAMPLE_RATE * (torbi_b200.evaluate.OCTAVE / torbi_b200.evaluate.CENTS_PER_BIN) + 1
    assert np.allclose(A.numpy(), cases.triangular_transition(96, width), atol=1e-7)
    result = torbi_b200.evaluate.agreement(
        [torch.tensor([1, 2, 3, 10])], [torch.tensor([1, 3, 5, 10])])
    assert result == {'rpa': {'0': 0.5, '1': 0.75, '2': 1.0}}


def test_banded_hint_detects_the_pitch_transition():
    import cases
    from torbi_b200.core import _is_banded
    assert _is_banded(torch.from_numpy(cases.log_triangular_transition(1440)))
    assert not _is_banded(torch.log_softmax(torch.randn(1440, 1440), dim=-1))
    assert not _is_banded(torch.from_numpy(cases.log_triangular_transition(96, 6.5)))   # too small to matter
    dense_default = torch.full((512, 512), -6.2)
    assert _is_banded(dense_default)      # all-constant: the window is empty, the fast path is exact and cheapest


def test_product_never_imports_the_oracle():
    # the oracle is test infrastructure: nothing under torbi_b200/ may reference it
    for path in (ROOT / 'torbi_b200').rglob('*'):
        if path.suffix in ('.py', '.cu', '.cuh', '.h'):
            text = path.read_text()
            assert 'oracle' not in text.lower(), path


def test_native_file_reader_and_writer(tmp_path):
    """csrc/ptfile.cpp against torch.save / torch.load: plain files through the fast path, views,
    other dtypes, legacy and non-tensor files through the per-file fallback; written files load
    back with torch.load (weights_only and mmap included)."""
    from torbi_b200.data import native
    rng = torch.Generator().manual_seed(3)
    lengths = [7, 1, 12, 5]
    tensors = [torch.randn((f, 24), generator=rng) for f in lengths]
    files = []
    for i, x in enumerate(tensors):
        files.append(tmp_path / f'{i}.pt')
        torch.save(x, files[-1])
    infos = native.inspect(files)
    assert [info.status for info in infos] == [0] * 4
    assert [tuple(info.shape[:info.ndim]) for info in infos] == [(f, 24) for f in lengths]
    assert all(info.dtype == _ffi.PT_F32 and info.nbytes == f * 24 * 4 for info, f in zip(infos, lengths))

    observation, frames = native.load_batch(files, threads=3)
    assert frames.tolist() == lengths and observation.shape == (4, 12, 24)
    for i, x in enumerate(tensors):
        assert torch.equal(observation[i, :lengths[i]], x)
        assert not observation[i, lengths[i]:].any()                       # zero padding like collate()
    # identical to the reference route (Dataset + collate)
    want = torbi_b200.data.collate([(torch.load(f), f) for f in files])[0]
    assert torch.equal(observation, want)
    # into a caller-provided flat buffer
    flat = torch.full((4 * 12 * 24 + 5,), 7.)
    again, _ = native.load_batch(files, out=flat)
    assert again.data_ptr() == flat.data_ptr() and torch.equal(again, want)

    # files outside the fast path fall back one by one: a view with an offset is still native,
    # transposed / legacy / wrapped objects are not
    base = torch.randn((10, 24), generator=rng)
    torch.save(base[2:7], tmp_path / 'view.pt')
    torch.save(base.t().contiguous().t(), tmp_path / 'strided.pt')
    torch.save(base, tmp_path / 'legacy.pt', _use_new_zipfile_serialization=False)
    torch.save({'x': base}, tmp_path / 'dict.pt')
    status = [info.status for info in native.inspect(
        [tmp_path / 'view.pt', tmp_path / 'strided.pt', tmp_path / 'legacy.pt', tmp_path / 'dict.pt',
         tmp_path / 'missing.pt'])]
    assert status == [0, 3, 2, 3, 1]
    mixed, frames = native.load_batch([tmp_path / 'view.pt', tmp_path / 'strided.pt', tmp_path / 'legacy.pt'])
    assert frames.tolist() == [5, 10, 10]
    assert torch.equal(mixed[0, :5], base[2:7]) and torch.equal(mixed[1], base) and torch.equal(mixed[2], base)
    # other dtypes / ranks: the caller gets None and uses the generic collate
    torch.save(base.half(), tmp_path / 'half.pt')
    assert native.inspect([tmp_path / 'half.pt'])[0].dtype == _ffi.PT_F16
    assert native.load_batch([files[0], tmp_path / 'half.pt']) is None

    outputs = [torch.randint(0, 1440, (n,), dtype=torch.int32, generator=rng) for n in (500, 1, 0, 70000)]
    outputs.append(torch.randint(0, 9, (1, 77), dtype=torch.int32, generator=rng))
    paths = [tmp_path / f'out{i}.pt' for i in range(len(outputs))]
    native.save_int32(outputs, paths, threads=2)
    for tensor, path in zip(outputs, paths):
        for kwargs in ({}, {'weights_only': True}, {'weights_only': True, 'mmap': True}):
            back = torch.load(path, **kwargs)
            assert back.dtype == torch.int32 and back.shape == tensor.shape and torch.equal(back, tensor)
    with pytest.raises(ValueError):
        native.save_int32([torch.zeros(3)], [tmp_path / 'bad.pt'])
    with pytest.raises(OSError):
        native.save_int32([outputs[0]], [tmp_path / 'no_such_dir' / 'x.pt'])


def test_native_reader_rejects_damaged_files(tmp_path):
    """Truncated, corrupted or empty containers are reported per file, never read out of bounds."""
    from torbi_b200.data import native
    good = tmp_path / 'good.pt'
    torch.save(torch.arange(24, dtype=torch.float32).reshape(4, 6), good)
    raw = good.read_bytes()
    cases_ = {
        'empty.pt': b'',
        'short.pt': raw[:10],
        'truncated.pt': raw[:len(raw) // 2],
        'no_directory.pt': raw[:-30],
        'garbage.pt': bytes(range(256)) * 8,
    }
    # a central directory that points past the end of the file
    bad = bytearray(raw)
    eocd = raw.rfind(b'PK\x05\x06')
    bad[eocd + 16:eocd + 20] = (len(raw) + 1000).to_bytes(4, 'little')
    cases_['bad_offset.pt'] = bytes(bad)
    # a pickle that is not a tensor record
    flipped = bytearray(raw)
    at = raw.find(b'_rebuild_tensor_v2')
    flipped[at:at + 8] = b'_REBUILD'
    cases_['other_pickle.pt'] = bytes(flipped)
    for name, data in cases_.items():
        (tmp_path / name).write_bytes(data)
    infos = native.inspect([tmp_path / name for name in cases_] + [good])
    assert all(info.status != 0 for info in infos[:-1]), [info.status for info in infos]
    assert infos[-1].status == 0


def test_planned_workspace_follows_the_kernel_choice():
    """torbi_b200_workspace_bytes_planned sizes the trailing region for the kernel the call will
    use: uint16 backpointers (2 bytes per entry) for the wide / generic kernels, fp32 score history
    (4 bytes) for the tiled / split / cluster kernels.  Config 5 needs 8.3 GB, not 16.5 GB."""
    lib = _ffi.library()

    def planned(frames, T, S, flags=0, sm_count=148):
        frames = np.ascontiguousarray(frames, dtype=np.int32)
        return lib.torbi_b200_workspace_bytes_planned(frames.ctypes.data, len(frames), T, S, flags, sm_count)

    upper = lib.torbi_b200_workspace_bytes(1, 1_000_000, 4096, 0)
    exact = planned([1_000_000], 1_000_000, 4096)
    assert plan([1_000_000], 1_000_000, 4096) == 'forward_wide'
    assert upper - exact == 2 * 1_000_000 * 4096 and exact < 8.3e9
    assert planned([2000] * 512, 2000, 1440) == lib.torbi_b200_workspace_bytes(512, 2000, 1440, 0)
    assert planned([2000] * 64, 2000, 1440, _ffi.FORCE_CLUSTER) == lib.torbi_b200_workspace_bytes(64, 2000, 1440, 0)
    assert planned([4, 4], 4, 20000) == lib.torbi_b200_workspace_bytes(2, 4, 20000, _ffi.BACKPOINTERS)
    assert planned([4, 9], 4, 100) == 0 and 'outside' in _ffi.last_error()
    assert planned([4], 4, 100, _ffi.OBS_F16 | _ffi.OBS_BF16) == 0


def test_cluster_kernel_plan():
    assert plan([2000] * 64, 2000, 1440, _ffi.FORCE_CLUSTER) == 'forward_cluster'
    assert plan([2000] * 512, 2000, 1480, _ffi.FORCE_CLUSTER) == 'forward_cluster'      # 74 pairs = 148 SMs
    assert plan([2000] * 64, 2000, 1481, _ffi.FORCE_CLUSTER) is None                     # 75 pairs do not fit
    assert plan([2000] * 64, 2000, 1440, _ffi.FORCE_CLUSTER | _ffi.BACKPOINTERS) is None
    assert plan([2000] * 64, 2000, 1440, _ffi.FORCE_CLUSTER | _ffi.BANDED) is None
    # the banded flag is a hint: dropped where the tiled kernel cannot run
    assert plan([3] * 32, 3, 12000, _ffi.BANDED) == plan([3] * 32, 3, 12000)
    assert plan([3] * 32, 3, 12000, _ffi.BANDED | _ffi.FORCE_TILED) is None


def test_torbi_alias_package():
    """`import torbi` is torbi_b200 (the reference's call sites run unchanged,
    /root/reference/torbi/__init__.py:72-78)."""
    import importlib
    import sys
    saved = {name: module for name, module in sys.modules.items() if name == 'torbi' or name.startswith('torbi.')}
    for name in saved:
        del sys.modules[name]
    try:
        torbi = importlib.import_module('torbi')
        assert torbi is torbi_b200
        for name in ('from_probabilities', 'from_file', 'from_file_to_file', 'from_files_to_files', 'from_dataloader',
                     'decode', 'chunk', 'data', 'BATCH_SIZE', 'MIN_CHUNK_SIZE', 'ENTROPY_THRESHOLD'):
            assert hasattr(torbi, name), name
        assert importlib.import_module('torbi.core') is importlib.import_module('torbi_b200.core')
        assert importlib.import_module('torbi.data').loader is torbi_b200.data.loader
    finally:
        for name in [n for n in sys.modules if n == 'torbi' or n.startswith('torbi.')]:
            del sys.modules[name]
        sys.modules.update(saved)


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
def test_native_reader_keeps_half_precision_files(tmp_path, dtype):
    """fp16 / bf16 posteriorgram files are read natively into a batch of their own dtype (2 bytes per
    element up to the kernel's observation load); mixed dtypes go to the generic collate."""
    from torbi_b200.data import native
    g = torch.Generator().manual_seed(3)
    tensors = [torch.randn(n, 12, generator=g).to(dtype) for n in (5, 9, 1)]
    files = []
    for i, tensor in enumerate(tensors):
        files.append(tmp_path / f'{i}.pt')
        torch.save(tensor, files[-1])
    pinned_like = torch.empty(1000, dtype=torch.float32)
    for out in (None, pinned_like):
        observation, frames = native.load_batch(files, threads=2, out=out)
        assert observation.dtype == dtype and observation.shape == (3, 9, 12) and frames.tolist() == [5, 9, 1]
        for i, tensor in enumerate(tensors):
            assert torch.equal(observation[i, :tensor.shape[0]], tensor)
            assert (observation[i, tensor.shape[0]:] == 0).all()
    torch.save(tensors[0].to(torch.float64), files[0])
    assert native.load_batch(files, threads=2) is None


@pytest.mark.parametrize('dtype', [torch.float32, torch.bfloat16])
def test_native_reader_packs_a_batch_without_padding(tmp_path, dtype):
    from torbi_b200.data import native
    g = torch.Generator().manual_seed(4)
    tensors = [torch.randn(n, 12, generator=g).to(dtype) for n in (5, 9, 1, 9)]
    files = []
    for i, tensor in enumerate(tensors):
        files.append(tmp_path / f'{i}.pt')
        torch.save(tensor, files[-1])
    for out in (None, torch.full((1000,), float('nan'))):
        observation, item_row, frames = native.load_batch_packed(files, threads=2, out=out)
        assert observation.dtype == dtype and observation.shape == (24, 12)
        assert item_row.tolist() == [0, 5, 14, 15] and frames.tolist() == [5, 9, 1, 9]
        assert torch.equal(observation, torch.cat(tensors))
    torch.save(torch.randn(3, 11), files[1])                       # another width: generic path
    assert native.load_batch_packed(files, threads=2) is None
    assert torbi_b200.core.size_capped_batches([10, 50, 20, 40, 30], 3, max_bytes=90) == [[1, 3], [4, 2, 0]]
    assert torbi_b200.core.size_capped_batches([500, 5], 8, max_bytes=90) == [[0], [1]]


def test_file_pipeline_stops_its_loader_when_the_decode_fails(tmp_path, monkeypatch):
    """A failing decode must not leave the loader thread blocked on its queue holding pinned buffers."""
    import threading
    from torbi_b200 import core
    files = []
    for i in range(12):
        files.append(str(tmp_path / f'{i}.pt'))
        torch.save(torch.rand(4, 6), files[-1])
    monkeypatch.setattr(torbi_b200, 'BATCH_SIZE', 2)

    def boom(**kwargs):
        raise RuntimeError('decode failed')
    monkeypatch.setattr(core, 'from_probabilities', boom)
    monkeypatch.setattr(core, '_from_probabilities_packed', lambda *args, **kwargs: boom())
    before = threading.active_count()
    with pytest.raises(RuntimeError, match='decode failed'):
        core._from_files_native(files, [100] * len(files), {f: f + '.out' for f in files}, None, None, True, 0,
                                False, 2)
    assert threading.active_count() == before

    def failing_loader():
        yield torch.zeros(1, 2, 3), torch.tensor([2]), [1], ['a']
        yield torch.zeros(1, 2, 3), torch.tensor([2]), [1], ['b']
        yield torch.zeros(1, 2, 3), torch.tensor([2]), [1], ['c']
    iterator = core._prefetched(failing_loader(), 1)
    next(iterator)
    iterator.close()                                               # consumer walks away early
    assert threading.active_count() == before
