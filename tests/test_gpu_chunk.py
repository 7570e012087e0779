"""GPU (-m gpu): entropy chunking on the device -- torbi_b200_entropy + the host re-check
of the frames near the threshold reproduce the split points the REFERENCE's torbi.chunk.split returned
(tests/golden/golden.npz), the pieces decode in place through the packed layout, and the file pipeline
with MIN_CHUNK_SIZE gives the files the reference's Dataset / collate / separate route gives."""
import importlib
import os

import numpy as np
import pytest
import torch

import cases
import torbi_b200
from torbi_b200 import _ffi

pytestmark = pytest.mark.gpu

chunk_module = importlib.import_module('torbi_b200.chunk')   # (the package re-exports the function over the module)


@pytest.mark.parametrize('name', list(cases.CHUNK_CASES))
def test_device_split_matches_reference_golden(name, golden):
    observation, settings = cases.CHUNK_CASES[name]()
    x = torch.from_numpy(observation).cuda()
    values = chunk_module.entropy_device(x).cpu().numpy()
    reference = golden[f'{name}/entropy']                                   # the reference's torch CPU values
    margin = chunk_module.rounding_margin(x.shape[1], torch.float32) * np.maximum(values[:, 1], 1.0)
    assert (np.abs(values[:, 0] - reference) <= margin).all(), 'device entropy outside its stated rounding margin'
    assert np.abs(values[:, 0] - reference).max() < 2e-5
    for k, (min_chunk, threshold) in enumerate(settings):
        expected = golden[f'{name}/split{k}'].tolist()
        stats = {}
        low = chunk_module.low_entropy_frames_device(x, threshold, stats)
        assert low.tolist() == (reference < threshold).tolist()
        assert chunk_module.split(x, min_chunk, threshold) == expected
        chunks = torbi_b200.chunk(x, min_chunk, threshold)
        assert [int(c.shape[0]) for c in chunks] == np.diff([0] + expected + [x.shape[0]]).tolist()
        assert all(c.data_ptr() == x[b].data_ptr() for c, b in zip(chunks, [0] + expected))   # views
        if 'hugger' not in name and x.shape[0] >= 400:
            assert stats['rechecked'] <= 0.02 * stats['frames'], stats      # the host re-check is the exception


@pytest.mark.parametrize('dtype', [torch.float16, torch.bfloat16, torch.float32])
@pytest.mark.parametrize('states', [1440, 333])
def test_device_split_equals_the_reference_ops_on_any_dtype(states, dtype):
    """fp16 / bf16 observations: the reference evaluates the entropy in the file's own dtype; the device
    route must make the same decisions (wider re-check band), here against the reference's ops on the CPU."""
    rng = np.random.default_rng(states)
    logits = rng.standard_normal((4000, states))
    logits[rng.random(4000) < 0.5] *= 9.0
    x = torch.from_numpy(cases.log_softmax(logits)).to(dtype)
    for min_chunk, threshold in ((50, 0.5), (7, 0.3), (200, 0.8)):
        want = chunk_module.split(x, min_chunk, threshold)                   # CPU tensors: the reference's own ops
        assert chunk_module.split(x.cuda(), min_chunk, threshold) == want


def test_device_split_peaky_100k_frames():
    """10^5 frames x 1440 states, mostly peaky with noisy stretches: identical split points, a small
    re-check set, no Python loop over the frames."""
    T, S = 100_000, 1440
    g = torch.Generator(device='cuda:0').manual_seed(9)
    logits = torch.randn(T, S, device='cuda:0', generator=g)
    gain = torch.where(torch.rand(T, 1, device='cuda:0', generator=g) < 0.6, 10.0, 1.0)
    x = torch.log_softmax(logits * gain, dim=-1)
    host = x.cpu()
    for min_chunk, threshold in ((500, 0.5), (64, 0.35)):
        want = chunk_module.split(host, min_chunk, threshold)
        stats = {}
        low = chunk_module.low_entropy_frames_device(x, threshold, stats)
        assert low.tolist() == (chunk_module.entropy(host.T) < threshold).tolist()
        assert chunk_module.split(x, min_chunk, threshold) == want
        assert len(want) > 50 and stats['rechecked'] < 0.01 * T, (len(want), stats)


def test_chunked_pieces_decode_in_place(oracle):
    """The pieces of a chunked sequence are row ranges of the same device buffer: decode_packed on them
    equals decoding every chunk on its own (what the reference's Dataset + collate + separate do)."""
    observation, _ = cases.CHUNK_CASES['chunk_s1440']()
    x = torch.from_numpy(observation[:1200]).cuda()
    points = chunk_module.split(x, 64, 0.5)
    assert len(points) >= 3
    bounds = [0] + points + [x.shape[0]]
    rows = torch.tensor(bounds[:-1], dtype=torch.int64)
    frames = torch.tensor(np.diff(bounds), dtype=torch.int32)
    A = torch.from_numpy(cases.log_triangular_transition(1440)).cuda()
    init = torch.full((1440,), float(np.log(1.0 / 1440 + np.finfo(np.float32).tiny)), device='cuda:0')
    got = torbi_b200.decode_packed(x, rows, frames, A, init).cpu().numpy()
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        want = oracle.c_decode(observation[None, lo:hi], [hi - lo], A.cpu().numpy(), init.cpu().numpy(), exact_scan=False)
        assert np.array_equal(got[lo:hi], want[0])


@pytest.mark.parametrize('dtype', [torch.float32, torch.float16])
def test_file_pipeline_with_device_chunking(tmp_path, dtype, oracle):
    """from_files_to_files with MIN_CHUNK_SIZE: native packed route (device entropy, pieces decoded in
    place) against the DataLoader route of the reference (CPU chunking, padded collate, separate) and,
    for one file, against the oracle chunk by chunk."""
    rng = np.random.default_rng(31)
    S = 96
    files, outputs, tensors = [], [], []
    for i, f in enumerate([300, 41, 7, 150, 300, 1, 90]):
        logits = rng.standard_normal((f, S))
        logits[rng.random(f) < 0.6] *= 12.0
        x = torch.from_numpy(cases.log_softmax(logits)).to(dtype)
        tensors.append(x)
        files.append(tmp_path / f'in{i}.pt')
        torch.save(x, files[-1])
        outputs.append(tmp_path / f'out{i}.pt')
    transition = torch.from_numpy(cases.triangular_transition(S, 9.5))
    torch.save(transition, tmp_path / 'transition.pt')
    old = torbi_b200.BATCH_SIZE, torbi_b200.MIN_CHUNK_SIZE
    torbi_b200.BATCH_SIZE, torbi_b200.MIN_CHUNK_SIZE = 4, 16
    torbi_b200.config.defaults.MIN_CHUNK_SIZE = 16
    try:
        from torbi_b200 import core
        core._PIPELINE_STATS.clear()
        torbi_b200.from_files_to_files(files, outputs, tmp_path / 'transition.pt', None, log_probs=True, gpu=0)
        assert core._PIPELINE_STATS.get('pieces', 0) > len(files)              # the packed route chunked on the device
        native = [torch.load(out) for out in outputs]
        os.environ['TORBI_B200_NATIVE_IO'] = '0'
        try:
            torbi_b200.from_files_to_files(files, outputs, tmp_path / 'transition.pt', None, log_probs=True, gpu=0)
        finally:
            del os.environ['TORBI_B200_NATIVE_IO']
        for got, out, x in zip(native, outputs, tensors):
            assert got.dtype == torch.int32 and got.shape == (x.shape[0],)
            assert torch.equal(got, torch.load(out).to(torch.int32))
        # file 0 against the oracle, chunk by chunk
        x = tensors[0]
        points = chunk_module.split(x, 16, torbi_b200.ENTROPY_THRESHOLD)
        assert len(points) >= 2
        A = torch.log(transition + torch.finfo(torch.float32).tiny)
        bounds = [0] + points + [x.shape[0]]
        want = np.concatenate([
            oracle.ref_from_probabilities(x[lo:hi].to(torch.float32).unsqueeze(0), None, A, None, True)[0]
            for lo, hi in zip(bounds[:-1], bounds[1:])])
        assert np.array_equal(native[0].numpy(), want)
    finally:
        torbi_b200.BATCH_SIZE, torbi_b200.MIN_CHUNK_SIZE = old
        torbi_b200.config.defaults.MIN_CHUNK_SIZE = old[1]


def test_file_pipeline_packed_equals_padded(tmp_path):
    rng = np.random.default_rng(5)
    S = 1440
    files, outputs = [], []
    for i, f in enumerate(rng.integers(1, 120, size=40)):
        files.append(tmp_path / f'in{i}.pt')
        torch.save(torch.from_numpy(cases.log_softmax(rng.standard_normal((int(f), S)))), files[-1])
        outputs.append(tmp_path / f'out{i}.pt')
    torbi_b200.from_files_to_files(files, outputs, None, None, log_probs=True, gpu=0)
    packed = [torch.load(out) for out in outputs]
    os.environ['TORBI_B200_PACKED'] = '0'
    try:
        torbi_b200.from_files_to_files(files, outputs, None, None, log_probs=True, gpu=0)
    finally:
        del os.environ['TORBI_B200_PACKED']
    for got, out in zip(packed, outputs):
        assert torch.equal(got, torch.load(out))


def test_evaluation_runner_on_a_synthetic_corpus(tmp_path):
    """python -m torbi_b200.evaluate (evaluate/core.py:14-123): un-chunked configuration against itself
    (RPA = 1), then the chunked configuration against the un-chunked decode and against the true contour."""
    torbi_b200.evaluate.synthetic_dataset('synthetic', tmp_path, files=12, median_frames=300, seed=3)
    result = torbi_b200.evaluate.datasets(['synthetic'], gpu=0, root=tmp_path)['synthetic']
    assert result['rpa'] == {'0': 1.0, '1': 1.0, '2': 1.0} and result['files'] == 12
    assert result['rtf']['torbi'] > result['rtf']['end_to_end'] > 1.0
    assert result['rpa_vs_truth']['2'] > 0.6
    old = torbi_b200.CONFIG, torbi_b200.MIN_CHUNK_SIZE
    torbi_b200.CONFIG, torbi_b200.MIN_CHUNK_SIZE = 'chunked', 32
    try:
        chunked = torbi_b200.evaluate.datasets(['synthetic'], gpu=0, root=tmp_path)['synthetic']
    finally:
        torbi_b200.CONFIG, torbi_b200.MIN_CHUNK_SIZE = old
    assert chunked['rpa']['2'] > 0.95 and chunked['frames'] == result['frames']
