"""GPU: parity of the CUDA path (through the C ABI) with the oracle and with the
golden vectors of the unmodified reference.  Integer outputs and fp32 forward
scores are compared BIT-EXACT; nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

import cases
import torbi_b200
from torbi_b200 import _ffi

pytestmark = pytest.mark.gpu

KERNELS = {'auto': 0, 'generic': _ffi.FORCE_GENERIC, 'tiled16': _ffi.TILE_ROWS_16,
           'tiled32': _ffi.TILE_ROWS_32, 'tiled64': _ffi.TILE_ROWS_64, 'wide': _ffi.FORCE_WIDE,
           'banded': _ffi.BANDED,
           'banded16': _ffi.BANDED | _ffi.TILE_ROWS_16,
           # the tiled kernel's backpointer mode (default is score history + recomputed backtrace)
           'tiled64bp': _ffi.TILE_ROWS_64 | _ffi.BACKPOINTERS, 'tiled16bp': _ffi.TILE_ROWS_16 | _ffi.BACKPOINTERS,
           'tiled32bp': _ffi.TILE_ROWS_32 | _ffi.BACKPOINTERS, 'bandedbp': _ffi.BANDED | _ffi.BACKPOINTERS,
           # split-K 16 x 40 tiles (small and mid-size batches)
           'split': _ffi.FORCE_SPLIT, 'split_streamed': _ffi.FORCE_SPLIT | _ffi.SPLIT_STREAMED}


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def cuda(x):
    return torch.as_tensor(np.asarray(x)).to('cuda:0')


def gpu_decode(obs, frames, A, init, flags=0):
    indices, posterior = torbi_b200.decode(
        cuda(obs), torch.as_tensor(np.asarray(frames)), cuda(A), cuda(init), flags=flags,
        return_posterior=True)
    torch.cuda.synchronize()
    return indices.cpu().numpy(), posterior.cpu().numpy()


def test_library_is_loaded_and_device_is_sm100():
    assert torbi_b200.LIBRARY.exists()
    assert _ffi.library().torbi_b200_device_count() >= 1
    assert torch.cuda.get_device_capability(0)[0] == 10


# ---------------------------------------------------------------------------
# golden vectors of the reference (tests/golden/golden.npz)
# ---------------------------------------------------------------------------

@pytest.mark.parametrize('kernel', list(KERNELS))
@pytest.mark.parametrize('name', list(cases.CASES))
def test_decode_matches_reference_golden(name, kernel, golden, oracle):
    """Gate (A): identical converted fp32 inputs -> bit-exact indices and scores."""
    case = cases.CASES[name]()
    obs, frames, A, init = oracle.ref_prepare(
        case['observation'], case['batch_frames'], case['transition'], case['initial'],
        case['log_probs'])
    indices, posterior = gpu_decode(obs.numpy(), frames.numpy(), A.numpy(), init.numpy(),
                                    flags=KERNELS[kernel])
    assert np.array_equal(indices, golden[f'{name}/indices'])
    assert np.array_equal(bits(posterior), bits(golden[f'{name}/posterior']))


@pytest.mark.parametrize('conversion', ['host', 'fused', 'torch'])
@pytest.mark.parametrize('name', list(cases.CASES))
def test_from_probabilities_matches_reference_golden(name, conversion, golden):
    """The public entry point against the reference's own from_probabilities
    output (CPU path).  'host' repeats the reference's CPU conversion exactly;
    'fused' converts inside the kernel (expf/logf of the CUDA math library)."""
    case = cases.CASES[name]()

    def tensor(x):
        return None if x is None else torch.from_numpy(np.array(x))

    observation = tensor(case['observation'])
    before = observation.clone()
    indices = torbi_b200.from_probabilities(
        observation, tensor(case['batch_frames']), tensor(case['transition']),
        tensor(case['initial']), log_probs=case['log_probs'], gpu=0, conversion=conversion)
    assert indices.dtype == torch.int32 and indices.device.type == 'cuda'
    assert torch.equal(observation, before)                     # caller's tensor untouched
    assert np.array_equal(indices.cpu().numpy(), golden[f'{name}/api_indices'])


def test_known_answer_on_gpu_like_reference_test_cuda():
    # /root/reference/tests/test_core.py:27-46
    case = cases.kat()
    bins = torbi_b200.from_probabilities(
        observation=torch.tensor(case['observation']).to('cuda:0'),
        transition=torch.tensor(case['transition']).to('cuda:0'),
        initial=torch.tensor(case['initial']).to('cuda:0'),
        log_probs=False, gpu=0)
    assert (bins == torch.tensor([1, 2, 2]).to('cuda:0')).all()
    # gpu=None with CUDA inputs uses their device
    bins = torbi_b200.from_probabilities(
        observation=torch.tensor(case['observation']).to('cuda:0'),
        transition=torch.tensor(case['transition']), initial=torch.tensor(case['initial']))
    assert bins.tolist() == [[1, 2, 2]]


# ---------------------------------------------------------------------------
# seeded random inputs against the C oracle
# ---------------------------------------------------------------------------

def random_problem(seed, B, T, S, ragged=True, quant=None, transition='dense'):
    rng = np.random.default_rng(seed)
    obs = cases.log_softmax(rng.standard_normal((B, T, S)))
    if transition == 'dense':
        A = cases.log_softmax(rng.standard_normal((S, S)))
    else:
        A = cases.log_triangular_transition(S, min(87.208, S / 4))
    init = cases.log_softmax(rng.standard_normal((S,)))
    if quant:
        obs = (np.round(obs * quant) / quant).astype(np.float32)
        A = (np.round(A * quant) / quant).astype(np.float32)
    frames = rng.integers(1, T + 1, size=B).astype(np.int32) if ragged else np.full(B, T, np.int32)
    frames[0] = T
    return obs, frames, A, init


SHAPES = [
    # (B, T, S, quant, transition)
    (1, 50, 1440, None, 'dense'),
    (8, 64, 1440, None, 'triangular'),
    (37, 20, 1440, 1, 'dense'),
    (130, 9, 1440, None, 'triangular'),
    (3, 40, 1000, None, 'dense'),
    (5, 33, 1441, 2, 'dense'),
    (4, 16, 4096, None, 'dense'),
    (70, 12, 96, 2, 'dense'),
    (600, 6, 160, None, 'dense'),
    (2, 300, 64, None, 'dense'),
]


@pytest.mark.parametrize('kernel', list(KERNELS))
@pytest.mark.parametrize('shape', SHAPES, ids=[f'B{s[0]}-T{s[1]}-S{s[2]}' for s in SHAPES])
def test_decode_matches_oracle(shape, kernel, oracle):
    B, T, S, quant, transition = shape
    obs, frames, A, init = random_problem(B * 1000 + S, B, T, S, quant=quant, transition=transition)
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    indices, posterior = gpu_decode(obs, frames, A, init, flags=KERNELS[kernel])
    assert np.array_equal(indices, want_idx)
    assert np.array_equal(bits(posterior), bits(want_post))


@pytest.mark.parametrize('outside', [-87.33655, -np.inf, -3.0])
@pytest.mark.parametrize('states,width', [(1440, 87.2), (1000, 3.0), (333, 40.5), (2048, 300.0)])
def test_banded_fast_path_is_exact(states, width, outside, oracle):
    """Banded matrices with different constants outside the band, quantised so that ties
    between in-window and out-of-window sources occur; banded == dense == oracle."""
    rng = np.random.default_rng(states)
    B, T = 7, 12
    i = np.arange(states)
    band = np.abs(i[:, None] - i[None, :]) < width
    A = np.where(band, np.round(rng.standard_normal((states, states)) * 2) / 2 - 1.0, outside).astype(np.float32)
    A[5, :] = np.float32(outside)                                  # a row with nothing in its band
    obs = (np.round(cases.log_softmax(rng.standard_normal((B, T, states))) * 2) / 2).astype(np.float32)
    init = cases.log_softmax(rng.standard_normal((states,)))
    frames = rng.integers(1, T + 1, size=B).astype(np.int32)
    frames[0] = T
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    for flags in (_ffi.BANDED, _ffi.BANDED | _ffi.TILE_ROWS_32, 0):
        indices, posterior = gpu_decode(obs, frames, A, init, flags=flags)
        assert np.array_equal(indices, want_idx)
        assert np.array_equal(bits(posterior), bits(want_post))


@pytest.mark.parametrize('kernel', ['tiled16', 'banded16', 'auto'])
def test_many_small_batches_use_every_cta_group(kernel, oracle):
    """Few states, many sequences: one CTA per group, 148 groups, every scheduling word in use."""
    obs, frames, A, init = random_problem(9, 3000, 3, 48)
    want = oracle.c_decode(obs, frames, A, init)
    indices, _ = gpu_decode(obs, frames, A, init, flags=KERNELS[kernel])
    assert np.array_equal(indices, want)


def test_very_wide_state_space(oracle):
    """S = 20000: beyond the tiled kernel (250 destination tiles > 148 SMs); the selector
    falls back to the wide kernel, and the generic kernel must agree."""
    obs, frames, A, init = random_problem(5, 2, 4, 20000, ragged=False)
    want = oracle.c_decode(obs, frames, A, init, exact_scan=False)
    for flags in (0, _ffi.FORCE_GENERIC, _ffi.FORCE_WIDE):
        indices, _ = gpu_decode(obs, frames, A, init, flags=flags)
        assert np.array_equal(indices, want)
    with pytest.raises(ValueError, match='tiled kernel does not support'):
        gpu_decode(obs, frames, A, init, flags=_ffi.FORCE_TILED)


def test_padding_is_never_read(oracle):
    # garbage (NaN) beyond batch_frames must not influence anything
    obs, frames, A, init = random_problem(7, 6, 30, 200)
    dirty = obs.copy()
    for b, f in enumerate(frames):
        dirty[b, f:] = np.nan
    a, pa = gpu_decode(obs, frames, A, init)
    b, pb = gpu_decode(dirty, frames, A, init)
    assert np.array_equal(a, b) and np.array_equal(bits(pa), bits(pb))
    assert np.array_equal(a, oracle.c_decode(obs, frames, A, init))


def test_each_item_equals_decoding_it_alone():
    obs, frames, A, init = random_problem(11, 9, 25, 1440, quant=1)
    together, _ = gpu_decode(obs, frames, A, init)
    for b in range(obs.shape[0]):
        alone, _ = gpu_decode(obs[b:b + 1, :frames[b]], frames[b:b + 1], A, init)
        assert np.array_equal(together[b, :frames[b]], alone[0])
        assert (together[b, frames[b] - 1:] == alone[0, -1]).all()


def test_fused_conversion_equals_torch_cuda_conversion():
    """Contract (B) of SURVEY H6: expf/+tiny/logf inside the kernel == torch's
    CUDA exp_/+=/log_ followed by a plain decode, bit for bit."""
    rng = np.random.default_rng(3)
    S = 256
    prob = np.exp(cases.log_softmax(rng.standard_normal((4, 40, S)) * 12.0).astype(np.float64)).astype(np.float32)
    A = cuda(cases.log_softmax(rng.standard_normal((S, S))))
    init = cuda(cases.log_softmax(rng.standard_normal((S,))))
    frames = torch.tensor([40, 7, 40, 1])
    x = cuda(prob)
    converted = torch.log(x)
    torch.exp_(converted)
    converted += torch.finfo(torch.float32).tiny
    torch.log_(converted)
    want, want_post = torbi_b200.decode(converted, frames, A, init, return_posterior=True)
    got, got_post = torbi_b200.decode(
        x, frames, A, init, flags=_ffi.OBS_IS_PROBABILITY | _ffi.EPS_FLOOR, return_posterior=True)
    assert torch.equal(want, got)
    assert torch.equal(want_post.view(torch.int32), got_post.view(torch.int32))


def test_host_entry_matches_device_entry(oracle):
    obs, frames, A, init = random_problem(21, 12, 30, 320)
    want = oracle.c_decode(obs, frames, A, init)
    host = torbi_b200.decode_host(
        torch.from_numpy(obs).pin_memory(), torch.from_numpy(frames), torch.from_numpy(A),
        torch.from_numpy(init), gpu=0)
    assert host.device.type == 'cpu' and np.array_equal(host.numpy(), want)
    devices = list(range(min(torch.cuda.device_count(), 8)))
    multi, post = torbi_b200.decode_host(
        torch.from_numpy(obs), torch.from_numpy(frames), torch.from_numpy(A),
        torch.from_numpy(init), gpu=devices + [devices[0]], return_posterior=True)
    assert np.array_equal(multi.numpy(), want)
    _, want_post = oracle.c_decode(obs, frames, A, init, want_posterior=True)
    assert np.array_equal(bits(post.numpy()), bits(want_post))


def test_host_entry_streams_large_pinned_batches():
    """Pinned batches >= 64 MB are copied in along T on a second stream while the
    forward kernel runs (it polls a frames-ready word); results must not change."""
    rng = np.random.default_rng(31)
    B, T, S = 20, 700, 1440                                   # 80 MB
    obs = torch.from_numpy(cases.log_softmax(rng.standard_normal((B, T, S)))).pin_memory()
    frames = torch.from_numpy(rng.integers(1, T + 1, size=B).astype(np.int32))
    frames[3] = T
    A = torch.from_numpy(cases.log_triangular_transition(S))
    init = torch.from_numpy(cases.log_softmax(rng.standard_normal((S,))))
    want = torbi_b200.decode(obs.cuda(), frames, A.cuda(), init.cuda()).cpu()
    for _ in range(2):
        got = torbi_b200.decode_host(obs, frames, A, init, gpu=0)
        assert torch.equal(got, want)
    pageable = torbi_b200.decode_host(obs.clone(), frames, A, init, gpu=0)   # not pinned: copied up front
    assert torch.equal(pageable, want)


def test_multi_gpu_list_api(oracle):
    obs, frames, A, init = random_problem(22, 10, 20, 128)
    want = oracle.c_decode(obs, frames, A, init)
    devices = list(range(torch.cuda.device_count())) * 2        # >= 2 shards even on one GPU
    got = torbi_b200.from_probabilities(
        torch.from_numpy(obs), torch.from_numpy(frames), torch.from_numpy(A),
        torch.from_numpy(init), log_probs=True, gpu=devices[:4], conversion='host')
    assert got.device == torch.device('cuda:0')
    # log(exp(x)+tiny) is the identity on these inputs, so the plain oracle applies
    assert np.array_equal(got.cpu().numpy(), want)


def test_dispatcher_op_on_gpu(oracle):
    obs, frames, A, init = random_problem(41, 5, 12, 160)
    got = torch.ops.torbi_b200.viterbi_decode(cuda(obs), torch.from_numpy(frames), cuda(A), cuda(init))
    assert np.array_equal(got.cpu().numpy(), oracle.c_decode(obs, frames, A, init))


def test_uniform_default_transition_through_the_banded_path(oracle):
    """transition=None on a big enough batch: the all-constant default matrix has an empty
    window in every tile, so the whole step is the outside scan -- still the reference's answer
    (lots of exact ties: every destination sees the same candidates)."""
    rng = np.random.default_rng(77)
    obs = torch.from_numpy(cases.log_softmax(rng.standard_normal((40, 9, 512))))
    frames = torch.from_numpy(rng.integers(1, 10, size=40).astype(np.int32))
    got = torbi_b200.from_probabilities(obs, frames, log_probs=True, gpu=0)
    assert _ffi.library().torbi_b200_last_kernel().decode().startswith('forward_tiled_banded')
    want = oracle.ref_from_probabilities(obs, frames, None, None, True)
    assert np.array_equal(got.cpu().numpy(), want)


def test_errors_on_gpu():
    obs = torch.rand(2, 4, 5, device='cuda:0')
    A, init = torch.rand(5, 5, device='cuda:0'), torch.rand(5, device='cuda:0')
    with pytest.raises(ValueError, match='batch_frames'):
        torbi_b200.decode(obs, torch.tensor([4, 0]), A, init)
    with pytest.raises(ValueError, match='batch_frames'):
        torbi_b200.decode(obs, torch.tensor([4, 5]), A, init)
    with pytest.raises(RuntimeError, match='same device'):
        torbi_b200.decode(obs, torch.tensor([4, 4]), A.cpu(), init)
    # a good call still works afterwards
    assert torbi_b200.decode(obs, torch.tensor([4, 4]), A, init).shape == (2, 4)


# ---------------------------------------------------------------------------
# full-size configs: size-independent properties + spot checks against the oracle
# ---------------------------------------------------------------------------

def replay_path_score(obs, A, init, path):
    """fp32 score of `path` accumulated in the recursion's own order:
    P = obs[0,i0] + init[i0];  P = obs[t,it] + (P + A[it, it-1])."""
    p = np.float32(obs[0, path[0]]) + np.float32(init[path[0]])
    for t in range(1, len(path)):
        p = np.float32(obs[t, path[t]]) + np.float32(p + np.float32(A[path[t], path[t - 1]]))
    return np.float32(p)


def torch_convert(x):
    """log(exp(x) + tiny) with torch's CUDA kernels (bit-identical to the fused
    conversion, see test_fused_conversion_equals_torch_cuda_conversion)."""
    y = x.clone()
    torch.exp_(y)
    y += torch.finfo(torch.float32).tiny
    torch.log_(y)
    return y


def full_size_problem(B, T, S, seed=1234):
    g = torch.Generator(device='cuda:0').manual_seed(seed)
    obs = torch.log_softmax(torch.randn(B, T, S, device='cuda:0', generator=g), dim=-1)
    A = cuda(cases.log_triangular_transition(S))
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device='cuda:0')
    return obs, A, init


def test_config2_full_size_properties(oracle):
    """BASELINE config 2: S=1440, B=512, T=2000, triangular transition."""
    B, T, S = 512, 2000, 1440
    obs, A, init = full_size_problem(B, T, S)
    frames = torch.full((B,), T, dtype=torch.int32)
    indices, posterior = torbi_b200.decode(obs, frames, A, init, flags=_ffi.EPS_FLOOR,
                                           return_posterior=True)
    again = torbi_b200.decode(obs, frames, A, init, flags=_ffi.EPS_FLOOR)
    assert torch.equal(indices, again)                           # deterministic
    assert int(indices.min()) >= 0 and int(indices.max()) < S
    # the final state is the first-occurrence argmax of the final scores
    assert torch.equal(indices[:, -1].long(), posterior.argmax(dim=1))
    # replaying each decoded path in fp32 reproduces the winning score exactly
    A_host, init_host = A.cpu().numpy(), init.cpu().numpy()
    for b in (0, 17, 255, 511):
        path = indices[b].cpu().numpy()
        score = replay_path_score(torch_convert(obs[b]).cpu().numpy(), A_host, init_host, path)
        assert bits(score) == bits(posterior[b].max().cpu().numpy())
    # spot check against the oracle: two whole items
    for b in (3, 400):
        want, want_post = oracle.c_decode(
            torch_convert(obs[b:b + 1]).cpu().numpy(), [T], A_host, init_host, exact_scan=False,
            want_posterior=True)
        assert np.array_equal(indices[b].cpu().numpy(), want[0])
        assert np.array_equal(bits(posterior[b].cpu().numpy()), bits(want_post[0]))


def test_config3_ragged_batch(oracle):
    """BASELINE config 3 (scaled to fit the test budget): S=1440, ragged lengths
    log-uniform in [200, 6000], batch 48; checked per item (the reference cannot
    index the padded batch, B*T*S >= 2^31 at full size)."""
    B, S = 48, 1440
    rng = np.random.default_rng(1234)
    frames = np.exp(rng.uniform(np.log(200), np.log(6000), size=B)).round().astype(np.int32)
    frames[0], frames[1] = 6000, 200
    T = int(frames.max())
    obs, A, init = full_size_problem(B, T, S, seed=4321)
    indices, posterior = torbi_b200.decode(obs, torch.from_numpy(frames), A, init,
                                           return_posterior=True)
    idx = indices.cpu().numpy()
    A_host, init_host = A.cpu().numpy(), init.cpu().numpy()
    for b in range(B):
        assert (idx[b, frames[b] - 1:] == idx[b, frames[b] - 1]).all()       # tail fill
    for b in (0, 1, 7, 30):
        f = int(frames[b])
        score = replay_path_score(obs[b, :f].cpu().numpy(), A_host, init_host, idx[b, :f])
        assert bits(score) == bits(posterior[b].max().cpu().numpy())
    for b in (1, 5):
        f = int(frames[b])
        want = oracle.c_decode(obs[b:b + 1, :f].cpu().numpy(), [f], A_host, init_host, exact_scan=False)
        assert np.array_equal(idx[b, :f], want[0])


# ---------------------------------------------------------------------------
# file drivers
# ---------------------------------------------------------------------------

def test_file_drivers(tmp_path, oracle):
    rng = np.random.default_rng(8)
    S = 96
    lengths = [30, 5, 17, 1, 30, 12]
    files = []
    for i, f in enumerate(lengths):
        x = torch.from_numpy(cases.log_softmax(rng.standard_normal((f, S))))
        path = tmp_path / f'in{i}.pt'
        torch.save(x, path)
        files.append(path)
    transition = torch.from_numpy(cases.triangular_transition(S, 9.5))       # probabilities
    torch.save(transition, tmp_path / 'transition.pt')
    initial = torch.from_numpy(cases.log_softmax(rng.standard_normal((S,))))  # log space
    torch.save(initial, tmp_path / 'initial.pt')
    outputs = [tmp_path / f'out{i}.pt' for i in range(len(files))]

    old = torbi_b200.BATCH_SIZE
    torbi_b200.BATCH_SIZE = 4                        # two batches
    try:
        torbi_b200.from_files_to_files(
            files, outputs, tmp_path / 'transition.pt', tmp_path / 'initial.pt',
            log_probs=True, gpu=0)
    finally:
        torbi_b200.BATCH_SIZE = old
    tiny = torch.finfo(torch.float32).tiny
    A = torch.log(transition + tiny).numpy()                                  # core.py:343-345
    for path, out, f in zip(files, outputs, lengths):
        got = torch.load(out)
        assert got.dtype == torch.int32 and got.shape == (f,)
        x = torch.load(path).unsqueeze(0)
        want = oracle.ref_from_probabilities(x, None, torch.from_numpy(A), initial, True)
        assert np.array_equal(got.numpy(), want[0])

    # several "GPUs" (the same one twice here): files sharded by size, one worker thread each
    for out in outputs:
        out.unlink()
    torbi_b200.from_files_to_files(
        files, outputs, tmp_path / 'transition.pt', tmp_path / 'initial.pt', log_probs=True, gpu=[0, 0])
    for path, out, f in zip(files, outputs, lengths):
        want = oracle.ref_from_probabilities(torch.load(path).unsqueeze(0), None, torch.from_numpy(A), initial, True)
        assert np.array_equal(torch.load(out).numpy(), want[0])

    # the DataLoader route of the reference instead of the native reader / writer: same files
    import os
    for out in outputs:
        out.unlink()
    os.environ['TORBI_B200_NATIVE_IO'] = '0'
    try:
        torbi_b200.from_files_to_files(
            files, outputs, tmp_path / 'transition.pt', tmp_path / 'initial.pt', log_probs=True, gpu=0)
    finally:
        del os.environ['TORBI_B200_NATIVE_IO']
    for path, out, f in zip(files, outputs, lengths):
        want = oracle.ref_from_probabilities(torch.load(path).unsqueeze(0), None, torch.from_numpy(A), initial, True)
        assert np.array_equal(torch.load(out).numpy(), want[0])

    # a half-precision and a legacy-format input in the batch: per-batch / per-file fallbacks
    odd = [tmp_path / 'half.pt', tmp_path / 'legacy.pt']
    torch.save(torch.load(files[0]).half(), odd[0])
    torch.save(torch.load(files[2]), odd[1], _use_new_zipfile_serialization=False)
    odd_out = [tmp_path / 'half.out.pt', tmp_path / 'legacy.out.pt']
    torbi_b200.from_files_to_files(
        [files[1], odd[1]], [outputs[1], odd_out[1]], tmp_path / 'transition.pt', tmp_path / 'initial.pt',
        log_probs=True, gpu=0)
    assert torch.equal(torch.load(odd_out[1]), torch.load(outputs[2]))
    torbi_b200.from_files_to_files(
        [files[1], odd[0]], [outputs[1], odd_out[0]], tmp_path / 'transition.pt', tmp_path / 'initial.pt',
        log_probs=True, gpu=0)
    want = oracle.ref_from_probabilities(
        torch.load(odd[0]).unsqueeze(0), None, torch.from_numpy(A), initial, True)
    assert np.array_equal(torch.load(odd_out[0]).numpy(), want[0])

    # from_file: transition logged WITHOUT tiny (core.py:247-250), shape (1, frames)
    got = torbi_b200.from_file(files[0], tmp_path / 'transition.pt', tmp_path / 'initial.pt',
                               log_probs=True, gpu=0)
    assert got.shape == (1, lengths[0])
    want = oracle.ref_from_probabilities(
        torch.load(files[0]).unsqueeze(0), None, torch.log(transition), initial, True)
    assert np.array_equal(got.cpu().numpy(), want)
    torbi_b200.from_file_to_file(files[2], tmp_path / 'single.pt', gpu=0, log_probs=True)
    assert torch.load(tmp_path / 'single.pt').shape == (1, lengths[2])

    # chunked mode: pieces decode independently and are re-joined per file
    torbi_b200.MIN_CHUNK_SIZE = 4
    try:
        torbi_b200.from_files_to_files(files[:3], outputs[:3], log_probs=True, gpu=0)
        for path, out, f in zip(files[:3], outputs[:3], lengths[:3]):
            x = torch.load(path)
            pieces = torbi_b200.chunk(x)
            want = np.concatenate([
                oracle.ref_from_probabilities(p.unsqueeze(0), None, None, None, True)[0]
                for p in pieces])
            assert np.array_equal(torch.load(out).numpy(), want)
    finally:
        torbi_b200.MIN_CHUNK_SIZE = None


def test_tiled_kernel_modes_report_their_backtrace(oracle):
    """Score-history mode (default) and backpointer mode of the tiled kernel agree bit for bit and
    report which backtrace ran; the workspace is sized per mode."""
    lib = _ffi.library()
    obs, frames, A, init = random_problem(77, 70, 40, 1440, quant=1)      # quantised: many exact ties
    want_idx, want_post = oracle.c_decode(obs, frames, A, init, exact_scan=False, want_posterior=True)
    for flags, name in ((_ffi.FORCE_TILED, 'backtrace_recompute'),
                        (_ffi.FORCE_TILED | _ffi.BACKPOINTERS, 'backtrace'),
                        (_ffi.FORCE_WIDE, 'backtrace'), (_ffi.FORCE_GENERIC, 'backtrace')):
        indices, posterior = gpu_decode(obs, frames, A, init, flags=flags)
        assert lib.torbi_b200_last_backtrace().decode() == name
        assert np.array_equal(indices, want_idx)
        assert np.array_equal(bits(posterior), bits(want_post))
    hist = lib.torbi_b200_workspace_bytes(70, 40, 1440, 0)
    bp = lib.torbi_b200_workspace_bytes(70, 40, 1440, _ffi.BACKPOINTERS)
    assert hist - bp == 2 * 70 * 40 * 1440


def test_eps_floor_shortcut_is_exact_for_every_fp32_input():
    """The kernels return x unchanged where logf(expf(x) + FLT_MIN) == x; the library checks that
    shortcut against the plain formula for all 2^32 bit patterns on this device."""
    import ctypes
    mismatches = ctypes.c_uint64(123)
    _ffi.check(_ffi.library().torbi_b200_selftest_conversion(ctypes.byref(mismatches)))
    assert mismatches.value == 0
