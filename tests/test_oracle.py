"""CPU: pin the oracle (oracle/viterbi_oracle.c, oracle/oracle.py) against the
reference's known-answer vector, the golden vectors produced by the unmodified
reference, and -- when oracle/_ref is present -- the reference operator itself."""
import hashlib

import numpy as np
import pytest

import cases


def sha(a):
    a = np.ascontiguousarray(a)
    return hashlib.sha256(str(a.dtype).encode() + str(a.shape).encode() + a.tobytes()).hexdigest()


def pack_trellis(trellis, frames):
    return np.concatenate([trellis[b, 1:f].reshape(-1) for b, f in enumerate(frames)]
                          + [np.zeros(0, np.int32)])


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.int32)


def test_known_answer(oracle):
    # /root/reference/tests/test_core.py:7-25
    case = cases.kat()
    for decoder in ('c', 'numpy'):
        out = oracle.ref_from_probabilities(
            case['observation'], None, case['transition'], case['initial'], False, decoder=decoder)
        assert out.tolist() == [[1, 2, 2]]


@pytest.mark.parametrize('name', list(cases.CASES))
def test_inputs_regenerate(name, golden):
    assert cases.digest(cases.CASES[name]()) == str(golden[f'{name}/digest'])


@pytest.mark.parametrize('name', list(cases.CASES))
@pytest.mark.parametrize('exact_scan', [True, False])
def test_c_oracle_matches_reference_golden(name, exact_scan, golden, oracle):
    case = cases.CASES[name]()
    obs, frames, A, init = oracle.ref_prepare(
        case['observation'], case['batch_frames'], case['transition'], case['initial'],
        case['log_probs'])
    # conversion restated with the same torch CPU ops as core.py:189-197
    assert sha(obs.numpy()) == str(golden[f'{name}/prep_sha'])
    indices, posterior, trellis = oracle.c_decode(
        obs.numpy(), frames.numpy(), A.numpy(), init.numpy(), exact_scan=exact_scan,
        want_posterior=True, want_trellis=True)
    assert np.array_equal(indices, golden[f'{name}/indices'])
    assert np.array_equal(indices, golden[f'{name}/api_indices'])
    assert np.array_equal(bits(posterior), bits(golden[f'{name}/posterior']))
    packed = pack_trellis(trellis, frames.numpy())
    if f'{name}/trellis' in golden:
        assert np.array_equal(packed, golden[f'{name}/trellis'])
    else:
        assert sha(packed.astype(np.int32)) == str(golden[f'{name}/trellis_sha'])


@pytest.mark.parametrize('name', [n for n in cases.CASES if '1440' not in n and '640' not in n])
def test_numpy_oracle_matches_reference_golden(name, golden, oracle):
    case = cases.CASES[name]()
    obs, frames, A, init = oracle.ref_prepare(
        case['observation'], case['batch_frames'], case['transition'], case['initial'],
        case['log_probs'])
    indices, posterior = oracle.numpy_decode(
        obs.numpy(), frames.numpy(), A.numpy(), init.numpy(), want_posterior=True)
    assert np.array_equal(indices, golden[f'{name}/indices'])
    assert np.array_equal(bits(posterior), bits(golden[f'{name}/posterior']))


def test_c_oracle_matches_compiled_reference(oracle):
    if not oracle.ref_available():
        pytest.skip('oracle/_ref not built (needs /root/reference)')
    rng = np.random.default_rng(99)
    for (B, T, S) in [(3, 21, 7), (2, 33, 65), (9, 12, 130)]:
        obs = cases.log_softmax(rng.standard_normal((B, T, S)))
        obs = (np.round(obs * 3) / 3).astype(np.float32)          # plenty of ties
        A = (np.round(cases.log_softmax(rng.standard_normal((S, S))) * 2) / 2).astype(np.float32)
        init = cases.log_softmax(rng.standard_normal((S,)))
        frames = rng.integers(1, T + 1, size=B).astype(np.int32)
        ref_idx, ref_post, ref_trellis = oracle.ref_decode(obs, frames, A, init, want_debug=True)
        idx, post, trellis = oracle.c_decode(obs, frames, A, init, want_posterior=True, want_trellis=True)
        assert np.array_equal(idx, ref_idx)
        assert np.array_equal(bits(post), bits(ref_post))
        assert np.array_equal(pack_trellis(trellis, frames), pack_trellis(ref_trellis, frames))


def test_semantics_orientation_ties_tail(oracle):
    # A2 orientation [dest][src], A3 lowest index wins, A1 tail fill
    S = 40
    A = np.full((S, S), -1e30, np.float32)
    A[1, 0] = 0.0
    init = np.full((S,), -1e30, np.float32)
    init[0] = 0.0
    obs = np.zeros((1, 2, S), np.float32)
    assert oracle.c_decode(obs, [2], A, init).tolist() == [[0, 1]]
    assert oracle.c_decode(obs, [2], A.T.copy(), init).tolist() == [[0, 0]]
    # tie between s=1 and s=32: lowest wins (the reference CUDA kernel returns 32)
    A = np.full((S, S), -5.0, np.float32)
    A[:, 1] = 0.0
    A[:, 32] = 0.0
    init = np.zeros((S,), np.float32)
    obs = np.zeros((1, 3, S), np.float32)
    assert oracle.c_decode(obs, [3], A, init).tolist() == [[1, 1, 0]]
    # ragged: positions >= frames-1 hold the final state
    rng = np.random.default_rng(5)
    obs = cases.log_softmax(rng.standard_normal((1, 9, S)))
    A = cases.log_softmax(rng.standard_normal((S, S)))
    idx = oracle.c_decode(obs, [4], A, init)
    assert (idx[0, 3:] == idx[0, 3]).all()
    alone = oracle.c_decode(obs[:, :4], [4], A, init)
    assert np.array_equal(idx[0, :4], alone[0])
    with pytest.raises(ValueError):
        oracle.c_decode(obs, [0], A, init)
