"""Deterministic input generators shared by the golden-vector generator
(make_golden.py, run once where /root/reference exists), the CPU tests and the
GPU parity tests.  Inputs are regenerated from seeds (numpy PCG64) and verified
against the sha256 stored beside the expected outputs, so the fixtures stay small.

Each case returns a dict with the arguments of torbi.from_probabilities
(/root/reference/torbi/core.py:110-118): observation (B,T,S), batch_frames or
None, transition or None, initial or None, log_probs.
"""
import hashlib

import numpy as np

TINY = np.finfo(np.float32).tiny


def log_softmax(x):
    x = x.astype(np.float64)
    x = x - x.max(axis=-1, keepdims=True)
    return (x - np.log(np.exp(x).sum(axis=-1, keepdims=True))).astype(np.float32)


def triangular_transition(states, width=87.208):
    """The pitch-decoding transition of /root/reference/torbi/evaluate/core.py:23-34:
    clip(w - |i - j|, 0) row-normalised.  w = MAX_OCTAVES_PER_SECOND * HOPSIZE /
    SAMPLE_RATE * (OCTAVE / CENTS_PER_BIN) + 1 with penn's constants
    (35.92, 80, 8000, 1200, 5) -> 87.208.  Returned as PROBABILITIES."""
    i = np.arange(states)
    a = np.clip(np.float32(width) - np.abs(i[:, None] - i[None, :]).astype(np.float32), 0, None)
    a = a.astype(np.float32)
    return a / a.sum(axis=1, keepdims=True, dtype=np.float32)


def log_triangular_transition(states, width=87.208):
    """log(A + tiny), what from_files_to_files passes on (core.py:343-345)."""
    return np.log(triangular_transition(states, width) + np.float32(TINY)).astype(np.float32)


def _noise(rng, B, T, S):
    return log_softmax(rng.standard_normal((B, T, S)))


def _ragged(rng, B, T):
    frames = rng.integers(1, T + 1, size=B).astype(np.int32)
    frames[0] = T
    return frames


def kat():
    # /root/reference/tests/test_core.py:7-25 (README.md:66-95): expects [1, 2, 2]
    return dict(
        observation=np.array([[[0.25, 0.5, 0.25], [0.25, 0.25, 0.5], [0.33, 0.33, 0.33]]], np.float32),
        batch_frames=None,
        transition=np.array([[0.5, 0.25, 0.25], [0.33, 0.34, 0.33], [0.25, 0.25, 0.5]], np.float32),
        initial=np.array([0.4, 0.35, 0.25], np.float32),
        log_probs=False)


def _dense(seed, B, T, S, ragged=False, quant=None, log_probs=True):
    def make():
        rng = np.random.default_rng(seed)
        obs = _noise(rng, B, T, S)
        A = log_softmax(rng.standard_normal((S, S)))
        init = log_softmax(rng.standard_normal((S,)))
        if quant:
            obs = (np.round(obs * quant) / quant).astype(np.float32)
            A = (np.round(A * quant) / quant).astype(np.float32)
            init = (np.round(init * quant) / quant).astype(np.float32)
        frames = _ragged(rng, B, T) if ragged else None
        if not log_probs:
            obs, A, init = np.exp(obs), np.exp(A), np.exp(init)
        return dict(observation=obs, batch_frames=frames, transition=A, initial=init,
                    log_probs=log_probs)
    return make


def _defaults(seed, B, T, S, ragged=False):
    def make():
        rng = np.random.default_rng(seed)
        return dict(observation=_noise(rng, B, T, S),
                    batch_frames=_ragged(rng, B, T) if ragged else None,
                    transition=None, initial=None, log_probs=True)
    return make


def _triangular(seed, B, T, S, width, ragged=False):
    def make():
        rng = np.random.default_rng(seed)
        return dict(observation=_noise(rng, B, T, S),
                    batch_frames=_ragged(rng, B, T) if ragged else None,
                    transition=log_triangular_transition(S, width), initial=None,
                    log_probs=True)
    return make


def _neg_inf(seed, B, T, S):
    def make():
        rng = np.random.default_rng(seed)
        obs = _noise(rng, B, T, S)
        A = log_softmax(rng.standard_normal((S, S)))
        A[rng.random((S, S)) < 0.8] = -np.inf
        A[np.arange(S), np.arange(S)] = np.float32(-1.0)       # keep every state reachable
        return dict(observation=obs, batch_frames=None, transition=A,
                    initial=log_softmax(rng.standard_normal((S,))), log_probs=True)
    return make


def _all_equal(B, T, S):
    def make():
        return dict(observation=np.full((B, T, S), -1.5, np.float32), batch_frames=None,
                    transition=np.full((S, S), -2.0, np.float32),
                    initial=np.full((S,), -0.5, np.float32), log_probs=True)
    return make


def _peaky(seed, B, T, S):
    """Near-one-hot probabilities: exercises the eps floor log(exp(x)+tiny) and
    the prob->log conversion away from the identity region."""
    def make():
        rng = np.random.default_rng(seed)
        logits = rng.standard_normal((B, T, S)) * 30.0
        obs = np.exp(log_softmax(logits).astype(np.float64)).astype(np.float32)
        A = np.exp(log_softmax(rng.standard_normal((S, S)) * 8.0).astype(np.float64)).astype(np.float32)
        init = np.exp(log_softmax(rng.standard_normal((S,))).astype(np.float64)).astype(np.float32)
        return dict(observation=obs, batch_frames=_ragged(rng, B, T), transition=A,
                    initial=init, log_probs=False)
    return make


def _drift(seed, B, T, S):
    """Large-magnitude scores: |P| grows ~ linearly in t so fp32 rounding ties
    are common and the lowest-index tie rule decides the path."""
    def make():
        rng = np.random.default_rng(seed)
        obs = (_noise(rng, B, T, S) - np.float32(4000.0)).astype(np.float32)
        return dict(observation=obs, batch_frames=None,
                    transition=log_triangular_transition(S, 9.5), initial=None, log_probs=True)
    return make


CASES = {
    'kat': kat,
    'dense_s5': _dense(1, 3, 17, 5, ragged=True),
    'dense_s1': _dense(2, 2, 9, 1, ragged=True),
    'dense_s2': _dense(3, 2, 12, 2),
    'dense_s31': _dense(4, 2, 20, 31, ragged=True),
    'dense_s32': _dense(5, 2, 20, 32),
    'dense_s33': _dense(6, 3, 20, 33, ragged=True),
    'dense_s100_prob': _dense(7, 2, 25, 100, ragged=True, log_probs=False),
    'ties_s33': _dense(8, 4, 30, 33, ragged=True, quant=2),
    'ties_s160': _dense(9, 5, 24, 160, ragged=True, quant=1),
    'all_equal_s70': _all_equal(2, 8, 70),
    'defaults_s64': _defaults(10, 3, 40, 64, ragged=True),
    'triangular_s96': _triangular(11, 3, 50, 96, 6.5, ragged=True),
    'neg_inf_s40': _neg_inf(12, 2, 30, 40),
    't1_s17': _dense(13, 3, 1, 17),
    'frames1_s20': lambda: dict(_dense(14, 3, 6, 20)(), batch_frames=np.array([6, 1, 2], np.int32)),
    'peaky_s48': _peaky(15, 3, 30, 48),
    'drift_s128': _drift(16, 2, 60, 128),
    'dense_s257': _dense(17, 9, 14, 257, ragged=True),
    'c1_s1440_defaults': _defaults(1234, 1, 48, 1440),
    'c2_s1440_triangular': _triangular(1235, 5, 40, 1440, 87.208, ragged=True),
    'ties_s1440': _dense(18, 3, 12, 1440, ragged=True, quant=1),
    'dense_s640': _dense(19, 70, 10, 640, ragged=True),
}


def digest(case):
    h = hashlib.sha256()
    for key in ('observation', 'batch_frames', 'transition', 'initial'):
        v = case[key]
        h.update(key.encode())
        if v is None:
            h.update(b'none')
        else:
            v = np.ascontiguousarray(v)
            h.update(str(v.dtype).encode() + str(v.shape).encode() + v.tobytes())
    h.update(str(bool(case['log_probs'])).encode())
    return h.hexdigest()


# ---------------------------------------------------------------------------
# Entropy chunking (reference: /root/reference/torbi/chunk.py:12-85).  Each case is
# (observation (frames, states) fp32 LOG probabilities, [(min_chunk_size, threshold), ...]).
# ---------------------------------------------------------------------------

def _chunk_case(seed, frames, states, peaky_share, gain, settings):
    def make():
        rng = np.random.default_rng(seed)
        logits = rng.standard_normal((frames, states))
        peaky = rng.random(frames) < peaky_share
        logits[peaky] *= gain
        return log_softmax(logits), settings
    return make


def _chunk_threshold_hugger(seed, frames, states, threshold):
    """Frames whose normalised entropy sits within a few fp32 ulp of the threshold: a two-point
    mixture (one state with mass p, the rest uniform) with p bisected so that H(p) ~ threshold,
    then nudged frame by frame.  Any evaluation that rounds differently from the reference's
    torch CPU ops flips decisions here."""
    def make():
        rng = np.random.default_rng(seed)
        lo, hi = 1.0 / states, 1.0 - 1e-9
        def H(p):
            q = (1.0 - p) / (states - 1)
            return -(p * np.log(p) + (states - 1) * q * np.log(q)) / np.log(states)
        for _ in range(200):
            mid = 0.5 * (lo + hi)
            if H(mid) > threshold:
                lo = mid
            else:
                hi = mid
        p = 0.5 * (lo + hi) * (1.0 + rng.uniform(-3e-7, 3e-7, size=frames))
        obs = np.empty((frames, states), np.float64)
        obs[:] = np.log((1.0 - p) / (states - 1))[:, None]
        obs[np.arange(frames), rng.integers(0, states, frames)] = np.log(p)
        return obs.astype(np.float32), [(4, threshold), (1, threshold), (37, threshold)]
    return make


CHUNK_CASES = {
    'chunk_s24': _chunk_case(101, 400, 24, 0.4, 12.0, [(8, 0.5), (25, 0.5), (3, 0.2), (500, 0.5), (1, 0.5)]),
    'chunk_s1440': _chunk_case(102, 3000, 1440, 0.5, 9.0, [(64, 0.5), (200, 0.35), (16, 0.7)]),
    'chunk_s5_short': _chunk_case(103, 7, 5, 0.5, 20.0, [(2, 0.5), (1, 0.8), (2, 0.8), (7, 0.5), (8, 0.5)]),
    'chunk_hugger_s64': _chunk_threshold_hugger(104, 2000, 64, 0.5),
    'chunk_hugger_s1440': _chunk_threshold_hugger(105, 1500, 1440, 0.5),
}


def chunk_digest(observation, settings):
    h = hashlib.sha256()
    observation = np.ascontiguousarray(observation)
    h.update(str(observation.dtype).encode() + str(observation.shape).encode() + observation.tobytes())
    h.update(repr([(int(m), float(t)) for m, t in settings]).encode())
    return h.hexdigest()
