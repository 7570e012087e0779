"""TEST INFRASTRUCTURE ONLY -- Python face of the CPU oracle.

Three interchangeable checkers for the torbi decode path, from most to least
authoritative:

  ref_decode          the UNMODIFIED reference operator torch.ops.torbi.viterbi_decode
                      compiled from /root/reference into oracle/_ref (build_ref.py)
  c_decode            oracle/viterbi_oracle.c -- plain-C restatement of
                      /root/reference/torbi/csrc/viterbi.cpp:35-234 with int64 offsets
  numpy_decode        vectorised numpy restatement (small cases, reads like the maths)

plus `ref_from_probabilities`, a restatement of the host-side preparation in
/root/reference/torbi/core.py:146-206 (defaults, log-space conversion, eps floor)
that feeds any of the three decoders.

Parity status: PINNED -- see the header of viterbi_oracle.c and tests/test_oracle.py.

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
may import this module.  torbi_b200/ never does.
"""
import ctypes
import math
import pathlib
import subprocess

import numpy as np

HERE = pathlib.Path(__file__).resolve().parent
LIB = HERE / '_build' / 'libtorbi_oracle.so'

_lib = None


def build(force: bool = False) -> pathlib.Path:
    """gcc-compile viterbi_oracle.c (seconds)."""
    if force or not LIB.exists() or \
            LIB.stat().st_mtime < (HERE / 'viterbi_oracle.c').stat().st_mtime:
        subprocess.run(['make', '-C', str(HERE), '-B', '_build/libtorbi_oracle.so'],
                       check=True, capture_output=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        lib = ctypes.CDLL(str(LIB))
        lib.torbi_oracle_decode.restype = ctypes.c_int
        lib.torbi_oracle_decode.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int64, ctypes.c_int64, ctypes.c_int64,
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_int, ctypes.c_int]
        lib.torbi_oracle_max_threads.restype = ctypes.c_int
        _lib = lib
    return _lib


def max_threads() -> int:
    return _load().torbi_oracle_max_threads()


def _as(a, dtype):
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def c_decode(observation, batch_frames, transition, initial, *, exact_scan=True,
             num_threads=0, want_posterior=False, want_trellis=False):
    """viterbi_decode_cpu (viterbi.cpp:182-234) via the C restatement.

    observation (B,T,S) fp32 log-probs; batch_frames (B,) int; transition (S,S)
    [dest][src]; initial (S,).  Returns indices (B,T) int32, optionally also the
    final posterior (B,S) fp32 and the int32 trellis (B,T,S).
    """
    obs = _as(observation, np.float32)
    B, T, S = obs.shape
    frames = _as(batch_frames, np.int32)
    A = _as(transition, np.float32)
    init = _as(initial, np.float32)
    assert A.shape == (S, S) and init.shape == (S,) and frames.shape == (B,)
    indices = np.empty((B, T), np.int32)
    posterior = np.empty((B, S), np.float32) if want_posterior else None
    trellis = np.empty((B, T, S), np.int32) if want_trellis else None
    rc = _load().torbi_oracle_decode(
        obs.ctypes.data, frames.ctypes.data, A.ctypes.data, init.ctypes.data,
        B, T, S, indices.ctypes.data,
        posterior.ctypes.data if want_posterior else None,
        trellis.ctypes.data if want_trellis else None,
        1 if exact_scan else 0, int(num_threads))
    if rc != 0:
        raise ValueError('torbi_oracle_decode: bad arguments (frames must be in [1, T])')
    out = (indices,)
    if want_posterior:
        out += (posterior,)
    if want_trellis:
        out += (trellis,)
    return out if len(out) > 1 else indices


def numpy_decode(observation, batch_frames, transition, initial, *, want_posterior=False):
    """Vectorised restatement of viterbi.cpp:35-160,218-221 for small cases.

    max over a NaN-free set is order independent and np.argmax returns the first
    occurrence, which is the reference's strict-`>` ascending scan from a
    zero-filled trellis (viterbi.cpp:94-101).
    """
    obs = _as(observation, np.float32)
    B, T, S = obs.shape
    frames = _as(batch_frames, np.int64)
    A = _as(transition, np.float32)
    init = _as(initial, np.float32)
    indices = np.zeros((B, T), np.int32)
    posterior = np.zeros((B, S), np.float32)
    for b in range(B):
        P = obs[b, 0] + init                                   # viterbi.cpp:72-76
        bp = np.zeros((int(frames[b]), S), np.int32)
        for t in range(1, int(frames[b])):
            cand = P[None, :] + A                              # [dest j, src s]  :81-86
            bp[t] = cand.argmax(axis=1)                        # :91-101
            P = obs[b, t] + cand.max(axis=1)                   # :102
        posterior[b] = P                                       # :111-113
        idx = int(P.argmax())                                  # :218
        indices[b, :] = idx                                    # :219-221 (repeat)
        for t in range(int(frames[b]) - 1, 0, -1):             # :140-160
            idx = int(bp[t, idx])
            indices[b, t - 1] = idx
    return (indices, posterior) if want_posterior else indices


def ref_available() -> bool:
    from . import build_ref
    return build_ref.available()


def ref_decode(observation, batch_frames, transition, initial, *, num_threads=None,
               want_debug=False):
    """The unmodified reference CPU op (oracle/_ref).  Inputs as torch CPU tensors
    or arrays; B*T*S must stay below 2^31 (reference int offsets)."""
    import torch
    from . import build_ref
    ops = build_ref.load_ref()
    obs = torch.as_tensor(np.asarray(observation), dtype=torch.float32).contiguous()
    B, T, S = obs.shape
    assert B * T * S < 2 ** 31, 'reference op overflows int offsets at this size'
    frames = torch.as_tensor(np.asarray(batch_frames), dtype=torch.int32)
    A = torch.as_tensor(np.asarray(transition), dtype=torch.float32)
    init = torch.as_tensor(np.asarray(initial), dtype=torch.float32)
    if num_threads:
        torch.set_num_threads(int(num_threads))                # viterbi.py:51-52
    indices = ops.viterbi_decode(obs, frames, A, init).numpy()
    if want_debug:
        posterior, trellis = ops.debug_trellis(obs, frames, A, init)
        return indices, posterior.numpy(), trellis.numpy()
    return indices


def ref_prepare(observation, batch_frames=None, transition=None, initial=None,
                log_probs=False):
    """Host-side preparation of core.py:146-197 restated with the same torch CPU
    ops in the same order (torch.log on the input dtype, cast to fp32, then the
    in-place exp_/+= tiny/log_ eps floor).  Returns fp32 torch CPU tensors
    (observation, batch_frames int32, transition, initial) ready for the op.
    The caller's observation is NOT mutated (the reference mutates it in place
    when it is already fp32, core.py:191-197)."""
    import torch
    observation = torch.as_tensor(observation)
    batch, frames, states = observation.shape                  # core.py:146
    if batch_frames is None:                                   # :152-158
        batch_frames = torch.full((batch,), frames, dtype=torch.int32)
    batch_frames = torch.as_tensor(batch_frames).to(dtype=torch.int32)
    tiny = torch.finfo(torch.float32).tiny
    if initial is None:                                        # :161-166
        initial = torch.full((states,), math.log((1. / states) + tiny), dtype=torch.float32)
    else:                                                      # :169-172
        initial = torch.as_tensor(initial)
        if not log_probs:
            initial = torch.log(initial)
    if transition is None:                                     # :175-180
        transition = torch.full((states, states), math.log(1. / states), dtype=torch.float32)
    else:                                                      # :183-186
        transition = torch.as_tensor(transition)
        if not log_probs:
            transition = torch.log(transition)
    if not log_probs:                                          # :189-190
        observation = torch.log(observation)
    observation = observation.to(dtype=torch.float32).clone()  # :191
    torch.exp_(observation)                                    # :195
    observation += tiny                                        # :196
    torch.log_(observation)                                    # :197
    return observation, batch_frames, transition, initial


def ref_from_probabilities(observation, batch_frames=None, transition=None, initial=None,
                           log_probs=False, decoder='c', **kw):
    """core.py:110-208 on CPU: ref_prepare + one of the three decoders."""
    obs, frames, A, init = ref_prepare(observation, batch_frames, transition, initial, log_probs)
    fn = {'c': c_decode, 'numpy': numpy_decode, 'ref': ref_decode}[decoder]
    return fn(obs.numpy(), frames.numpy(), A.numpy(), init.numpy(), **kw)
