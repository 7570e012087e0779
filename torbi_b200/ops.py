"""PyTorch dispatcher front for the C ABI: registers

    torbi_b200::viterbi_decode(Tensor observation, Tensor batch_frames,
                               Tensor transition, Tensor initial) -> Tensor

with the schema of the reference operator (/root/reference/torbi/csrc/ops.cpp:16-18)
and a CUDA implementation that calls libtorbi_b200.so, so call sites written
against `torch.ops.torbi.viterbi_decode` port by changing the namespace.  There
is deliberately no CPU / MPS registration (the reference registers both,
viterbi.cpp:237-239, viterbi.mm:288-290): dispatching a CPU tensor raises.
A fake (meta) kernel makes the op traceable by torch.compile / export.
"""
import torch

from .viterbi import decode

_LIBRARY = torch.library.Library('torbi_b200', 'DEF')
_LIBRARY.define(
    'viterbi_decode(Tensor observation, Tensor batch_frames, Tensor transition, Tensor initial) -> Tensor')


def _viterbi_decode_cuda(observation, batch_frames, transition, initial):
    return decode(observation, batch_frames, transition, initial)


def _viterbi_decode_meta(observation, batch_frames, transition, initial):
    return observation.new_empty(observation.shape[:2], dtype=torch.int32)


_LIBRARY.impl('viterbi_decode', _viterbi_decode_cuda, 'CUDA')
_LIBRARY.impl('viterbi_decode', _viterbi_decode_meta, 'Meta')
