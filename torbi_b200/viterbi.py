"""Operator boundary: `decode` mirrors /root/reference/torbi/viterbi.py:5-53 but
calls the C ABI of libtorbi_b200.so instead of torch.ops.torbi.viterbi_decode."""
from typing import Optional, Sequence, Union

import torch

from . import _ffi


def _host_frames(batch_frames: torch.Tensor, batch: int) -> torch.Tensor:
    # The C ABI takes batch_frames in host memory (it schedules by length).  A
    # CUDA batch_frames costs one small device->host copy.
    frames = batch_frames.detach().to(device='cpu', dtype=torch.int32).contiguous()
    if frames.dim() != 1 or frames.shape[0] != batch:
        raise ValueError(
            f'batch_frames must have shape ({batch},), got {tuple(batch_frames.shape)}')
    return frames


def _check_inputs(observation, transition, initial):
    if observation.dim() != 3:
        raise ValueError(
            f'observation must have shape (batch, frames, states), got {tuple(observation.shape)}')
    batch, frames, states = observation.shape
    if tuple(transition.shape) != (states, states):
        raise ValueError(f'transition must have shape ({states}, {states}), got {tuple(transition.shape)}')
    if tuple(initial.shape) != (states,):
        raise ValueError(f'initial must have shape ({states},), got {tuple(initial.shape)}')
    for name, tensor in (('transition', transition), ('initial', initial)):
        # the reference op raises on anything but fp32 (data_ptr<float>(), viterbi.cpp:46-50)
        if tensor.dtype != torch.float32:
            raise ValueError(f'{name} must be float32, got {tensor.dtype}')
    # ... while the reference's from_probabilities casts the observation from any dtype (core.py:191):
    # fp16 / bf16 observations are taken as they are and upcast inside the kernels
    if observation.dtype not in _OBS_FLAGS:
        raise ValueError(f'observation must be float32, float16 or bfloat16, got {observation.dtype}')
    return batch, frames, states


_OBS_FLAGS = {torch.float32: 0, torch.float16: _ffi.OBS_F16, torch.bfloat16: _ffi.OBS_BF16}


def _sm_count(device) -> int:
    return torch.cuda.get_device_properties(device).multi_processor_count


def decode(
    observation: torch.Tensor,
    batch_frames: torch.Tensor,
    transition: torch.Tensor,
    initial: torch.Tensor,
    num_threads: int = 0,
    *,
    flags: int = 0,
    return_posterior: bool = False,
    checkpoint_every: Optional[int] = None,
) -> torch.Tensor:
    """Decode a time-varying categorical distribution

    Args:
        observation: :math:`(N, T, S)` float32 (or float16 / bfloat16) CUDA tensor,
            log-probabilities (or as described by `flags`)
        batch_frames :math:`(N)`
            Sequence length of each batch item
        transition :math:`(S, S)`
            Log transition matrix, indexed [destination][source]
        initial :math:`(S)`
            Log initial distribution
        num_threads
            Ignored (CPU thread count in the reference, viterbi.py:51-52)
        flags
            Bit-or of the TORBI_B200_* flags of include/torbi_b200.h (fused
            log-space conversion, kernel selection)
        return_posterior
            Also return the final forward scores :math:`(N, S)` float32
        checkpoint_every
            Also return the forward scores :math:`P_t` for t = k * checkpoint_every,
            shape :math:`(N, (T - 1) // checkpoint_every, S)` float32 (rows at or past an
            item's length are NaN) -- what lets a long decode be verified in segments

    Return:
        indices: :math:`(N, T)` int32 on the device of `observation`
        (then the posterior and / or the checkpoints when requested)
    """
    del num_threads
    batch, frames, states = _check_inputs(observation, transition, initial)
    if observation.device.type != 'cuda':
        raise RuntimeError(
            'torbi_b200.decode needs CUDA tensors: this build has no CPU decode path '
            '(use torbi_b200.decode_host / from_probabilities(gpu=...) for host buffers)')
    device = observation.device
    if transition.device != device or initial.device != device:
        raise RuntimeError('observation, transition and initial must be on the same device')
    lib = _ffi.library()
    observation = observation.contiguous()
    transition = transition.contiguous()
    initial = initial.contiguous()
    host_frames = _host_frames(batch_frames, batch)
    flags |= _OBS_FLAGS[observation.dtype]
    with torch.cuda.device(device):
        indices = torch.empty((batch, frames), dtype=torch.int32, device=device)
        posterior = torch.empty((batch, states), dtype=torch.float32, device=device) \
            if return_posterior else None
        # sized for the kernel this call will use: uint16 backpointers for the wide / generic
        # kernels (small batches, one long sequence), fp32 score history otherwise
        workspace_bytes = lib.torbi_b200_workspace_bytes_planned(
            host_frames.data_ptr(), batch, frames, states, flags, _sm_count(device))
        if workspace_bytes == 0:
            _ffi.check(_ffi.ERR_INVALID_ARGUMENT)
        workspace = torch.empty(workspace_bytes, dtype=torch.uint8, device=device)
        stream = torch.cuda.current_stream(device)
        args = (observation.data_ptr(), host_frames.data_ptr(), transition.data_ptr(),
                initial.data_ptr(), batch, frames, states, flags, indices.data_ptr(),
                posterior.data_ptr() if return_posterior else None,
                workspace.data_ptr(), workspace_bytes, stream.cuda_stream)
        checkpoints = None
        if checkpoint_every is None:
            _ffi.check(lib.torbi_b200_decode(*args))
        else:
            checkpoints = torch.full((batch, max(0, (frames - 1) // max(1, checkpoint_every)), states),
                                     float('nan'), dtype=torch.float32, device=device)
            _ffi.check(lib.torbi_b200_decode_checkpointed(*args, checkpoints.data_ptr(), checkpoint_every))
        # the workspace goes back to torch's caching allocator here; it is only
        # handed out again in stream order, behind the kernels just enqueued
    out = (indices,) + ((posterior,) if return_posterior else ()) + ((checkpoints,) if checkpoints is not None else ())
    return out if len(out) > 1 else indices


def decode_host(
    observation: torch.Tensor,
    batch_frames: torch.Tensor,
    transition: torch.Tensor,
    initial: torch.Tensor,
    gpu: Union[int, Sequence[int]] = 0,
    *,
    flags: int = 0,
    return_posterior: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Decode HOST tensors through torbi_b200_decode_host[_multi]: the
    host->device copies, the decode and the device->host copy of the indices all
    happen inside the C call.  `gpu` is one ordinal or a sequence of ordinals
    (items are sharded longest-first across them, no collective).  Pinned
    tensors copy at full PCIe rate.  Returns int32 (N, T) on the CPU.
    """
    batch, frames, states = _check_inputs(observation, transition, initial)
    for name, tensor in (('observation', observation), ('transition', transition), ('initial', initial)):
        if tensor.device.type != 'cpu':
            raise RuntimeError(f'decode_host: {name} must be a CPU tensor')
    lib = _ffi.library()
    observation = observation.contiguous()
    transition = transition.contiguous()
    initial = initial.contiguous()
    host_frames = _host_frames(batch_frames, batch)
    flags |= _OBS_FLAGS[observation.dtype]
    if out is None:
        out = torch.empty((batch, frames), dtype=torch.int32)
    elif out.dtype != torch.int32 or tuple(out.shape) != (batch, frames) or not out.is_contiguous() \
            or out.device.type != 'cpu':
        raise ValueError('out must be a contiguous CPU int32 tensor of shape (batch, frames)')
    posterior = torch.empty((batch, states), dtype=torch.float32) if return_posterior else None
    post_ptr = posterior.data_ptr() if return_posterior else None
    if isinstance(gpu, int):
        status = lib.torbi_b200_decode_host(
            observation.data_ptr(), host_frames.data_ptr(), transition.data_ptr(),
            initial.data_ptr(), batch, frames, states, flags, out.data_ptr(), post_ptr, gpu)
    else:
        devices = torch.tensor(list(gpu), dtype=torch.int32)
        status = lib.torbi_b200_decode_host_multi(
            observation.data_ptr(), host_frames.data_ptr(), transition.data_ptr(),
            initial.data_ptr(), batch, frames, states, flags, out.data_ptr(), post_ptr,
            devices.data_ptr(), devices.numel())
    _ffi.check(status)
    return (out, posterior) if return_posterior else out


def _packed_arguments(observation, item_row, batch_frames, transition, initial):
    if observation.dim() != 2:
        raise ValueError(f'observation must have shape (rows, states), got {tuple(observation.shape)}')
    rows, states = observation.shape
    _check_inputs(observation[None], transition, initial)
    frames = batch_frames.detach().to(device='cpu', dtype=torch.int32).contiguous()
    first = torch.as_tensor(item_row).detach().to(device='cpu', dtype=torch.int64).contiguous()
    if frames.dim() != 1 or first.shape != frames.shape:
        raise ValueError('item_row and batch_frames must be 1-D and of equal length')
    return rows, states, first, frames


def decode_packed(
    observation: torch.Tensor,
    item_row: torch.Tensor,
    batch_frames: torch.Tensor,
    transition: torch.Tensor,
    initial: torch.Tensor,
    *,
    flags: int = 0,
    return_posterior: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Decode a PACKED batch (torbi_b200_decode_packed): `observation` is one (rows, S) CUDA
    buffer, item b is its rows item_row[b] .. item_row[b] + batch_frames[b]; returns (rows,)
    int32 with every item's states in its own rows.  No padding to the longest item
    (the reference pads, data/collate.py:26-31) -- the pieces of a chunked sequence or the
    files of a batch are decoded where they lie."""
    rows, states, first, frames = _packed_arguments(observation, item_row, batch_frames, transition, initial)
    if observation.device.type != 'cuda':
        raise RuntimeError('torbi_b200.decode_packed needs CUDA tensors (decode_host_packed takes host buffers)')
    device = observation.device
    if transition.device != device or initial.device != device:
        raise RuntimeError('observation, transition and initial must be on the same device')
    lib = _ffi.library()
    observation, transition, initial = observation.contiguous(), transition.contiguous(), initial.contiguous()
    flags |= _OBS_FLAGS[observation.dtype]
    batch = frames.numel()
    with torch.cuda.device(device):
        if out is None:
            out = torch.zeros((rows,), dtype=torch.int32, device=device)
        elif out.dtype != torch.int32 or tuple(out.shape) != (rows,) or not out.is_contiguous() or out.device != device:
            raise ValueError('out must be a contiguous int32 tensor of shape (rows,) on the device of observation')
        posterior = torch.empty((batch, states), dtype=torch.float32, device=device) if return_posterior else None
        workspace_bytes = lib.torbi_b200_workspace_bytes_packed(
            first.data_ptr(), frames.data_ptr(), batch, rows, states, flags, _sm_count(device))
        if workspace_bytes == 0:
            _ffi.check(_ffi.ERR_INVALID_ARGUMENT)
        workspace = torch.empty(workspace_bytes, dtype=torch.uint8, device=device)
        _ffi.check(lib.torbi_b200_decode_packed(
            observation.data_ptr(), first.data_ptr(), frames.data_ptr(), transition.data_ptr(), initial.data_ptr(),
            batch, rows, states, flags, out.data_ptr(), posterior.data_ptr() if return_posterior else None,
            workspace.data_ptr(), workspace_bytes, torch.cuda.current_stream(device).cuda_stream))
    return (out, posterior) if return_posterior else out


def decode_host_packed(
    observation: torch.Tensor,
    item_row: torch.Tensor,
    batch_frames: torch.Tensor,
    transition: torch.Tensor,
    initial: torch.Tensor,
    gpu: int = 0,
    *,
    flags: int = 0,
    return_posterior: bool = False,
    out: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """decode_packed for HOST tensors (torbi_b200_decode_host_packed): copies, decode and the copy
    back of the (rows,) indices inside the C call."""
    rows, states, first, frames = _packed_arguments(observation, item_row, batch_frames, transition, initial)
    for name, tensor in (('observation', observation), ('transition', transition), ('initial', initial)):
        if tensor.device.type != 'cpu':
            raise RuntimeError(f'decode_host_packed: {name} must be a CPU tensor')
    lib = _ffi.library()
    observation, transition, initial = observation.contiguous(), transition.contiguous(), initial.contiguous()
    flags |= _OBS_FLAGS[observation.dtype]
    batch = frames.numel()
    if out is None:
        out = torch.zeros((rows,), dtype=torch.int32)
    elif out.dtype != torch.int32 or tuple(out.shape) != (rows,) or not out.is_contiguous() or out.device.type != 'cpu':
        raise ValueError('out must be a contiguous CPU int32 tensor of shape (rows,)')
    posterior = torch.empty((batch, states), dtype=torch.float32) if return_posterior else None
    _ffi.check(lib.torbi_b200_decode_host_packed(
        observation.data_ptr(), first.data_ptr(), frames.data_ptr(), transition.data_ptr(), initial.data_ptr(),
        batch, rows, states, flags, out.data_ptr(), posterior.data_ptr() if return_posterior else None, gpu))
    return (out, posterior) if return_posterior else out
