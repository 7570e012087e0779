"""Native file I/O for the file API: the torch.save containers of
/root/reference/torbi/core.py:244 (inputs) and :466-473 (outputs) read and written by
host threads inside libtorbi_b200.so (csrc/ptfile.cpp) instead of one Python thread.
Files the fast path does not cover fall back to torch.load, one by one."""
import ctypes
import os
from typing import List, Sequence, Tuple

import torch

from .. import _ffi

DTYPES = {_ffi.PT_F32: torch.float32, _ffi.PT_F16: torch.float16, _ffi.PT_BF16: torch.bfloat16,
          _ffi.PT_F64: torch.float64, _ffi.PT_I32: torch.int32, _ffi.PT_I64: torch.int64}


def _paths(files) -> Tuple[ctypes.Array, list]:
    encoded = [os.fsencode(os.fspath(file)) for file in files]
    return (ctypes.c_char_p * len(encoded))(*encoded), encoded


def inspect(files: Sequence, threads: int = 0):
    """One _ffi.PtInfo per file (status != 0: not readable by the fast path)."""
    lib = _ffi.library()
    array, _keep = _paths(files)
    infos = (_ffi.PtInfo * len(files))()
    status = lib.torbi_b200_pt_inspect(array, len(files), infos, threads)
    if status < 0:
        _ffi.check(status)
    return infos


def load_batch(files: Sequence, pin_memory: bool = False, threads: int = 0, out: torch.Tensor = None):
    """Zero-padded (B, Tmax, S) batch of 2-D posteriorgram files plus their frame counts -- what
    Dataset + collate() produce (data/dataset.py:17-22, collate.py:19-33), assembled by the native
    reader in the files' own dtype: fp32, or fp16 / bf16 (kept at 2 bytes per element all the way
    to the kernel's observation load).  `out`: a flat (pinned) fp32 buffer to carve the batch
    from, if big enough.  Returns (observation, batch_frames) or None when the batch needs the
    generic path (files of mixed or other dtypes, mixed widths, or not 2-D)."""
    infos = inspect(files, threads)
    fallback = [i for i, info in enumerate(infos) if info.status != _ffi.PT_OK]
    loaded = {i: torch.load(files[i], map_location='cpu') for i in fallback}
    shapes, dtypes = [], set()
    for i, info in enumerate(infos):
        if i in loaded:
            shapes.append(tuple(loaded[i].shape))
            dtypes.add(loaded[i].dtype)
        else:
            shapes.append(tuple(info.shape[:info.ndim]))
            dtypes.add(DTYPES[info.dtype])
    if len(dtypes) != 1 or not dtypes <= {torch.float32, torch.float16, torch.bfloat16} or \
            any(len(shape) != 2 for shape in shapes) or len({shape[1] for shape in shapes}) != 1:
        return None
    dtype = next(iter(dtypes))
    itemsize = 4 if dtype == torch.float32 else 2
    states = shapes[0][1]
    batch_frames = torch.tensor([shape[0] for shape in shapes])
    max_frames = int(batch_frames.max())
    numel = len(files) * max_frames * states
    if out is not None and out.numel() * out.element_size() >= numel * itemsize:
        observation = out.view(dtype)[:numel].view(len(files), max_frames, states)
    else:
        observation = torch.empty((len(files), max_frames, states), dtype=dtype, pin_memory=pin_memory)
    native = [i for i in range(len(files)) if i not in loaded]
    if native:
        array, _keep = _paths([files[i] for i in native])
        native_infos = (_ffi.PtInfo * len(native))(*[infos[i] for i in native])
        row_bytes = max_frames * states * itemsize
        destinations = (ctypes.c_void_p * len(native))(*[observation.data_ptr() + i * row_bytes for i in native])
        capacities = (ctypes.c_int64 * len(native))(*([row_bytes] * len(native)))
        failed = _ffi.library().torbi_b200_pt_read(array, native_infos, len(native), destinations, capacities, threads)
        if failed:
            raise OSError(f'torbi_b200: reading {failed} of {len(native)} files failed')
    for i, tensor in loaded.items():
        observation[i, :tensor.shape[0]] = tensor
        observation[i, tensor.shape[0]:] = 0
    return observation, batch_frames


def load_batch_packed(files: Sequence, pin_memory: bool = False, threads: int = 0, out: torch.Tensor = None):
    """The files of a batch back to back in ONE (rows, S) buffer -- no zero padding to the longest
    file (the reference pads, collate.py:26-31): returns (observation, item_row, batch_frames) with
    file i in rows item_row[i] .. item_row[i] + batch_frames[i], in the files' own dtype (fp32, fp16
    or bf16), or None when the batch needs the generic path (see load_batch).  This is the layout
    torbi_b200.decode_packed / decode_host_packed take."""
    infos = inspect(files, threads)
    shapes, dtypes = [], set()
    for info in infos:
        if info.status != _ffi.PT_OK:
            return None
        shapes.append(tuple(info.shape[:info.ndim]))
        dtypes.add(DTYPES[info.dtype])
    if len(dtypes) != 1 or not dtypes <= {torch.float32, torch.float16, torch.bfloat16} or \
            any(len(shape) != 2 or shape[0] < 1 for shape in shapes) or len({shape[1] for shape in shapes}) != 1:
        return None
    dtype = next(iter(dtypes))
    itemsize = 4 if dtype == torch.float32 else 2
    states = shapes[0][1]
    batch_frames = torch.tensor([shape[0] for shape in shapes], dtype=torch.int32)
    item_row = torch.zeros(len(files), dtype=torch.int64)
    item_row[1:] = torch.cumsum(batch_frames.to(torch.int64), 0)[:-1]
    rows = int(batch_frames.sum())
    if out is not None and out.numel() * out.element_size() >= rows * states * itemsize:
        observation = out.view(dtype)[:rows * states].view(rows, states)
    else:
        observation = torch.empty((rows, states), dtype=dtype, pin_memory=pin_memory)
    array, _keep = _paths(files)
    row_bytes = states * itemsize
    destinations = (ctypes.c_void_p * len(files))(*[observation.data_ptr() + int(r) * row_bytes for r in item_row])
    capacities = (ctypes.c_int64 * len(files))(*[int(f) * row_bytes for f in batch_frames])
    failed = _ffi.library().torbi_b200_pt_read(array, infos, len(files), destinations, capacities, threads)
    if failed:
        raise OSError(f'torbi_b200: reading {failed} of {len(files)} files failed')
    return observation, item_row, batch_frames


def save_int32(tensors: List[torch.Tensor], files: Sequence, threads: int = 0):
    """torch.save() of contiguous CPU int32 tensors (core.py:466-473), written by native threads."""
    tensors = [tensor.contiguous() for tensor in tensors]
    for tensor in tensors:
        if tensor.dtype != torch.int32 or tensor.device.type != 'cpu' or tensor.dim() > 4:
            raise ValueError('save_int32 takes CPU int32 tensors of at most 4 dimensions')
    array, _keep = _paths(files)
    data = (ctypes.c_void_p * len(tensors))(*[tensor.data_ptr() for tensor in tensors])
    shapes = (ctypes.c_int64 * (4 * len(tensors)))()
    ndims = (ctypes.c_int32 * len(tensors))()
    for i, tensor in enumerate(tensors):
        ndims[i] = tensor.dim()
        for d, size in enumerate(tensor.shape):
            shapes[4 * i + d] = size
    failed = _ffi.library().torbi_b200_pt_write_i32(array, data, shapes, ndims, len(tensors), threads)
    if failed:
        raise OSError(f'torbi_b200: writing {failed} of {len(tensors)} files failed')
