"""Tensor API and file / batch drivers with the signatures of
/root/reference/torbi/core.py:110-473, on top of the sm_100a decoder.

Differences from the reference, all on purpose:
  * no CPU / MPS path: `gpu=None` works only for CUDA inputs, else raises;
  * `gpu` may also be a sequence of ordinals or 'all': independent items / files
    are sharded across those GPUs, longest first, with no collective beyond
    gathering the int32 index arrays;
  * the log-space conversion of the observation (core.py:189-197) is fused into
    the forward kernel instead of four elementwise passes, and the caller's
    tensor is never mutated (the reference mutates it in place, core.py:195-197);
    `conversion='host'` reproduces the reference's CPU torch ops exactly instead.
"""
import concurrent.futures
import functools
import math
import os
import queue
import threading
from typing import Dict, List, Optional, Sequence, Union

import torch

import torbi_b200
from . import _ffi, timing
from .viterbi import decode, decode_packed

__all__ = [
    'from_probabilities', 'from_file', 'from_file_to_file', 'from_files_to_files',
    'from_dataloader', 'save', 'save_masked', 'release_pinned_buffers']

Gpu = Optional[Union[int, str, Sequence[int]]]


###############################################################################
# Device selection and sharding
###############################################################################


def _resolve_devices(gpu: Gpu, observation: Optional[torch.Tensor] = None) -> List[int]:
    """`gpu` of the reference API (core.py:147-150) -> list of CUDA ordinals."""
    if isinstance(gpu, str):
        if gpu == 'all':
            count = torch.cuda.device_count()
            if count == 0:
                raise RuntimeError("gpu='all' but no CUDA device is visible")
            return list(range(count))
        # the reference accepts gpu='mps' (core.py:147-148)
        raise RuntimeError(f"gpu={gpu!r}: torbi_b200 is CUDA sm_100a only (no MPS / CPU dispatch)")
    if gpu is None:
        if observation is not None and observation.is_cuda:
            return [observation.device.index]
        raise RuntimeError(
            'gpu=None selects the CPU path in torbi; torbi_b200 has no CPU fallback. '
            'Pass gpu=<index> (or a list of indices, or "all").')
    if isinstance(gpu, int):
        return [gpu]
    devices = [int(g) for g in gpu]
    if not devices:
        raise ValueError('gpu must name at least one device')
    return devices


def lpt_shards(costs: Sequence[int], shards: int) -> List[List[int]]:
    """Longest-processing-time-first assignment of items to `shards` bins.
    Returns per-bin item indices in ascending order (original order kept)."""
    shards = max(1, min(shards, len(costs)))
    order = sorted(range(len(costs)), key=lambda i: -int(costs[i]))
    load = [0] * shards
    bins: List[List[int]] = [[] for _ in range(shards)]
    for item in order:
        k = load.index(min(load))
        bins[k].append(item)
        load[k] += int(costs[item])
    return [sorted(b) for b in bins]


def size_capped_batches(sizes: Sequence[int], batch_size: int, max_bytes: int = 1 << 29) -> List[List[int]]:
    """Batches of file indices, largest files first, at most `batch_size` files and at most
    `max_bytes` in total (a single larger file makes a batch of its own): the batching of the packed
    (unpadded) layout, where a batch costs the sum of its files rather than count x longest."""
    order = sorted(range(len(sizes)), key=lambda i: -int(sizes[i]))
    batches: List[List[int]] = []
    total = 0
    for i in order:
        if batches and len(batches[-1]) < batch_size and total + int(sizes[i]) <= max_bytes:
            batches[-1].append(i)
            total += int(sizes[i])
        else:
            batches.append([i])
            total = int(sizes[i])
    return batches


def size_bucketed_batches(sizes: Sequence[int], batch_size: int,
                          max_padded_bytes: int = 2 << 30) -> List[List[int]]:
    """Batches of file indices, largest files first, at most `batch_size` files and
    at most `max_padded_bytes` once zero-padded to the largest member (the first of
    each batch).  Keeps the padded batch of collate() small for long files."""
    order = sorted(range(len(sizes)), key=lambda i: -int(sizes[i]))
    batches: List[List[int]] = []
    for i in order:
        if batches and len(batches[-1]) < batch_size and \
                (len(batches[-1]) + 1) * int(sizes[batches[-1][0]]) <= max_padded_bytes:
            batches[-1].append(i)
        else:
            batches.append([i])
    return batches


###############################################################################
# Tensor API
###############################################################################


def _conversion_mode(conversion: Optional[str]) -> str:
    mode = conversion or os.environ.get('TORBI_B200_CONVERSION', 'fused')
    if mode not in ('fused', 'host', 'torch'):
        raise ValueError(f"conversion must be 'fused', 'host' or 'torch', got {mode!r}")
    return mode


def _prepare_observation(observation, log_probs, device, mode):
    """core.py:189-197.  Returns (fp32 tensor on `device`, kernel flags)."""
    tiny = torch.finfo(torch.float32).tiny
    if mode == 'fused':
        flags = _ffi.EPS_FLOOR
        if not log_probs:
            if observation.dtype == torch.float32:
                flags |= _ffi.OBS_IS_PROBABILITY
            else:
                # the reference takes the log in the INPUT dtype, then casts (core.py:190-191)
                observation = torch.log(observation)
        # fp16 / bf16 observations travel and stay as 2-byte elements; the kernels upcast in their
        # observation load, which is what the reference's .to(float32) does (core.py:191), exactly
        dtype = observation.dtype if observation.dtype in (torch.float16, torch.bfloat16) else torch.float32
        return observation.to(device=device, dtype=dtype), flags
    if mode == 'host':
        # exactly the reference's CPU ops, then one host->device copy
        observation = observation.detach().to('cpu')
        if not log_probs:
            observation = torch.log(observation)
        observation = observation.to(dtype=torch.float32, copy=True)
    else:
        # the reference's own CUDA route: same ops, run by torch on the device
        if not log_probs:
            observation = torch.log(observation)
        observation = observation.to(device=device, dtype=torch.float32, copy=True)
    torch.exp_(observation)
    observation += tiny
    torch.log_(observation)
    return observation.to(device), 0


def _is_banded(transition: torch.Tensor) -> bool:
    """Whether most of the (log) transition matrix equals one constant -- e.g. the pitch
    matrix of evaluate/core.py:23-34, where 88 % of the entries are log(tiny).  Only a
    speed hint: the banded kernel is exact for any matrix."""
    states = transition.shape[0]
    if states < 256:
        return False
    as_bits = transition.contiguous().view(torch.int32)
    return bool((as_bits != as_bits[0, -1]).float().mean() < 0.35)


def _prepare_model(states, batch, transition, initial, log_probs, device, banded):
    """Defaults and log-space conversion of `initial` / `transition` (core.py:161-186) -> fp32 tensors on
    `device` plus the banded hint."""
    tiny = torch.finfo(torch.float32).tiny
    if initial is None:
        initial = torch.full(
            (states,), math.log((1. / states) + tiny), dtype=torch.float32, device=device)
    else:
        if not log_probs:
            initial = torch.log(initial)
        initial = initial.to(device=device, dtype=torch.float32)
    if transition is None:
        transition = torch.full(
            (states, states), math.log(1. / states), dtype=torch.float32, device=device)
    else:
        if not log_probs:
            transition = torch.log(transition)
        transition = transition.to(dtype=torch.float32)
    if banded is None:
        banded = batch >= 32 and _is_banded(transition)
    return transition.to(device=device), initial, banded


def _from_probabilities_packed(observation, item_row, batch_frames, transition, initial, log_probs, device,
                               banded=None, min_chunk_size=None, entropy_threshold=None, stats=None):
    """from_probabilities on a PACKED batch: `observation` (rows, states), item i in rows item_row[i] ..
    + batch_frames[i] (host or device tensor; fp32 / fp16 / bf16).  With `min_chunk_size` every item is
    first cut at its low-entropy frames (chunk.py:12-85, entropy on the device over the whole buffer in
    one pass) and the pieces -- row ranges of the same buffer -- are decoded as independent items, which
    is what the reference's Dataset + collate + separate do (data/dataset.py:21-23, collate.py:13-17,
    36-45).  Returns (rows,) int32 on `device`: every item's states in its own rows."""
    import importlib
    chunk_module = importlib.import_module(__package__ + '.chunk')   # (the package attribute `chunk` is the function)
    rows, states = observation.shape
    frames_list = [int(f) for f in batch_frames]
    rows_list = [int(r) for r in item_row]
    model_transition, model_initial, banded = _prepare_model(
        states, len(frames_list), transition, initial, log_probs, device, banded)
    flags = _ffi.EPS_FLOOR
    if not log_probs:
        if observation.dtype == torch.float32:
            flags |= _ffi.OBS_IS_PROBABILITY
        else:
            # the reference takes the log in the INPUT dtype, then casts (core.py:190-191)
            observation = torch.log(observation)
    raw = observation
    dtype = observation.dtype if observation.dtype in (torch.float16, torch.bfloat16) else torch.float32
    if min_chunk_size is not None and not log_probs and observation.dtype != torch.float32:
        raise ValueError('chunked decoding of non-fp32 probability files is not supported on the packed route')
    observation = observation.to(device=device, dtype=dtype, non_blocking=True)
    if min_chunk_size is not None:
        # the reference chunks the observation exactly as stored (before any conversion)
        del raw
        low = chunk_module.low_entropy_frames_device(
            observation, torbi_b200.ENTROPY_THRESHOLD if entropy_threshold is None else entropy_threshold, stats)
        piece_row, piece_frames = [], []
        for row0, frames in zip(rows_list, frames_list):
            points = chunk_module.split_points_from_candidates(low[row0:row0 + frames], frames, min_chunk_size)
            bounds = [0] + points + [frames]
            for start, stop in zip(bounds[:-1], bounds[1:]):
                piece_row.append(row0 + start)
                piece_frames.append(stop - start)
        rows_list, frames_list = piece_row, piece_frames
        if stats is not None:
            stats['pieces'] = stats.get('pieces', 0) + len(piece_row)
    if banded:
        flags |= _ffi.BANDED
    with timing.context('torbi', device):                    # core.py:200
        return decode_packed(observation, torch.tensor(rows_list, dtype=torch.int64),
                             torch.tensor(frames_list, dtype=torch.int32), model_transition, model_initial, flags=flags)


def _from_probabilities_one(observation, batch_frames, transition, initial, log_probs,
                            device, mode, return_posterior=False, banded=None):
    batch, frames, states = observation.shape
    tiny = torch.finfo(torch.float32).tiny

    # Default to uniform initial probabilities (core.py:161-172)
    if initial is None:
        initial = torch.full(
            (states,), math.log((1. / states) + tiny), dtype=torch.float32, device=device)
    else:
        if not log_probs:
            initial = torch.log(initial)
        initial = initial.to(device=device, dtype=torch.float32)

    # Default to uniform transition probabilities (core.py:175-186)
    if transition is None:
        transition = torch.full(
            (states, states), math.log(1. / states), dtype=torch.float32, device=device)
    else:
        if not log_probs:
            transition = torch.log(transition)
        transition = transition.to(dtype=torch.float32)
    if banded is None:
        # small batches are better served by the wide kernel, which the flag would rule out
        banded = batch >= 32 and _is_banded(transition)
    transition = transition.to(device=device)

    observation, flags = _prepare_observation(observation, log_probs, device, mode)
    if banded:
        flags |= _ffi.BANDED
    with timing.context('torbi', device):                    # core.py:200
        return decode(observation, batch_frames, transition, initial, flags=flags,
                      return_posterior=return_posterior)


def from_probabilities(
    observation: torch.Tensor,
    batch_frames: Optional[torch.Tensor] = None,
    transition: Optional[torch.Tensor] = None,
    initial: Optional[torch.Tensor] = None,
    log_probs: bool = False,
    gpu: Gpu = None,
    num_threads: Optional[int] = 1,
    *,
    conversion: Optional[str] = None,
    banded: Optional[bool] = None,
) -> torch.Tensor:
    """Decode a time-varying categorical distribution

    Arguments
        observation
            Time-varying categorical distribution
            shape=(batch, frames, states)
        batch_frames
            Number of frames in each batch item; defaults to all
            shape=(batch,)
        transition
            Categorical transition matrix; defaults to uniform
            shape=(states, states), indexed [destination][source]
        initial
            Categorical initial distribution; defaults to uniform
            shape=(states,)
        log_probs
            Whether inputs are in (natural) log space
        gpu
            GPU index to use for decoding; a sequence of indices or 'all' shards
            the batch items across those GPUs. None uses the device of a CUDA
            `observation` (there is no CPU path).
        num_threads
            Ignored (CPU threads in the reference)
        conversion
            'fused' (default): log / eps floor computed inside the kernel;
            'host': the reference's exact CPU torch ops; 'torch': torch ops on
            the GPU. Also settable through TORBI_B200_CONVERSION.
        banded
            Use the exact fast path for banded transition matrices (results are
            bit-identical either way). None (default) enables it when most of the
            matrix is one constant.

    Returns
        indices
            The decoded bin indices, int32, on the (first) GPU
            shape=(batch, frames)
    """
    del num_threads
    batch, frames, states = observation.shape           # core.py:146 (ValueError if not 3-D)
    devices = _resolve_devices(gpu, observation)
    mode = _conversion_mode(conversion)

    if batch_frames is None:                             # core.py:152-158
        batch_frames = torch.full((batch,), frames, dtype=torch.int32)
    host_frames = batch_frames.detach().to(device='cpu', dtype=torch.int32)

    if len(devices) == 1 or batch == 1:
        device = torch.device(f'cuda:{devices[0]}')
        return _from_probabilities_one(
            observation, host_frames, transition, initial, log_probs, device, mode, banded=banded)

    # Several GPUs: independent items, longest first, one host thread per GPU;
    # the only cross-device traffic is the gather of the int32 rows.
    shards = lpt_shards(host_frames.tolist(), len(devices))
    home = torch.device(f'cuda:{devices[0]}')
    indices = torch.empty((batch, frames), dtype=torch.int32, device=home)
    errors: List[BaseException] = []

    def work(k, items):
        try:
            device = torch.device(f'cuda:{devices[k]}')
            select = torch.tensor(items, dtype=torch.long)
            part = observation.index_select(0, select.to(observation.device))
            with torch.cuda.device(device):
                rows = _from_probabilities_one(
                    part, host_frames[select], transition, initial, log_probs, device, mode,
                    banded=banded)
                torch.cuda.current_stream(device).synchronize()
            indices.index_copy_(0, select.to(home), rows.to(home))
        except BaseException as error:  # noqa: BLE001 - re-raised on the caller's thread
            errors.append(error)

    threads = [threading.Thread(target=work, args=(k, items)) for k, items in enumerate(shards)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    if errors:
        raise errors[0]
    return indices


def from_file(
    input_file: Union[str, os.PathLike],
    transition_file: Optional[Union[str, os.PathLike]] = None,
    initial_file: Optional[Union[str, os.PathLike]] = None,
    log_probs: bool = False,
    gpu: Gpu = None,
    num_threads: Optional[int] = 1
) -> torch.Tensor:
    """Decode a time-varying categorical distribution file

    Arguments
        input_file
            Time-varying categorical distribution file
            shape=(frames, states)
        transition_file
            Categorical transition matrix file; defaults to uniform
            shape=(states, states)
        initial_file
            Categorical initial distribution file; defaults to uniform
            shape=(states,)
        log_probs
            Whether inputs are in (natural) log space
        gpu
            GPU index to use for decoding
        num_threads
            Ignored

    Returns
        indices
            The decoded bin indices
            shape=(1, frames)   (as the reference actually returns, core.py:244,261)
    """
    # Load observations (core.py:244)
    observation = torch.load(input_file, map_location='cpu').unsqueeze(dim=0)

    # Load transition probabilities; logged WITHOUT tiny here (core.py:247-250)
    if transition_file:
        transition = torch.load(transition_file, map_location='cpu')
        if log_probs:
            transition = torch.log(transition)
    else:
        transition = None

    # Load initial distribution, used raw (core.py:255-258)
    initial = torch.load(initial_file, map_location='cpu') if initial_file else None

    return from_probabilities(
        observation=observation,
        transition=transition,
        initial=initial,
        log_probs=log_probs,
        gpu=gpu,
        num_threads=num_threads)


def from_file_to_file(
    input_file: Union[str, os.PathLike],
    output_file: Union[str, os.PathLike],
    transition_file: Optional[Union[str, os.PathLike]] = None,
    initial_file: Optional[Union[str, os.PathLike]] = None,
    log_probs: bool = False,
    gpu: Gpu = None,
    num_threads: Optional[int] = None
) -> None:
    """Decode a time-varying categorical distribution file and save
    (core.py:270-307; the saved tensor has shape (1, frames) like the reference's)"""
    indices = from_file(
        input_file, transition_file, initial_file, log_probs, gpu=gpu, num_threads=num_threads)
    torch.save(indices, output_file)


def from_files_to_files(
    input_files: List[Union[str, os.PathLike]],
    output_files: List[Union[str, os.PathLike]],
    transition_file: Optional[Union[str, os.PathLike]] = None,
    initial_file: Optional[Union[str, os.PathLike]] = None,
    log_probs: bool = False,
    gpu: Gpu = None,
    num_threads: Optional[int] = None
) -> None:
    """Decode time-varying categorical distribution files and save

    Arguments as /root/reference/torbi/core.py:310-339.  With several GPUs the
    files are sharded across them by size (a proxy for frame count), longest
    first; each GPU runs its own loader / decode / save loop.
    """
    # Load transition probabilities; logged WITH tiny here (core.py:341-347)
    if transition_file:
        transition = torch.load(transition_file, map_location='cpu')
        if log_probs:
            transition = torch.log(transition + torch.finfo(transition.dtype).tiny)
    else:
        transition = None

    # Load initial distribution (core.py:350-353)
    initial = torch.load(initial_file, map_location='cpu') if initial_file else None

    # Preserve file mapping (core.py:356-358)
    mapping = {
        input_file: output_file
        for input_file, output_file in zip(input_files, output_files)}

    devices = _resolve_devices(gpu)

    # Results do not depend on batch composition (every item decodes as if alone),
    # so files are batched by size -- a proxy for frame count -- which keeps the
    # zero padding of collate() (collate.py:26-31) and the spread of lengths inside a
    # batch small.  The reference batches in the given order (loader.py:19-25).
    sizes = [os.path.getsize(file) for file in input_files]

    # In-process loading: the native pipeline (C++ reader / writer threads, pinned ring of packed
    # batch buffers, entropy chunking on the device when MIN_CHUNK_SIZE is set).  Worker-process
    # loaders keep the DataLoader route of the reference.
    native = torbi_b200.NUM_WORKERS == 0 and os.environ.get('TORBI_B200_NATIVE_IO', '1') != '0'

    def run(files, file_sizes, device, progress, io_threads):
        if native:
            _from_files_native(files, file_sizes, mapping, transition, initial, log_probs, device,
                               progress, io_threads)
        else:
            from_dataloader(
                dataloader=torbi_b200.data.loader(
                    files, batch_sampler=size_bucketed_batches(file_sizes, torbi_b200.BATCH_SIZE),
                    collate_fn=_pinned_collate()),
                output_files=mapping,
                transition=transition,
                initial=initial,
                log_probs=log_probs,
                gpu=device,
                num_threads=num_threads,
                progress=progress,
                prefetch=2,
                save_workers=4)

    cores = os.cpu_count() or 1
    if len(devices) == 1:
        run(input_files, sizes, devices[0], True, cores)
        return

    shards = lpt_shards(sizes, len(devices))
    errors: List[BaseException] = []

    def work(k, items):
        try:
            run([input_files[i] for i in items], [sizes[i] for i in items], devices[k], k == 0,
                max(2, cores // len(devices)))
        except BaseException as error:  # noqa: BLE001
            errors.append(error)

    threads = [threading.Thread(target=work, args=(k, items)) for k, items in enumerate(shards)]
    for thread in threads:
        thread.start()
    for thread in threads:
        thread.join()
    if errors:
        raise errors[0]


def _from_files_native(input_files, sizes, mapping, transition, initial, log_probs, device, progress,
                       io_threads, batch_bytes=1 << 29, depth=2):
    """The file pipeline of one GPU.  A loader thread fills a ring of `depth + 1` pinned batch
    buffers through the native reader (csrc/ptfile.cpp: central directory + pickle parsed in
    C++, payloads pread() straight into the batch by `io_threads` host threads) while the GPU
    decodes the previous batch; results are written by the native writer on a third thread.
    Batches are PACKED: the files lie back to back in one (rows, states) buffer, nothing is padded
    to the longest file (collate.py:26-31 pads), neither on the bus nor in the score history; with
    MIN_CHUNK_SIZE the files are cut at their low-entropy frames on the device (chunk.py:12-85) and
    the pieces decoded in place.  Per file the result is the reference's (torch.load -> [chunk] ->
    pad -> decode -> [join] -> torch.save of an int32 (frames,) tensor, core.py:376-463), only the
    plumbing differs.  TORBI_B200_PACKED=0 keeps the padded batches of the reference."""
    from .data import native as native_io

    packed = os.environ.get('TORBI_B200_PACKED', '1') != '0'
    chunked = torbi_b200.MIN_CHUNK_SIZE is not None
    if packed:
        batches = size_capped_batches(sizes, torbi_b200.BATCH_SIZE, max_bytes=batch_bytes)
    else:
        batches = size_bucketed_batches(sizes, torbi_b200.BATCH_SIZE, max_padded_bytes=batch_bytes)
    if not batches:
        return
    # file size bounds the payload, the first file of a batch is its largest
    if packed:
        numel = max(sum(int(sizes[i]) for i in batch) for batch in batches) // 4 + 1
    else:
        numel = max(len(batch) * int(sizes[batch[0]]) for batch in batches) // 4 + 1
    if packed and sum(int(size) for size in sizes) >= batch_bytes:
        numel = max(numel, batch_bytes // 4 + 1)      # one ring-buffer size for every large call: the kept buffers fit
    pin = torch.cuda.is_available()
    free: queue.Queue = queue.Queue()
    timing = os.environ.get('TORBI_B200_PIPELINE_TIMING') == '1'
    import time as _time
    t_start = _time.perf_counter()
    ring = _take_pinned_buffers(depth + 1, numel, pin)
    if timing:
        print(f'pipeline: {len(input_files)} files, {len(batches)} batches, ring of {depth + 1} x {numel * 4 >> 20} MB '
              f'ready after {_time.perf_counter() - t_start:.3f} s', flush=True)
    for buffer in ring:
        free.put(buffer)
    ready: queue.Queue = queue.Queue(maxsize=depth)
    done = object()
    stop = threading.Event()          # set when the consumer gives up: the producer must not block for ever

    def handoff(q, item=None, get=False):
        """put / get that gives up once `stop` is set"""
        while not stop.is_set():
            try:
                return (q.get(timeout=0.2),) if get else q.put(item, timeout=0.2) or (None,)
            except (queue.Empty, queue.Full):
                continue
        return None

    def produce():
        try:
            for batch in batches:
                files = [input_files[i] for i in batch]
                got = handoff(free, get=True)
                if got is None:
                    return
                buffer = got[0]
                loaded = None
                if packed:
                    loaded = native_io.load_batch_packed(files, threads=io_threads, out=buffer)
                elif not chunked:
                    loaded = native_io.load_batch(files, threads=io_threads, out=buffer)
                    if loaded is not None:
                        loaded = (loaded[0], None, loaded[1], None)
                if loaded is None:
                    # mixed dtypes / widths / legacy files: the reference's own Dataset (which chunks on
                    # the CPU when MIN_CHUNK_SIZE is set) and collate (file dtype kept, cast later)
                    dataset = torbi_b200.data.Dataset(files)
                    observation, batch_frames, batch_chunks, _ = torbi_b200.data.collate(
                        [dataset[i] for i in range(len(files))], pin_memory=pin)
                    loaded = (observation, None, batch_frames, batch_chunks)
                elif len(loaded) == 3:
                    loaded = (loaded[0], loaded[1], loaded[2], None)
                if handoff(ready, (loaded, files, buffer)) is None:
                    return
            handoff(ready, done)
        except BaseException as error:  # noqa: BLE001 - re-raised by the consumer
            handoff(ready, error)

    producer = threading.Thread(target=produce, daemon=True)
    producer.start()

    if transition is not None:
        log_transition = transition if log_probs else torch.log(transition)
        banded = _is_banded(log_transition.to(torch.float32))
    else:
        banded = None
    bar = _progress(len(input_files)) if progress else None
    saver = concurrent.futures.ThreadPoolExecutor(1)
    pending = []
    try:
        while True:
            item = ready.get()
            if item is done:
                break
            if isinstance(item, BaseException):
                raise item
            (observation, item_row, batch_frames, batch_chunks), files, buffer = item
            if timing:
                print(f'pipeline: batch of {len(files)} files ({observation.numel() * observation.element_size() >> 20} MB) '
                      f'loaded at {_time.perf_counter() - t_start:.3f} s', flush=True)
            use_banded = (banded and len(files) >= 32) if banded is not None else None
            if item_row is not None:
                indices = _from_probabilities_packed(
                    observation, item_row, batch_frames, transition, initial, log_probs,
                    torch.device(f'cuda:{device}'), banded=use_banded,
                    min_chunk_size=torbi_b200.MIN_CHUNK_SIZE, stats=_PIPELINE_STATS)
                indices = indices.cpu()        # also orders the host->device copy before the buffer is reused
                free.put(buffer)
                rows = [indices[int(row):int(row) + int(frames)] for row, frames in zip(item_row, batch_frames)]
            else:
                indices = from_probabilities(
                    observation=observation,
                    batch_frames=batch_frames,
                    transition=transition,
                    initial=initial,
                    log_probs=log_probs,
                    gpu=device,
                    banded=use_banded)
                indices = indices.cpu()
                free.put(buffer)
                if batch_chunks is not None and chunked:
                    rows = [r.to(torch.int32) for r in torbi_b200.data.separate(                # core.py:438-448
                        indices=indices, batch_chunks=batch_chunks, batch_frames=batch_frames)]
                else:
                    rows = [indices[i, :int(frames)] for i, frames in enumerate(batch_frames)]  # core.py:452-457
            pending.append(saver.submit(native_io.save_int32, rows, [mapping[file] for file in files],
                                        io_threads))
            if bar is not None:
                bar.update(len(files))
        if timing:
            print(f'pipeline: last batch decoded at {_time.perf_counter() - t_start:.3f} s', flush=True)
        for future in pending:
            future.result()
        if timing:
            print(f'pipeline: last file written at {_time.perf_counter() - t_start:.3f} s', flush=True)
    finally:
        # a failed decode / save must not leave the loader thread blocked on a full queue holding
        # pinned batch buffers: tell it to stop, then wait for it
        stop.set()
        producer.join()
        saver.shutdown(wait=True)
        _return_pinned_buffers(ring)
        if bar is not None:
            bar.close()


# Page-locking 3 x 512 MB takes ~0.4 s -- as long as decoding a thousand files -- so the ring buffers of the
# file pipeline are kept between calls (release_pinned_buffers() frees them).
_PINNED_POOL: List[torch.Tensor] = []
_PINNED_LOCK = threading.Lock()


def _take_pinned_buffers(count, numel, pin):
    taken = []
    with _PINNED_LOCK:
        for buffer in list(_PINNED_POOL):
            if len(taken) < count and buffer.numel() >= numel and buffer.is_pinned() == pin:
                _PINNED_POOL.remove(buffer)
                taken.append(buffer)
    while len(taken) < count:
        taken.append(torch.empty(numel, dtype=torch.float32, pin_memory=pin))
    return taken


def _return_pinned_buffers(buffers, keep=6):
    with _PINNED_LOCK:
        _PINNED_POOL.extend(buffers)
        _PINNED_POOL.sort(key=lambda buffer: -buffer.numel())
        del _PINNED_POOL[keep:]


def release_pinned_buffers():
    """Free the page-locked ring buffers the file pipeline keeps between calls."""
    with _PINNED_LOCK:
        _PINNED_POOL.clear()


# counters of the last file pipelines (frames seen / re-checked by the device chunker, pieces decoded)
_PIPELINE_STATS: Dict[str, int] = {}


###############################################################################
# Utilities
###############################################################################


def _pinned_collate():
    """collate() that assembles the batch directly in page-locked memory when a GPU is there
    (in-process loaders only: worker processes hand tensors over through shared memory)."""
    pin = torch.cuda.is_available() and torbi_b200.NUM_WORKERS == 0
    return functools.partial(torbi_b200.data.collate, pin_memory=pin)


def _prefetched(dataloader, depth):
    """Iterate `dataloader` on a background thread, `depth` batches ahead, with the
    padded observation pinned so its host->device copy runs at full PCIe rate."""
    if depth <= 0:
        yield from dataloader
        return
    batches: queue.Queue = queue.Queue(maxsize=depth)
    done = object()
    stop = threading.Event()          # set when the consumer stops iterating (error, early exit)

    def put(item):
        while not stop.is_set():
            try:
                batches.put(item, timeout=0.2)
                return True
            except queue.Full:
                continue
        return False

    def produce():
        try:
            for observation, batch_frames, batch_chunks, files in dataloader:
                if torch.cuda.is_available() and not observation.is_pinned():
                    observation = observation.pin_memory()
                if not put((observation, batch_frames, batch_chunks, files)):
                    return
            put(done)
        except BaseException as error:  # noqa: BLE001 - re-raised by the consumer
            put(error)

    producer = threading.Thread(target=produce, daemon=True)
    producer.start()
    try:
        while True:
            item = batches.get()
            if item is done:
                return
            if isinstance(item, BaseException):
                raise item
            yield item
    finally:
        stop.set()
        producer.join()


def from_dataloader(
    dataloader: torch.utils.data.DataLoader,
    output_files: Dict[
        Union[str, bytes, os.PathLike],
        Union[str, bytes, os.PathLike]],
    transition: Optional[torch.Tensor] = None,
    initial: Optional[torch.Tensor] = None,
    log_probs: bool = False,
    gpu: Gpu = None,
    num_threads: Optional[int] = 1,
    progress: bool = True,
    prefetch: int = 0,
    save_workers: int = 0
) -> None:
    """Decode time-varying categorical distributions from dataloader
    (core.py:376-463): decode each collated batch, then save one int32
    `(frames,)` tensor per input file.

    `prefetch` > 0 loads / collates / pins that many batches ahead on a background
    thread while the GPU decodes; `save_workers` > 0 writes the outputs from a
    thread pool.  Both default to the reference's serial behaviour."""
    bar = _progress(len(dataloader.dataset)) if progress else None
    savers = concurrent.futures.ThreadPoolExecutor(save_workers) if save_workers > 0 else None
    pending = []

    def submit(function, *args):
        if savers is None:
            function(*args)
        else:
            pending.append(savers.submit(function, *args))

    try:
        for observation, batch_frames, batch_chunks, input_filenames in _prefetched(dataloader, prefetch):

            indices = from_probabilities(
                observation=observation,
                batch_frames=batch_frames,
                transition=transition,
                initial=initial,
                log_probs=log_probs,
                gpu=gpu,
                num_threads=num_threads)

            filenames = [output_files[file] for file in input_filenames]
            indices = indices.cpu().detach()

            if torbi_b200.MIN_CHUNK_SIZE is not None:
                # Combine chunks (core.py:438-448)
                separated = torbi_b200.data.separate(
                    indices=indices, batch_chunks=batch_chunks, batch_frames=batch_frames)
                for item, filename in zip(separated, filenames):
                    submit(save, item, filename)
            else:
                # core.py:452-457
                for item, filename, frames in zip(indices, filenames, batch_frames.cpu()):
                    submit(save_masked, item, filename, frames)

            if bar is not None:
                bar.update(len(input_filenames))
        for future in pending:
            future.result()
    finally:
        if savers is not None:
            savers.shutdown(wait=True)
        if bar is not None:
            bar.close()


def _progress(total):
    try:
        import tqdm
        return tqdm.tqdm(total=total, desc=torbi_b200.CONFIG, dynamic_ncols=True)
    except ImportError:
        return None


def save(tensor, file):
    """Save tensor (core.py:466-468)"""
    torch.save(tensor.clone(), file)


def save_masked(tensor, file, length):
    """Save masked tensor (core.py:471-473)"""
    torch.save(tensor[..., :length].clone(), file)
