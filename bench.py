#!/usr/bin/env python
"""Benchmark of the one hot path: timesteps decoded per second, 1440 states,
batch 512, 2000 frames, triangular transition (BASELINE.json config 2).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl graft|reference]

A "step" is one pass of the hot path over one batch of synthetic input.  For
N > 1 the driver launches one process per GPU (torchrun); every rank decodes its
own batch of 512 sequences (weak scaling: the path shards over independent
sequences, no data-path collective) and rank 0 prints ONE JSON line.  The same
invocation also measures STRONG scaling -- one global batch of 512 sequences,
512 / N per GPU -- and reports it under "strong"; at N = 1 the line carries
BASELINE configs 1, 3 and 5 under "configs" (device-timed, with clocks).
Every rank binds its host threads to the NUMA node of its GPU before it
allocates pinned memory ("host_binding" in config).

  value      whole-job timesteps/s with inputs resident in HBM, CUDA events, max over ranks
  e2e        the same through the C ABI host entry (torbi_b200_decode_host): pinned host
             buffers, H2D of the observations and D2H of the indices inside the timed region
  roofline   the forward kernel against the FP32 max-plus issue roofline and,
             secondarily, its algorithmic HBM bytes against the measured copy bandwidth
  cpu_baseline  the reference's own CPU operator (oracle/_ref) on this box's host cores,
             on a bounded sample of the same workload (rank 0, N = 1 only)

`--impl reference` times the reference's CPU implementation of the path instead.
"""
import argparse
import json
import os
import pathlib
import statistics
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))

METRIC = 'timesteps decoded/sec, 1440 states batch 512'
UNIT = 'timesteps/s'


def parse_args():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gpus', type=int, default=1)
    parser.add_argument('--steps', type=int, default=5)
    parser.add_argument('--warmup', type=int, default=3)
    parser.add_argument('--impl', default='graft', choices=['graft', 'reference'])
    parser.add_argument('--batch', type=int, default=512)
    parser.add_argument('--frames', type=int, default=2000)
    parser.add_argument('--states', type=int, default=1440)
    parser.add_argument('--scaling', default='weak', choices=['weak', 'strong'],
                        help='strong: --batch is the GLOBAL batch, split evenly over the ranks')
    parser.add_argument('--no-e2e', action='store_true')
    parser.add_argument('--no-e2e-half', action='store_true', help='skip the fp16-input end-to-end leg')
    parser.add_argument('--cpu-seconds', type=float, default=None,
                        help='CPU work per reference step / cpu_baseline sample (default: sized automatically)')
    parser.add_argument('--no-cpu-baseline', action='store_true')
    parser.add_argument('--kernel', default='auto', choices=['auto', 'generic', 'tiled', 'wide', 'split', 'cluster', 'flow'])
    parser.add_argument('--no-strong', action='store_true', help='skip the strong-scaling leg (N > 1)')
    parser.add_argument('--no-configs', action='store_true', help='skip BASELINE configs 1 / 3 / 5 (N = 1)')
    parser.add_argument('--skip-config3', action='store_true')
    parser.add_argument('--skip-config5', action='store_true')
    parser.add_argument('--config5-frames', type=int, default=1_000_000)
    parser.add_argument('--skip-config4', action='store_true')
    parser.add_argument('--config4-files', type=int, default=2000,
                        help='files of the synthetic corpus (BASELINE config 4 is ~40000 files / ~20 M timesteps; '
                             'the default is 1/20 of it, same length distribution, on /dev/shm)')
    parser.add_argument('--tile-rows', type=int, default=0, choices=[0, 16, 32, 64])
    parser.add_argument('--split-streamed', action='store_true', help='split kernel: streamed 16 x 40 tiles')
    parser.add_argument('--backpointers', action='store_true',
                        help="tiled kernel in backpointer mode (uint16 trellis) instead of score history + recomputed backtrace")
    parser.add_argument('--banded', action='store_true', help='exact banded-transition fast path (not the contract metric)')
    return parser.parse_args()


def measured_peaks():
    path = ROOT / 'MEASURED_PEAKS.json'
    if path.exists():
        peaks = json.loads(path.read_text())
        return dict(hbm_gbs=float(peaks['hbm_gbs']), sm_max_mhz=float(peaks['sm_max_mhz']),
                    source='measured (MEASURED_PEAKS.json)')
    return dict(hbm_gbs=6650.0, sm_max_mhz=1965.0, source='fallback (B200_PROFILING.md)')


def profiled_traffic(kernel, args, backtrace='backtrace_recompute'):
    """DRAM bytes (read + write) of one launch of the forward kernel from the
    committed `ncu --set full` capture of this same workload, or None."""
    path = ROOT / 'profiles' / ('r02_forward_tiled_hist_ncu_full.json' if backtrace == 'backtrace_recompute'
                                else 'forward_tiled_ncu_full.json')
    if not path.exists():
        path = ROOT / 'profiles' / 'forward_tiled_hist_ncu_full.json'
    if kernel != 'forward_tiled' or not path.exists() or \
            (args.batch, args.frames, args.states) != (512, 2000, 1440):
        return None
    m = json.loads(path.read_text())
    scale = {'Gbyte': 1e9, 'Mbyte': 1e6, 'Kbyte': 1e3, 'byte': 1.0}
    total = 0.0
    for key in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
        total += float(m[key]['value'].replace(',', '')) * scale[m[key]['unit']]
    return total


def workload_config(args, world):
    return {
        'workload': 'BASELINE config 2: from_probabilities hot path, log-softmax(randn) observations, '
                    'triangular pitch transition log(A+tiny), uniform initial, fused eps floor',
        'states': args.states, 'batch_per_gpu': args.batch, 'frames': args.frames,
        'global_batch': args.batch * world, 'sharding': f'independent sequences, {world} rank(s), no collective',
        'l2_policy': 'inputs (5.9 GB/step) larger than L2; no flush needed',
    }


# ---------------------------------------------------------------------------
# clocks during the timed region (B200_PROFILING.md recipe)
# ---------------------------------------------------------------------------

class ClockSampler:
    QUERY = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
             'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
             'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.QUERY}', '--format=csv,noheader,nounits',
                 '-i', str(self.index), '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm, sm_max, power, reasons = [], [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.lines:
            f = [x.strip() for x in line.split(',')]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); sm_max.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for name, value in zip(names, f[5:9]):
                if value.lower() == 'active':
                    reasons.add(name)
        if not sm:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': [], 'samples': 0}
        return {'sm_mhz': statistics.median(sm), 'sm_max_mhz': max(sm_max),
                'power_w_max': max(power), 'reasons': sorted(reasons), 'samples': len(sm)}


# ---------------------------------------------------------------------------
# CPU reference (oracle/_ref = the unmodified reference operator)
# ---------------------------------------------------------------------------

def cpu_reference_runner():
    """Returns (fn(obs, frames, A, init), kind, cores)."""
    from oracle import oracle
    cores = os.cpu_count() or 1
    if oracle.ref_available():
        def run(obs, frames, A, init):
            return oracle.ref_decode(obs, frames, A, init, num_threads=cores)
        return run, 'reference', cores

    def run(obs, frames, A, init):
        return oracle.c_decode(obs, frames, A, init, exact_scan=True, num_threads=cores)
    return run, 'port', cores


def cpu_sample(args, seconds, seed=1234):
    """A bounded sample of the workload sized for about `seconds` of CPU work:
    the first rows of the batch, truncated in time.  CPU throughput does not
    depend on batch size (the reference loops over items sequentially,
    viterbi.cpp:65)."""
    import cases
    run, kind, cores = cpu_reference_runner()
    S = args.states
    rng = np.random.default_rng(seed)
    A = cases.log_triangular_transition(S)
    init = np.full((S,), np.log(1.0 / S + np.finfo(np.float32).tiny), np.float32)
    probe = cases.log_softmax(rng.standard_normal((1, 33, S)))
    run(probe[:, :5], np.array([5], np.int32), A, init)                 # warm the thread pool
    rate = 0.0
    for _ in range(3):                                                  # (the first passes still ramp up)
        t0 = time.perf_counter()
        run(probe, np.array([33], np.int32), A, init)
        rate = max(rate, 32 / (time.perf_counter() - t0))               # recursion steps / s
    total = int(max(64, min(rate * seconds, 8 * args.frames)))
    B = max(1, min(8, total // args.frames)) if total >= args.frames else 1
    T = max(2, min(args.frames, total // B))
    obs = cases.log_softmax(rng.standard_normal((B, T, S)))
    frames = np.full((B,), T, np.int32)

    def step():
        t0 = time.perf_counter()
        run(obs, frames, A, init)
        return time.perf_counter() - t0
    return step, B * T, kind, cores, f'{B} sequences x {T} frames of the same synthetic workload (S={S})'


def bench_reference(args):
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if rank != 0:
        return
    budget = 150.0 / max(1, args.steps + args.warmup)
    step, timesteps, kind, cores, sample = cpu_sample(args, args.cpu_seconds or min(20.0, budget))
    for _ in range(args.warmup):
        step()
    elapsed = sum(step() for _ in range(args.steps))
    value = timesteps * args.steps / elapsed
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * elapsed / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32',
        'data': 'synthetic', 'config': workload_config(args, world),
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind, 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------

def bind_to_gpu_numa_node(local, world):
    """Pin this rank's host threads (and, through first touch, the pinned buffers it allocates from
    here on) to the CPUs of the NUMA node its GPU hangs off.  With N ranks on one box all inheriting
    the full CPU mask, the pinned staging buffers otherwise land on whatever node the allocating
    thread happened to run on and half of the host->device traffic crosses the socket interconnect.
    Falls back to an even split of the inherited mask when sysfs has no NUMA information."""
    info = {'method': 'none'}
    if os.environ.get('TORBI_B200_NO_BIND') == '1' or not hasattr(os, 'sched_setaffinity'):
        return info
    allowed = sorted(os.sched_getaffinity(0))
    try:
        props = torch.cuda.get_device_properties(local)
        bus = f'{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0'
        node = int(pathlib.Path(f'/sys/bus/pci/devices/{bus}/numa_node').read_text())
        if node >= 0:
            cpus = set()
            for part in pathlib.Path(f'/sys/devices/system/node/node{node}/cpulist').read_text().strip().split(','):
                lo, _, hi = part.partition('-')
                cpus.update(range(int(lo), int(hi or lo) + 1))
            cpus &= set(allowed)
            if cpus:
                os.sched_setaffinity(0, cpus)
                return {'method': 'numa', 'numa_node': node, 'cpus': len(cpus), 'pci': bus}
    except (OSError, ValueError, AttributeError):
        pass
    if world > 1 and len(allowed) >= world:
        share = len(allowed) // world
        os.sched_setaffinity(0, allowed[local * share:(local + 1) * share])
        info = {'method': 'even split of the inherited mask', 'cpus': share}
    return info


def synthetic_observations(B, T, S, device, seed, chunk_bytes=2 << 30):
    g = torch.Generator(device=device).manual_seed(seed)
    obs = torch.empty((B, T, S), dtype=torch.float32, device=device)
    rows = max(1, int(chunk_bytes // (T * S * 4)))
    for start in range(0, B, rows):
        n = min(rows, B - start)
        if rows == 1 and T * S * 4 > chunk_bytes:                 # one long sequence: chunk along T
            step = max(1, int(chunk_bytes // (S * 4)))
            for t0 in range(0, T, step):
                x = torch.randn((min(step, T - t0), S), device=device, generator=g)
                obs[start, t0:t0 + x.shape[0]] = torch.log_softmax(x, dim=-1)
                del x
        else:
            x = torch.randn((n, T, S), device=device, generator=g)
            obs[start:start + n] = torch.log_softmax(x, dim=-1)
            del x
    return obs


def make_workload(args, device, seed):
    import cases
    B, T, S = args.batch, args.frames, args.states
    obs = synthetic_observations(B, T, S, device, seed)
    A = torch.from_numpy(cases.log_triangular_transition(S)).to(device)
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)),
                      dtype=torch.float32, device=device)
    frames = torch.full((B,), T, dtype=torch.int32)
    return obs, frames, A, init


def timed_decode(lib, obs, frames, A, init, flags, steps, warmup, barrier, local):
    """`steps` decodes of one resident batch bracketed by CUDA events (barrier + synchronize on both
    sides), clocks sampled during the timed region, per-kernel device times from the library."""
    import ctypes
    import torbi_b200
    from torbi_b200 import _ffi
    fwd_ms, back_ms = [], []
    indices = None
    for _ in range(warmup):
        indices = torbi_b200.decode(obs, frames, A, init, flags=flags)
    barrier()
    launches0 = lib.torbi_b200_launch_count()
    torch.cuda.reset_peak_memory_stats()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clocks:
        barrier()
        start.record()
        for _ in range(steps):
            indices = torbi_b200.decode(obs, frames, A, init, flags=flags)
            f, b = ctypes.c_float(), ctypes.c_float()
            _ffi.check(lib.torbi_b200_last_timing(ctypes.byref(f), ctypes.byref(b)))
            fwd_ms.append(f.value)
            back_ms.append(b.value)
        end.record()
        barrier()
    return {
        'elapsed_ms': start.elapsed_time(end), 'fwd_ms': statistics.mean(fwd_ms), 'back_ms': statistics.mean(back_ms),
        'kernel': lib.torbi_b200_last_kernel().decode(), 'backtrace': lib.torbi_b200_last_backtrace().decode(),
        'launches': int(lib.torbi_b200_launch_count() - launches0), 'clocks': clocks.summary(),
        'checksum': int(indices.sum(dtype=torch.int64)), 'indices': indices,
        'peak_hbm_gb': torch.cuda.max_memory_allocated() / 1e9,
    }


def fp32_peak(sm_count, peaks):
    return sm_count * 128 * peaks['sm_max_mhz'] * 1e6 / 1e12       # T lane-op/s


def other_configs(lib, device, local, peaks, sm_count, barrier, args):
    """BASELINE configs 1, 3 and 5 on this GPU (N = 1 only), device-timed like the headline: inputs
    resident in HBM, CUDA events, clocks sampled during the timed region.  Each entry names its kernel,
    its fraction of the FP32 max-plus issue roofline (2 S^2 lane-ops per timestep) and the peak HBM in use."""
    import cases
    from torbi_b200 import _ffi
    out = {}
    flags = _ffi.EPS_FLOOR | _ffi.TIME_KERNELS

    def entry(name, obs, frames, A, init, steps, warmup, note):
        S = obs.shape[2]
        r = timed_decode(lib, obs, frames, A, init, flags, steps, warmup, barrier, local)
        timesteps = int(frames.sum())
        value = timesteps * steps / (r['elapsed_ms'] * 1e-3)
        out[name] = {
            'workload': note, 'value': value, 'unit': UNIT, 'ms_per_step': r['elapsed_ms'] / steps, 'steps': steps,
            'warmup': warmup, 'kernel': r['kernel'], 'backtrace': r['backtrace'], 'forward_ms': r['fwd_ms'],
            'backtrace_ms': r['back_ms'],
            'frac_fp32_maxplus': 2.0 * S * S * timesteps / (r['fwd_ms'] * 1e-3) / 1e12 / fp32_peak(sm_count, peaks),
            'peak_hbm_gb': r['peak_hbm_gb'], 'clocks': r['clocks'], 'checksum': r['checksum']}
        if r['kernel'] == 'forward_wide' and obs.shape[0] == 1:
            # one sequence through the wide kernel: a step = the grid barrier (1.25 us measured for 148 CTAs,
            # profiles/r02_microbench_sync.txt) + the longer of the scan and the L2 stream of the transition rows
            # shared memory cannot hold (13.7 TB/s measured L2->SM stream rate, same file)
            clk = peaks['sm_max_mhz'] * 1e6
            ldp = -(-S // 96) * 96
            rows = -(-S // sm_count)
            resident = min(rows, (226 * 1024 - 4 * ldp) // (4 * ldp))
            streamed_bytes = (rows - resident) * ldp * 4.0 * -(-S // rows)
            floor_us = 1.25 + 1e6 * max(S * S / sm_count / 64.0 / clk, streamed_bytes / 13.7e12)
            us = 1e3 * r['fwd_ms'] / max(1, timesteps - 1)
            out[name]['latency_bound'] = {
                'us_per_step': us, 'floor_us_per_step': floor_us, 'frac': floor_us / us,
                'streamed_bytes_per_step': streamed_bytes, 'l2_stream_gbs': streamed_bytes / (us * 1e-6) / 1e9,
                'note': 'single-sequence recursion is latency / L2-stream bound, not FP32 bound: floor = grid barrier '
                        '(1.25 us measured) + max(scan at 64 cells/clk/SM, L2 stream of the non-resident transition rows at '
                        'the measured 13.7 TB/s)'}
        if r['kernel'] == 'forward_flow':
            # one sequence: a timestep cannot start before the previous one's scores have crossed L2.  Floor of
            # a step = L2 store + L2 load latency (2 x ~250 cycles, B300_MICROARCH.md "L2 hit latency") + the
            # longer of the CTA's scan (S^2 / #SM cells at 64 cells/clk/SM) and the streaming of the transition
            # rows that neither registers nor shared memory hold (measured L2->SM stream rate,
            # profiles/r02_microbench_sync.txt: 13.7 TB/s at 64 KB per CTA)
            clk = peaks['sm_max_mhz'] * 1e6
            rows = -(-S // sm_count)
            resident = (20 if S <= 2048 else 10 if S <= 4096 else 4) + (220 * 1024) // (4 * (-(-S // 96) * 96))
            streamed_bytes = max(0, rows - resident) * S * 4.0 * sm_count
            floor_us = 1e6 * (500.0 / clk + max(S * S / sm_count / 64.0 / clk, streamed_bytes / 13.7e12))
            us = 1e3 * r['fwd_ms'] / max(1, timesteps - 1)
            out[name]['latency_bound'] = {
                'us_per_step': us, 'floor_us_per_step': floor_us, 'frac': floor_us / us,
                'streamed_bytes_per_step': streamed_bytes,
                'note': 'single-sequence recursion is latency bound: floor = L2 round trip (2 x 250 cycles) + max(scan of '
                        'S^2/#SM cells at 64 cells/clk/SM, L2 stream of the non-resident transition rows at 13.7 TB/s)'}

    # config 1: 1440 states, batch 1, 2000 frames, default uniform transition / initial
    S, T = 1440, 2000
    obs = synthetic_observations(1, T, S, device, 1234)
    A = torch.full((S, S), float(np.log(1.0 / S)), device=device)
    init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device=device)
    entry('c1', obs, torch.tensor([T], dtype=torch.int32), A, init, 10, 3,
          'config 1 on the GPU: S=1440, B=1, T=2000, uniform transition / initial (barrier-latency bound, not FP32 bound)')
    del obs

    # config 3: 1440 states, batch 256, 200..20000 frames (log-uniform), triangular transition
    if not args.skip_config3:
        rng = np.random.default_rng(1234)
        lengths = np.exp(rng.uniform(np.log(200), np.log(20000), size=256)).round().astype(np.int32)
        lengths[0], lengths[1] = 20000, 200
        obs = synthetic_observations(256, 20000, S, device, 4321)
        A = torch.from_numpy(cases.log_triangular_transition(S)).to(device)
        entry('c3', obs, torch.from_numpy(lengths), A, init, 3, 3,
              'config 3: S=1440, B=256, 200..20000 frames (log-uniform, seed 1234), triangular transition, dense path; '
              'bound by the 20000 sequential steps of its longest item')
        del obs
    torch.cuda.empty_cache()

    # config 5: 4096 states, batch 1, 10^6 frames, dense transition
    if not args.skip_config5:
        S, T = 4096, args.config5_frames
        obs = synthetic_observations(1, T, S, device, 55)
        g = torch.Generator(device=device).manual_seed(56)
        A = torch.log_softmax(torch.randn(S, S, device=device, generator=g), dim=1)
        init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device=device)
        # warm-up on three short prefixes (same kernels, same clocks) instead of three 10^6-frame passes
        import torbi_b200
        for _ in range(3):
            torbi_b200.decode(obs[:, :20000].contiguous(), torch.tensor([20000], dtype=torch.int32), A, init, flags=flags)
        entry('c5', obs, torch.tensor([T], dtype=torch.int32), A, init, 1, 0,
              f'config 5: S=4096, B=1, T={T}, dense transition (L2-stream / barrier bound, not FP32 bound); '
              'warm-up = 3 decodes of a 20000-frame prefix')
        del obs
    torch.cuda.empty_cache()
    return out


def config4_corpus(files, world):
    """Frame counts of the synthetic config-4 corpus: log-normal, median 430, sigma 0.55,
    clipped to [50, 5000], rescaled to ~500 frames per file on average (20 M timesteps / 40 k files)."""
    rng = np.random.default_rng(1234)
    frames = np.clip(np.exp(rng.normal(np.log(430), 0.55, size=files)), 50, 5000)
    frames = np.maximum(50, (frames * (500.0 * files / frames.sum())).round()).astype(np.int64)
    import torbi_b200
    sizes = (frames * 1440 * 4).tolist()
    return frames, torbi_b200.lpt_shards(sizes, world)          # the sharding from_files_to_files applies


def config4_leg(args, rank, world, local, device, barrier, max_over_ranks):
    """BASELINE config 4: from_files_to_files over a synthetic 1440-state corpus, the files sharded
    across the N GPUs by size (longest-processing-time first, as core.from_files_to_files does across
    `gpu=[...]`); here one process per GPU, each decoding its shard from /dev/shm end to end
    (native reader -> pinned packed batches -> decode -> native writer).  Whole-job timesteps/s =
    all frames / max over ranks of the wall time."""
    import shutil
    import tempfile
    import cases
    import torbi_b200
    frames, shards = config4_corpus(args.config4_files, world)
    mine = shards[rank] if rank < len(shards) else []
    root = pathlib.Path(tempfile.mkdtemp(prefix=f'torbi_c4_r{rank}_', dir='/dev/shm' if os.path.isdir('/dev/shm') else None))
    try:
        g = torch.Generator(device=device).manual_seed(4000 + rank)
        inputs, outputs = [], []
        for i in mine:
            x = torch.log_softmax(torch.randn((int(frames[i]), 1440), device=device, generator=g), dim=-1).cpu()
            torch.save(x, root / f'{i}.pt')
            inputs.append(root / f'{i}.pt')
            outputs.append(root / f'{i}.out.pt')
        torch.save(torch.from_numpy(cases.triangular_transition(1440)), root / 'transition.pt')
        # warm-up on a slice (arena, pinned ring buffers, library load), then the timed pass over the whole shard
        warm = min(len(inputs), 400)         # enough files for full-size (512 MB) pinned ring buffers, which are kept
        torbi_b200.from_files_to_files(inputs[:warm], outputs[:warm], root / 'transition.pt', log_probs=True, gpu=local)
        barrier()
        with ClockSampler(local) as clocks:
            t0 = time.perf_counter()
            torbi_b200.from_files_to_files(inputs, outputs, root / 'transition.pt', log_probs=True, gpu=local)
            torch.cuda.synchronize()
            seconds = time.perf_counter() - t0
            barrier()
        check = torch.load(outputs[0]) if outputs else None
        assert check is None or (check.dtype == torch.int32 and check.shape == (int(frames[mine[0]]),))
    finally:
        shutil.rmtree(root, ignore_errors=True)
    seconds, = max_over_ranks(seconds)
    total = int(frames.sum())
    return {'workload': f'config 4 at {args.config4_files}/40000 scale: {args.config4_files} files, {total} timesteps, S=1440 fp32 '
                        f'log-probs on /dev/shm, sharded over {world} GPU(s) by size (LPT), end to end incl. file read / write',
            'value': total / seconds, 'unit': UNIT, 'seconds': seconds, 'files': int(args.config4_files),
            'input_gb': total * 1440 * 4 / 1e9, 'read_gbs': total * 1440 * 4 / 1e9 / seconds, 'clocks': clocks.summary()}


def bench_graft(args):
    import torch.distributed as dist
    import torbi_b200
    from torbi_b200 import _ffi

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; torbi_b200 has no CPU path '
                         '(use --impl reference for the CPU reference)')
    torch.cuda.set_device(local)
    device = torch.device(f'cuda:{local}')
    binding = bind_to_gpu_numa_node(local, world)
    if world > 1:
        # NCCL may print its version banner on stdout (NCCL_DEBUG=VERSION on some boxes); the contract is
        # ONE JSON line there, so stdout points at stderr while the communicator comes up
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group('nccl', device_id=device)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(*values):
        times = torch.tensor(values, dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(times, op=dist.ReduceOp.MAX)
        return [float(v) for v in times]

    lib = _ffi.library()
    global_batch_strong = args.batch
    if args.scaling == 'strong':
        if args.batch % world:
            raise SystemExit('--scaling strong needs --batch divisible by the number of ranks')
        args.batch //= world                       # per-rank share from here on
    B, T, S = args.batch, args.frames, args.states
    obs, frames, A, init = make_workload(args, device, 1234 + rank)
    flags = _ffi.EPS_FLOOR | _ffi.TIME_KERNELS
    flags |= {'auto': 0, 'generic': _ffi.FORCE_GENERIC, 'tiled': _ffi.FORCE_TILED, 'wide': _ffi.FORCE_WIDE,
              'split': _ffi.FORCE_SPLIT, 'cluster': _ffi.FORCE_CLUSTER, 'flow': _ffi.FORCE_FLOW}[args.kernel]
    flags |= {0: 0, 16: _ffi.TILE_ROWS_16, 32: _ffi.TILE_ROWS_32, 64: _ffi.TILE_ROWS_64}[args.tile_rows]
    if args.banded:
        flags |= _ffi.BANDED
    if args.backpointers:
        flags |= _ffi.BACKPOINTERS
    if args.split_streamed:
        flags |= _ffi.SPLIT_STREAMED
    timesteps_per_step = int(frames.sum())

    main = timed_decode(lib, obs, frames, A, init, flags, args.steps, args.warmup, barrier, local)
    elapsed_ms = main['elapsed_ms']
    checksum = main['checksum']
    kernel, backtrace = main['kernel'], main['backtrace']

    # end to end through the C ABI host entry: pinned host buffers
    e2e_ms = half_ms = None
    h2d = d2h = 0
    if not args.no_e2e:
        host_obs = torch.empty((B, T, S), dtype=torch.float32, pin_memory=True)
        host_obs.copy_(obs)
        host_A, host_init = A.cpu().pin_memory(), init.cpu().pin_memory()
        host_out = torch.empty((B, T), dtype=torch.int32, pin_memory=True)
        e2e_flags = flags & ~_ffi.TIME_KERNELS
        h2d = host_obs.numel() * 4 + host_A.numel() * 4 + host_init.numel() * 4 + B * 4
        d2h = host_out.numel() * 4
        for _ in range(max(1, min(args.warmup, 2))):
            torbi_b200.decode_host(host_obs, frames, host_A, host_init, gpu=local, flags=e2e_flags, out=host_out)
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            torbi_b200.decode_host(host_obs, frames, host_A, host_init, gpu=local, flags=e2e_flags, out=host_out)
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        assert int(host_out.sum(dtype=torch.int64)) == checksum, 'host entry disagrees with device entry'
        del host_obs
        # the same end to end with the observations stored as fp16 on the host (the reference casts whatever
        # dtype it is given to fp32 on the host, core.py:191; here the 2-byte elements cross PCIe and are
        # upcast in the kernel): half the host->device bytes.  Not the contract number (other input values).
        half_ms = None
        if not args.no_e2e_half:
            host_half = torch.empty((B, T, S), dtype=torch.float16, pin_memory=True)
            host_half.copy_(obs)
            for _ in range(2):
                torbi_b200.decode_host(host_half, frames, host_A, host_init, gpu=local, flags=e2e_flags, out=host_out)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                torbi_b200.decode_host(host_half, frames, host_A, host_init, gpu=local, flags=e2e_flags, out=host_out)
            barrier()
            half_ms = 1e3 * (time.perf_counter() - t0)
            del host_half
        lib.torbi_b200_release()

    # Strong scaling in the same invocation (the conservative reading of "batch 512 at 1/2/4/8 B200",
    # SURVEY 8d): ONE global batch of 512 sequences, 512 / N per GPU (rank r contributes the first 512 / N
    # sequences of the batch it already holds); at N = 1 this IS the headline run.
    strong = None
    if args.scaling == 'weak' and not args.no_strong and global_batch_strong % world == 0:
        share = global_batch_strong // world
        if world == 1:
            s_run, s_frames = main, frames
        else:
            s_obs = obs[:share].contiguous() if share < B else obs
            s_frames = frames[:share]
            strong_flags = flags & ~(_ffi.TILE_ROWS_64 | _ffi.TILE_ROWS_32 | _ffi.TILE_ROWS_16) if args.kernel == 'auto' else flags
            s_run = timed_decode(lib, s_obs, s_frames, A, init, strong_flags, args.steps, args.warmup, barrier, local)
            del s_obs
        s_elapsed, = max_over_ranks(s_run['elapsed_ms'])
        strong = s_run, s_frames, share, s_elapsed

    elapsed_ms, e2e_max, half_max = max_over_ranks(elapsed_ms, e2e_ms or 0.0, half_ms or 0.0)
    del obs
    torch.cuda.empty_cache()
    c4 = None
    if not args.skip_config4 and not args.no_configs and args.scaling == 'weak':
        c4 = config4_leg(args, rank, world, local, device, barrier, max_over_ranks)

    if rank == 0:
        peaks = measured_peaks()
        total = timesteps_per_step * args.steps * world
        value = total / (elapsed_ms * 1e-3)
        sm_count = torch.cuda.get_device_properties(device).multi_processor_count
        # FP32 max-plus issue roofline: 2*S^2 lane-ops per timestep,
        # peak = SMs x 128 lanes x max SM clock
        kernel_ms = main['fwd_ms']
        ops = 2.0 * S * S * timesteps_per_step
        achieved = ops / (kernel_ms * 1e-3) / 1e12
        peak = fp32_peak(sm_count, peaks)
        # forward kernel: observation read + fp32 score history (or uint16 backpointers) written
        hbm_bytes = (4.0 * S + (4.0 if backtrace == 'backtrace_recompute' else 2.0) * S) * timesteps_per_step
        hbm_achieved = hbm_bytes / (kernel_ms * 1e-3) / 1e9
        config = workload_config(args, world)
        config['host_binding'] = binding
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': elapsed_ms / args.steps, 'higher_is_better': True,
            'scaling': args.scaling, 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
            'config': config,
            'e2e': None if e2e_ms is None else {
                'value': total / (e2e_max * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d,
                'd2h_bytes_per_step': d2h, 'ms_per_step': e2e_max / args.steps,
                'h2d_gbs_per_gpu': h2d * args.steps / (e2e_max * 1e-3) / 1e9,
                'api': 'torbi_b200_decode_host (C ABI, pinned host buffers)',
                'fp16_inputs': None if not half_ms else {
                    'value': total / (half_max * 1e-3), 'unit': UNIT, 'h2d_bytes_per_step': h2d - B * T * S * 2,
                    'note': 'same call with the observations stored as fp16 on the host (TORBI_B200_OBS_F16): '
                            'half the host->device bytes; other input values than the contract number'}},
            'gpu_launches': main['launches'],
            'clocks': main['clocks'],
            'roofline': {
                'bound': 'fp32_alu_issue', 'kernel': kernel, 'achieved': achieved, 'peak': peak,
                'unit': 'T lane-op/s', 'frac': achieved / peak, 'traffic': profiled_traffic(kernel, args, backtrace),
                'traffic_note': 'DRAM read+write bytes per launch, ncu --set full capture in profiles/ (not live)',
                'kernel_ms': kernel_ms, 'kernel_share_of_step': kernel_ms * args.steps / main['elapsed_ms'],
                'backtrace_ms': main['back_ms'], 'backtrace': backtrace,
                'ops_per_launch': ops, 'peak_source': f"{sm_count} SMs x 128 lanes x {peaks['sm_max_mhz']} MHz, {peaks['source']}",
                'timesteps_per_s_at_peak': peak * 1e12 / (2.0 * S * S),
                'hbm': {'bound': 'hbm', 'achieved': hbm_achieved, 'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                        'frac': hbm_achieved / peaks['hbm_gbs'], 'bytes_per_launch': hbm_bytes},
            },
            'peak_hbm_gb': main['peak_hbm_gb'],
            'checksum': checksum,
        }
        if strong is not None:
            s_run, s_frames, share, s_elapsed = strong
            s_total = int(s_frames.sum()) * args.steps * world
            line['strong'] = {
                'value': s_total / (s_elapsed * 1e-3), 'unit': UNIT, 'global_batch': global_batch_strong,
                'per_gpu_batch': share, 'ms_per_step': s_elapsed / args.steps, 'kernel': s_run['kernel'],
                'frac': 2.0 * S * S * int(s_frames.sum()) / (s_run['fwd_ms'] * 1e-3) / 1e12 / peak,
                'kernel_ms': s_run['fwd_ms'], 'backtrace_ms': s_run['back_ms'], 'clocks': s_run['clocks'],
                'note': 'strong scaling: the SAME global batch of 512 sequences split evenly over the GPUs, '
                        'device-resident, CUDA events, max over ranks; value = whole-job timesteps/s'}
        if world == 1 and not args.no_configs:
            line['configs'] = other_configs(lib, device, local, peaks, sm_count, barrier, args)
        if c4 is not None:
            line.setdefault('configs', {})['c4'] = c4
        if world == 1 and not args.no_cpu_baseline:
            step_cpu, timesteps, kind, cores, sample = cpu_sample(args, args.cpu_seconds or 15.0)
            seconds = step_cpu()
            line['cpu_baseline'] = {'value': timesteps / seconds, 'unit': UNIT, 'cores': cores,
                                    'kind': kind, 'sample': sample,
                                    'scope': "the reference's CPU operator only (torbi.decode); its from_probabilities "
                                             'adds three more CPU passes over the observations (core.py:189-197)'}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == 'reference':
        bench_reference(args)
    else:
        bench_graft(args)


if __name__ == '__main__':
    main()
