#!/usr/bin/env python
"""Time the reference's OWN CUDA operator on this GPU, synchronised.

oracle/_ref/cuda/torbi_ref_cuda.so is /root/reference/torbi/csrc/{ops.cpp,cuda/viterbi.cu}
compiled unmodified for sm_100a (oracle/build_ref.py cuda).  The reference times its op without
a device synchronisation (core.py:200-206), so its published GPU numbers are launch latency;
here every repetition is bracketed by CUDA events and a synchronize.  Its ties differ from its
CPU path (viterbi.cu:98-112), so this is a speed bar only, not a parity oracle.

    python tools/time_reference_cuda.py [--batch 512 --frames 2000 --states 1440 --reps 3]
Prints one JSON line per configuration.
"""
import argparse
import json
import pathlib
import sys

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--configs', default='512x2000x1440,64x2000x1440,1x2000x1440')
    parser.add_argument('--reps', type=int, default=3)
    args = parser.parse_args()
    import cases
    from oracle import build_ref
    if not build_ref.cuda_available():
        print(json.dumps({'impl': 'reference_cuda', 'unavailable': f'{build_ref.REF_CUDA_SO} not built'}))
        return
    torch.ops.load_library(str(build_ref.REF_CUDA_SO))
    op = torch.ops.torbi.viterbi_decode
    device = torch.device('cuda:0')
    for config in args.configs.split(','):
        B, T, S = (int(v) for v in config.split('x'))
        g = torch.Generator(device=device).manual_seed(1234)
        obs = torch.empty((B, T, S), dtype=torch.float32, device=device)
        for start in range(0, B, 64):
            x = torch.randn((min(64, B - start), T, S), device=device, generator=g)
            obs[start:start + x.shape[0]] = torch.log_softmax(x, dim=-1)
        A = torch.from_numpy(cases.log_triangular_transition(S)).to(device)
        init = torch.full((S,), float(np.log(1.0 / S + np.finfo(np.float32).tiny)), device=device)
        frames = torch.full((B,), T, dtype=torch.int32, device=device)
        op(obs[:, :8].contiguous(), torch.full((B,), 8, dtype=torch.int32, device=device), A, init)   # warm-up
        torch.cuda.synchronize()
        times = []
        for _ in range(args.reps):
            start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            start.record()
            indices = op(obs, frames, A, init)
            end.record()
            torch.cuda.synchronize()
            times.append(start.elapsed_time(end))
        ms = float(np.median(times))
        print(json.dumps({
            'impl': 'reference_cuda', 'kernel': 'viterbi_make_trellis_kernel + viterbi_backtrace_trellis_kernel (unmodified, sm_100a)',
            'batch': B, 'frames': T, 'states': S, 'ms': ms, 'timesteps_per_s': B * T / (ms * 1e-3),
            'frac_of_fp32_maxplus_roofline': 2.0 * S * S * B * T / (ms * 1e-3) / (148 * 128 * 1.965e9),
            'checksum': int(indices.sum(dtype=torch.int64)), 'reps_ms': times}), flush=True)
        del obs, indices


if __name__ == '__main__':
    main()
