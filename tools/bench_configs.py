#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs (1, 3, 4, 5) through the
public API, one JSON line each -> gpurun_out/configs.jsonl.  bench.py stays the
contract benchmark (config 2); these numbers feed DESIGN.md / README.md.

    python tools/bench_configs.py [--configs 1 3 4 5] [--scale-c4 0.05] [--frames-c5 100000]
"""
import argparse
import json
import pathlib
import sys
import tempfile
import time

import numpy as np
import torch

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / 'tests' / 'golden'))

import cases  # noqa: E402
import torbi_b200  # noqa: E402
from torbi_b200 import _ffi  # noqa: E402


def timed(fn, repeats=3, warmup=1):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    times = []
    for _ in range(repeats):
        start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        start.record()
        out = fn()
        end.record()
        torch.cuda.synchronize()
        times.append(start.elapsed_time(end) * 1e-3)
    return out, float(np.median(times))


def log_softmax_noise(shape, seed):
    g = torch.Generator(device='cuda:0').manual_seed(seed)
    out = torch.empty(shape, dtype=torch.float32, device='cuda:0')
    step = max(1, (1 << 28) // (shape[1] * shape[2]))
    for b in range(0, shape[0], step):
        x = torch.randn((min(step, shape[0] - b),) + tuple(shape[1:]), device='cuda:0', generator=g)
        out[b:b + x.shape[0]] = torch.log_softmax(x, dim=-1)
    return out


def kernel_name():
    return _ffi.library().torbi_b200_last_kernel().decode()


def config1():
    obs = log_softmax_noise((1, 2000, 1440), 1234)
    out, seconds = timed(lambda: torbi_b200.from_probabilities(obs, log_probs=True, gpu=0))
    return dict(config=1, workload='S=1440 B=1 T=2000 uniform transition/initial', timesteps=2000,
                seconds=seconds, timesteps_per_s=2000 / seconds, kernel=kernel_name())


def config3():
    rng = np.random.default_rng(1234)
    B, S = 256, 1440
    frames = np.exp(rng.uniform(np.log(200), np.log(20000), size=B)).round().astype(np.int32)
    frames[0], frames[1] = 20000, 200
    T = int(frames.max())
    obs = log_softmax_noise((B, T, S), 1234)
    A = torch.from_numpy(cases.log_triangular_transition(S)).cuda()
    frames_t = torch.from_numpy(frames)
    out, seconds = timed(lambda: torbi_b200.from_probabilities(obs, frames_t, A, None, log_probs=True, gpu=0),
                         repeats=2)
    total = int(frames.sum())
    return dict(config=3, workload=f'S=1440 B=256 ragged 200..20000 frames (sum {total})', timesteps=total,
                seconds=seconds, timesteps_per_s=total / seconds, kernel=kernel_name(),
                padded_gb=B * T * S * 4 / 1e9)


def config4(scale):
    rng = np.random.default_rng(1234)
    S = 1440
    files = max(8, int(40000 * scale))
    frames = np.clip(np.exp(rng.normal(np.log(430), 0.55, size=files)), 50, 5000)
    frames = np.maximum(50, (frames * (20e6 * scale / frames.sum())).round()).astype(np.int64)
    root = pathlib.Path(tempfile.mkdtemp(prefix='torbi_c4_', dir='/dev/shm'))
    inputs, outputs = [], []
    g = torch.Generator(device='cuda:0').manual_seed(4)
    for i, f in enumerate(frames):
        x = torch.log_softmax(torch.randn((int(f), S), device='cuda:0', generator=g), dim=-1).cpu()
        torch.save(x, root / f'{i}.pt')
        inputs.append(root / f'{i}.pt')
        outputs.append(root / f'{i}.out.pt')
    torch.save(torch.from_numpy(cases.triangular_transition(S)), root / 'transition.pt')
    torch.cuda.synchronize()
    start = time.perf_counter()
    torbi_b200.from_files_to_files(inputs, outputs, root / 'transition.pt', log_probs=True, gpu=0)
    seconds = time.perf_counter() - start
    total = int(frames.sum())
    check = torch.load(outputs[0])
    assert check.shape == (int(frames[0]),) and check.dtype == torch.int32
    import shutil
    shutil.rmtree(root, ignore_errors=True)
    return dict(config=4, workload=f'{files} files, {total} timesteps, S=1440 on /dev/shm (scale {scale} of the corpus), 1 GPU, '
                                   'end to end incl. torch.load / torch.save',
                timesteps=total, seconds=seconds, timesteps_per_s=total / seconds, kernel=kernel_name(),
                input_gb=total * S * 4 / 1e9)


def config5(T):
    S = 4096
    obs = log_softmax_noise((1, T, S), 5)
    g = torch.Generator(device='cuda:0').manual_seed(6)
    A = torch.log_softmax(torch.randn((S, S), device='cuda:0', generator=g), dim=1)
    out, seconds = timed(lambda: torbi_b200.from_probabilities(obs, None, A, None, log_probs=True, gpu=0),
                         repeats=2)
    return dict(config=5, workload=f'S=4096 B=1 T={T} dense transition (BASELINE config 5 is T=1e6)', timesteps=T,
                seconds=seconds, timesteps_per_s=T / seconds, kernel=kernel_name())


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--configs', type=int, nargs='+', default=[1, 3, 4, 5])
    parser.add_argument('--scale-c4', type=float, default=0.05)
    parser.add_argument('--frames-c5', type=int, default=100000)
    args = parser.parse_args()
    out_dir = ROOT / 'gpurun_out'
    out_dir.mkdir(exist_ok=True)
    runners = {1: config1, 3: config3, 4: lambda: config4(args.scale_c4), 5: lambda: config5(args.frames_c5)}
    with open(out_dir / 'configs.jsonl', 'a') as f:
        for c in args.configs:
            line = runners[c]()
            line['gpu'] = torch.cuda.get_device_name(0)
            print(json.dumps(line), flush=True)
            f.write(json.dumps(line) + '\n')
            torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
