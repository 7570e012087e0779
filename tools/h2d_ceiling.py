#!/usr/bin/env python
"""Host -> device copy ceiling with N concurrent ranks (one per GPU), with and without binding each
rank to the NUMA node of its GPU: plain cudaMemcpyAsync of one pinned buffer per rank, all ranks
copying at the same time.  This is the ceiling for the end-to-end number of bench.py at N GPUs
(every rank streams its 5.9 GB batch per step).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tools/h2d_ceiling.py [--gb 4] [--reps 4]
Rank 0 prints one JSON line per mode.
"""
import argparse
import json
import os
import pathlib
import sys
import time

import torch
import torch.distributed as dist

ROOT = pathlib.Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402  (bind_to_gpu_numa_node)


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument('--gb', type=float, default=4.0)
    parser.add_argument('--reps', type=int, default=4)
    args = parser.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    device = torch.device(f'cuda:{local}')
    if world > 1:
        dist.init_process_group('nccl', device_id=device)
    inherited = sorted(os.sched_getaffinity(0))
    numel = int(args.gb * (1 << 30)) // 4
    dst = torch.empty(numel, dtype=torch.float32, device=device)
    for mode in ('unbound', 'bound'):
        os.sched_setaffinity(0, inherited)
        binding = {'method': 'none'}
        if mode == 'bound':
            binding = bench.bind_to_gpu_numa_node(local, world)
        src = torch.empty(numel, dtype=torch.float32, pin_memory=True)
        src.fill_(1.0)                                           # first touch from the (bound) thread
        dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.reps):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        seconds = time.perf_counter() - t0
        if world > 1:
            dist.barrier()
        mine = torch.tensor([numel * 4 * args.reps / seconds / 1e9], dtype=torch.float64, device=device)
        rates = [torch.zeros_like(mine) for _ in range(world)]
        if world > 1:
            dist.all_gather(rates, mine)
        else:
            rates = [mine]
        if rank == 0:
            per_rank = [round(float(r), 2) for r in rates]
            print(json.dumps({'mode': mode, 'ranks': world, 'gb_per_copy': args.gb, 'reps': args.reps,
                              'gbs_per_rank': per_rank, 'gbs_min': min(per_rank), 'gbs_sum': round(sum(per_rank), 1),
                              'binding_rank0': binding, 'inherited_cpus': len(inherited),
                              'needed_gbs_per_rank_to_hide_under_the_kernel': round(5.9 / 0.138, 1)}), flush=True)
        del src
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
