"""CPU, world_size 2, gloo: the N > 1 host logic (LPT sharding agreed without
communication, gather of the index rows).  The decode itself is stood in for by
the C oracle -- the GPU parity tests cover the kernels."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import cases


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world_size, port, result_file):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        from oracle import oracle
        from torbi_b200 import distributed
        rng = np.random.default_rng(77)                      # same inputs on every rank
        B, T, S = 7, 24, 48
        obs = cases.log_softmax(rng.standard_normal((B, T, S)))
        A = cases.log_softmax(rng.standard_normal((S, S)))
        init = cases.log_softmax(rng.standard_normal((S,)))
        frames = np.array([24, 3, 17, 24, 9, 1, 20], np.int32)

        items = distributed.rank_shard(frames.tolist())
        local = oracle.c_decode(obs[items], frames[items], A, init)
        full = distributed.gather_indices(torch.from_numpy(local), items, B)

        want = oracle.c_decode(obs, frames, A, init)
        assert np.array_equal(full.numpy(), want)
        everyone = [None] * world_size
        dist.all_gather_object(everyone, items)
        assert sorted(sum(everyone, [])) == list(range(B))   # a partition of the batch
        loads = [int(frames[it].sum()) for it in everyone]
        assert max(loads) <= 0.65 * int(frames.sum())         # balanced by length
        # files of a corpus dealt by size: a partition, balanced, the same on every rank without a message;
        # the per-rank driver hands exactly that share to from_files_to_files
        import pathlib
        import torbi_b200
        root = pathlib.Path(result_file).parent
        inputs, outputs = [], []
        for i, n in enumerate([40, 3, 17, 25, 9, 31, 2]):
            inputs.append(root / f'in{i}.pt')
            outputs.append(root / f'out{i}.pt')
            if rank == 0:
                torch.save(torch.zeros(n, 8), inputs[-1])
        dist.barrier()
        mine_in, mine_out = distributed.shard_files(inputs, outputs)
        shares = [None] * world_size
        dist.all_gather_object(shares, [str(f) for f in mine_in])
        assert sorted(sum(shares, [])) == sorted(str(f) for f in inputs)
        assert [f.name.replace('in', 'out') for f in mine_in] == [f.name for f in mine_out]
        seen = []
        original = torbi_b200.from_files_to_files
        torbi_b200.from_files_to_files = lambda i, o, *a, **k: seen.append((list(i), list(o), k.get('gpu')))
        try:
            distributed.from_files_to_files(inputs, outputs, log_probs=True, gpu=rank)
        finally:
            torbi_b200.from_files_to_files = original
        assert seen == [(mine_in, mine_out, rank)]
        if rank == 0:
            with open(result_file, 'w') as f:
                f.write('ok')
    finally:
        dist.destroy_process_group()


def test_rank_sharding_and_gather_world_size_2(tmp_path):
    result = tmp_path / 'result.txt'
    mp.spawn(_worker, args=(2, _free_port(), str(result)), nprocs=2, join=True)
    assert result.read_text() == 'ok'
