#!/bin/bash
# 8 GPUs of one box: the contract bench (weak + strong + end to end + config 4 sharded) and the host->device ceiling
out=gpurun_out/r02n8; mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8 > $out/gpus.txt; nproc >> $out/gpus.txt; numactl -H 2>/dev/null | head -12 >> $out/gpus.txt
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > $out/bench_n8.json 2> $out/bench_n8.err; echo "rc=$?"; cut -c1-300 $out/bench_n8.json; tail -2 $out/bench_n8.err | cut -c1-300
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tools/h2d_ceiling.py --gb 4 --reps 4 > $out/h2d_ceiling_n8.jsonl 2> $out/h2d_ceiling_n8.err; cat $out/h2d_ceiling_n8.jsonl | cut -c1-400
