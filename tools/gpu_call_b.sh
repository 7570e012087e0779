#!/bin/bash
# round 2, GPU call B: parity of the round-2 pieces first (bounded), then cluster vs split numbers, then round-1 suite
out=gpurun_out/r02b; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
date +%s > $out/t0
echo "== cluster kernel parity"; timeout 240 python -m pytest tests/test_gpu_round2.py -x -q -k "cluster_kernel" > $out/pytest_cluster.log 2>&1; echo "rc=$?" >> $out/pytest_cluster.log; tail -4 $out/pytest_cluster.log
echo "== round-2 tests (all but cluster parity)"; timeout 900 python -m pytest tests/test_gpu_round2.py -q --timeout 400 -k "not cluster_kernel" --durations=15 > $out/pytest_round2.log 2>&1; echo "rc=$?" >> $out/pytest_round2.log; tail -30 $out/pytest_round2.log
echo "== cluster vs split, 64 sequences"
for k in split cluster; do timeout 200 python bench.py --batch 64 --kernel $k --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/bench_b64_$k.json 2> $out/bench_b64_$k.err; python -c "
import json,sys
d=json.loads(open('$out/bench_b64_$k.json').read().strip().splitlines()[-1]); print('$k', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms'])" ; done 2>&1 | tee $out/cluster_vs_split.txt
for bonus in 0 2 3; do TORBI_B200_CLUSTER_BONUS=$bonus timeout 200 python bench.py --batch 64 --kernel cluster --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cluster bonus $bonus', d['value'], d['roofline']['frac'])" ; done 2>&1 | tee $out/cluster_bonus.txt
for b in 32 128 256; do for k in split cluster; do timeout 200 python bench.py --batch $b --kernel $k --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b $k', d['value'], d['roofline']['frac'])" ; done; done 2>&1 | tee $out/cluster_batches.txt
echo "== cluster phases"; MODES="--kernel=cluster" BATCH=64 timeout 200 tools/profile_phases.sh run > $out/cluster_phases.txt 2>&1; head -12 $out/cluster_phases.txt
echo "== microbench"; timeout 200 tools/_build/microbench_sync > $out/microbench_sync.txt 2>&1; tail -40 $out/microbench_sync.txt
echo "== round-1 tests"; timeout 800 python -m pytest tests/test_gpu_parity.py -q --timeout 300 > $out/pytest_round1.log 2>&1; echo "rc=$?" >> $out/pytest_round1.log; tail -5 $out/pytest_round1.log
date +%s > $out/t1
