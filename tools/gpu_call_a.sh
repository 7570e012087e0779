#!/bin/bash
# round 2, first GPU call: new kernels' parity first (bounded), then the whole suite, microbenchmarks,
# the reference's own CUDA op, ncu captures of the wide kernel, and first bench numbers.
out=gpurun_out/r02a; mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > $out/gpu.txt 2>&1
echo "== cluster kernel parity (bounded)"; timeout 300 python -m pytest tests/test_gpu_round2.py -x -q -k "cluster_kernel" > $out/pytest_cluster.log 2>&1; echo "rc=$?" >> $out/pytest_cluster.log; tail -5 $out/pytest_cluster.log
echo "== round-2 tests"; timeout 1500 python -m pytest tests/test_gpu_round2.py -q -k "not cluster_kernel" > $out/pytest_round2.log 2>&1; echo "rc=$?" >> $out/pytest_round2.log; tail -15 $out/pytest_round2.log
echo "== round-1 tests"; timeout 900 python -m pytest tests/test_gpu_parity.py -q > $out/pytest_round1.log 2>&1; echo "rc=$?" >> $out/pytest_round1.log; tail -5 $out/pytest_round1.log
echo "== microbench"; timeout 300 tools/_build/microbench_sync > $out/microbench_sync.txt 2>&1; tail -40 $out/microbench_sync.txt
echo "== reference CUDA op"; timeout 600 python tools/time_reference_cuda.py > $out/reference_cuda.jsonl 2> $out/reference_cuda.err; cat $out/reference_cuda.jsonl
echo "== cluster vs split, 64 sequences"
for k in split cluster; do timeout 300 python bench.py --batch 64 --kernel $k --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/bench_b64_$k.json 2> $out/bench_b64_$k.err; python -c "
import json,sys
d=json.loads(open('$out/bench_b64_$k.json').read().strip().splitlines()[-1]); print('$k', d['value'], d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['kernel_ms'])" ; done
for bonus in 0 2 3; do TORBI_B200_CLUSTER_BONUS=$bonus timeout 300 python bench.py --batch 64 --kernel cluster --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cluster bonus $bonus', d['value'], d['roofline']['frac'])" ; done 2>&1 | tee $out/cluster_bonus.txt
for b in 32 128 256; do for k in split cluster; do timeout 300 python bench.py --batch $b --kernel $k --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b $k', d['value'], d['roofline']['frac'])" ; done; done 2>&1 | tee $out/cluster_batches.txt
echo "== cluster phases"; MODES="--kernel=cluster" BATCH=64 timeout 300 tools/profile_phases.sh run > $out/cluster_phases.txt 2>&1; cat $out/cluster_phases.txt | head -12
echo "== ncu wide"
for cfg in "1440 c1" "4096 c5"; do set -- $cfg; timeout 600 ncu --set full --clock-control none --import-source on -k regex:forward_wide -c 1 -f -o $out/wide_$2 python bench.py --batch 1 --frames 400 --states $1 --kernel wide --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/ncu_wide_$2.log 2>&1; tail -2 $out/ncu_wide_$2.log; done
echo "== bench N=1 (with configs)"; timeout 900 python bench.py --steps 5 --warmup 3 > $out/bench_n1.json 2> $out/bench_n1.err; tail -c 3000 $out/bench_n1.json; tail -5 $out/bench_n1.err
