#!/bin/bash
# round 2, GPU call C: parity of everything added in round 2 (verbose, per-test timeouts), then flow-kernel timings
out=gpurun_out/r02c; mkdir -p $out
date +%s > $out/t0
echo "== flow kernel"; timeout 500 python -m pytest tests/test_gpu_round2.py -q -x --timeout 120 -k "flow" > $out/pytest_flow.log 2>&1; echo "rc=$?" >> $out/pytest_flow.log; tail -6 $out/pytest_flow.log
echo "== cluster kernel"; timeout 300 python -m pytest tests/test_gpu_round2.py -q -x --timeout 100 -k "cluster_kernel" > $out/pytest_cluster.log 2>&1; echo "rc=$?" >> $out/pytest_cluster.log; tail -4 $out/pytest_cluster.log
echo "== chunk / entropy / evaluate"; timeout 600 python -m pytest tests/test_gpu_chunk.py -v --timeout 200 > $out/pytest_chunk.log 2>&1; echo "rc=$?" >> $out/pytest_chunk.log; grep -c PASSED $out/pytest_chunk.log; grep "FAILED\|Error\|rc=" $out/pytest_chunk.log | head -20
echo "== config 1 / 5 timings: flow vs wide"
for k in flow wide; do
  timeout 120 python bench.py --batch 1 --frames 2000 --states 1440 --kernel $k --steps 10 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>$out/bench_c1_$k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c1 $k', d['value'], 'timesteps/s', d['roofline']['kernel'], d['roofline']['kernel_ms'], 'ms fwd', d['roofline']['backtrace_ms'], 'ms back')"
  timeout 200 python bench.py --batch 1 --frames 100000 --states 4096 --kernel $k --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>$out/bench_c5_$k.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5/10 $k', d['value'], 'timesteps/s', d['roofline']['kernel'], d['roofline']['kernel_ms'], 'ms fwd', d['roofline']['backtrace_ms'], 'ms back')"
done 2>&1 | tee $out/flow_vs_wide.txt
for b in 2 4 8 16; do for k in flow wide split; do timeout 120 python bench.py --batch $b --frames 2000 --states 1440 --kernel $k --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('batch $b $k', d['value'], d['roofline']['kernel_ms'])"; done; done 2>&1 | tee $out/flow_small_batches.txt
for s in 256 640 2048 8192; do for k in flow wide; do timeout 120 python bench.py --batch 1 --frames 4000 --states $s --kernel $k --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('states $s $k', d['value'], d['roofline']['kernel_ms'])"; done; done 2>&1 | tee $out/flow_states.txt
echo "== round-2 tests (rest)"; timeout 1500 python -m pytest tests/test_gpu_round2.py -v --timeout 400 -k "not cluster_kernel and not flow_kernel" --durations=10 > $out/pytest_round2.log 2>&1; echo "rc=$?" >> $out/pytest_round2.log; grep -c PASSED $out/pytest_round2.log; grep "FAILED\|rc=\|Timeout" $out/pytest_round2.log | head -20; tail -14 $out/pytest_round2.log
date +%s > $out/t1
