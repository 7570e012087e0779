#!/bin/bash
# round 2, GPU call F: flow-kernel timings against the wide / split kernels
out=gpurun_out/r02f; mkdir -p $out
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong"
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "", d["roofline"]["backtrace_ms"] if d else "")'
for k in flow wide; do
  timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel $k $B 2>$out/err_c1_$k.txt | python -c "$show" "c1 $k"
  timeout 150 python bench.py --batch 1 --frames 100000 --states 4096 --kernel $k --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>$out/err_c5_$k.txt | python -c "$show" "c5/10 $k"
done 2>&1 | tee $out/flow_vs_wide.txt
for b in 2 4 8 16; do for k in flow wide split; do timeout 100 python bench.py --batch $b --frames 2000 --states 1440 --kernel $k $B 2>/dev/null | python -c "$show" "batch $b $k"; done; done 2>&1 | tee $out/flow_small_batches.txt
for s in 256 640 2048 8192; do for k in flow wide; do timeout 100 python bench.py --batch 1 --frames 4000 --states $s --kernel $k $B 2>/dev/null | python -c "$show" "states $s $k"; done; done 2>&1 | tee $out/flow_states.txt
