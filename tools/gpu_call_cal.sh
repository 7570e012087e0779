#!/bin/bash
# kernel selection at other state counts: every forced kernel against what `auto` picks (device-resident, T = 500)
out=gpurun_out/r02cal; mkdir -p $out
run() { timeout 120 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-e2e --no-configs --no-strong --frames 500 "$@" 2>/dev/null | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); print('$*', '->', d['roofline']['kernel'], round(d['value']), 'ts/s', round(d['roofline']['kernel_ms'],2), 'ms')
except Exception as e:
    print('$*', '-> failed / unsupported')"; }
for S in 256 4096; do
  for B in 512 64; do
    run --states $S --batch $B
    run --states $S --batch $B --kernel tiled --tile-rows 64
    run --states $S --batch $B --kernel tiled --tile-rows 32
    run --states $S --batch $B --kernel tiled --tile-rows 16
    run --states $S --batch $B --kernel split
  done
done 2>&1 | tee $out/calibration.txt
run --states 640 --batch 512 | tee -a $out/calibration.txt
run --states 640 --batch 512 --kernel tiled --tile-rows 64 | tee -a $out/calibration.txt
run --states 640 --batch 512 --kernel tiled --tile-rows 32 | tee -a $out/calibration.txt
run --states 640 --batch 512 --kernel split | tee -a $out/calibration.txt
