#!/bin/bash
out=gpurun_out/r02w; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
echo "== wide parity"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout 200 -k "wide or config5 or very_wide or golden" > $out/pytest_wide.log 2>&1; echo "rc=$?" >> $out/pytest_wide.log; tail -3 $out/pytest_wide.log
timeout 150 python bench.py --batch 1 --frames 50000 --states 4096 --kernel wide --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "c5/20 wide (raw emission)" | tee $out/wide_raw_emission.txt
for b in 1 4 8; do timeout 100 python bench.py --batch $b --frames 2000 --states 1440 --kernel wide --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "S=1440 batch $b wide"; done | tee -a $out/wide_raw_emission.txt
timeout 100 python bench.py --batch 1 --frames 4000 --states 8192 --kernel wide --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "S=8192 wide" | tee -a $out/wide_raw_emission.txt
