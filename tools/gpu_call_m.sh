#!/bin/bash
out=gpurun_out/r02m; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "", d["roofline"]["frac"] if d else "")'
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong"
echo "== parity (plain emission path: half precision, S % 4 != 0, packed)"; timeout 400 python -m pytest tests/test_gpu_round2.py tests/test_gpu_parity.py -q -x --timeout 200 -k "half or unaligned or packed or S1441 or S333 or decode_matches_oracle" > $out/pytest.log 2>&1; echo "rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
timeout 100 python bench.py $B 2>/dev/null | python -c "$show" "config 2 tiled" | tee $out/timings.txt
timeout 100 python bench.py $B 2>/dev/null | python -c "$show" "config 2 tiled (again)" | tee -a $out/timings.txt
timeout 100 python bench.py --batch 64 $B 2>/dev/null | python -c "$show" "64 sequences split" | tee -a $out/timings.txt
timeout 100 python bench.py --batch 256 $B 2>/dev/null | python -c "$show" "256 sequences" | tee -a $out/timings.txt
timeout 100 python bench.py --backpointers $B 2>/dev/null | python -c "$show" "config 2 backpointer mode" | tee -a $out/timings.txt
