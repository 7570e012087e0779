#!/bin/bash
out=gpurun_out/r02w; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
echo "== wide parity"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout 200 -k "wide or config5 or very_wide" > $out/pytest_wide.log 2>&1; echo "rc=$?" >> $out/pytest_wide.log; tail -3 $out/pytest_wide.log
for v in 4 2 6; do
  lib=$PWD/torbi_b200/libtorbi_b200.so; [ $v != 4 ] && lib=$PWD/tools/_build/libtorbi_b200_wide$v.so
  TORBI_B200_LIBRARY=$lib timeout 150 python bench.py --batch 1 --frames 50000 --states 4096 --kernel wide --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "c5/20 wide ahead=$v"
  TORBI_B200_LIBRARY=$lib timeout 100 python bench.py --batch 1 --frames 4000 --states 8192 --kernel wide --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "S=8192 wide ahead=$v"
  TORBI_B200_LIBRARY=$lib timeout 100 python bench.py --batch 4 --frames 4000 --states 4096 --kernel wide --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "S=4096 batch 4 wide ahead=$v"
done | tee $out/wide_ahead.txt
