#!/bin/bash
out=gpurun_out/r02k; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong"
echo "== flow parity"; timeout 300 python -m pytest tests/test_gpu_round2.py -q -x --timeout 100 -k "flow" > $out/pytest_flow.log 2>&1; echo "rc=$?" >> $out/pytest_flow.log; tail -3 $out/pytest_flow.log
for s in 640 1440 2048; do timeout 100 python bench.py --batch 1 --frames 4000 --states $s --kernel flow $B 2>/dev/null | python -c "$show" "states $s flow"; done | tee $out/flow_timings2.txt
echo "S=1440 phases"; TORBI_B200_LIBRARY=$PWD/tools/_build/libtorbi_b200_prof.so timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel flow --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>&1 | grep "^cta" | sort | uniq | head -2 | tee $out/flow_phases2.txt
