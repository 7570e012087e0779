#!/bin/bash
# flow kernel: parity, phase profile, timings
out=gpurun_out/r02g; mkdir -p $out
echo "== flow parity"; timeout 300 python -m pytest tests/test_gpu_round2.py -q -x --timeout 100 -k "flow or packed_host" > $out/pytest_flow.log 2>&1; echo "rc=$?" >> $out/pytest_flow.log; tail -3 $out/pytest_flow.log
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong"
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel flow $B 2>/dev/null | python -c "$show" "c1 flow" | tee $out/flow_timings.txt
timeout 100 python bench.py --batch 1 --frames 50000 --states 4096 --kernel flow --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "c5/20 flow" | tee -a $out/flow_timings.txt
for b in 4 16; do timeout 100 python bench.py --batch $b --frames 2000 --states 1440 --kernel flow $B 2>/dev/null | python -c "$show" "batch $b flow"; done | tee -a $out/flow_timings.txt
for s in 256 640 2048 8192; do timeout 100 python bench.py --batch 1 --frames 4000 --states $s --kernel flow $B 2>/dev/null | python -c "$show" "states $s flow"; done | tee -a $out/flow_timings.txt
for g in 0 1; do TORBI_B200_FLOW_DIRECT=$g timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel flow $B 2>/dev/null | python -c "$show" "c1 flow direct=$g"; done | tee -a $out/flow_timings.txt
echo "== phases"
echo "S=1440"; TORBI_B200_LIBRARY=$PWD/tools/_build/libtorbi_b200_prof.so timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel flow --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>&1 | grep "^cta" | sort | uniq | head -3 | tee $out/flow_phases.txt
echo "S=4096"; TORBI_B200_LIBRARY=$PWD/tools/_build/libtorbi_b200_prof.so timeout 100 python bench.py --batch 1 --frames 5000 --states 4096 --kernel flow --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>&1 | grep "^cta" | sort | uniq | head -3 | tee -a $out/flow_phases.txt
