#!/bin/bash
out=gpurun_out/r02r; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
B="--steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong"
for r in 1 2 4 8 16 32; do
  TORBI_B200_FLOW_REPLICAS=$r timeout 100 python bench.py --batch 1 --frames 4000 --states 1440 --kernel flow $B 2>/dev/null | python -c "$show" "S=1440 replicas $r"
done | tee $out/replicas.txt
for r in 4 16; do
  TORBI_B200_FLOW_REPLICAS=$r timeout 100 python bench.py --batch 1 --frames 4000 --states 640 --kernel flow $B 2>/dev/null | python -c "$show" "S=640 replicas $r"
  TORBI_B200_FLOW_REPLICAS=$r timeout 100 python bench.py --batch 1 --frames 4000 --states 2048 --kernel flow $B 2>/dev/null | python -c "$show" "S=2048 replicas $r"
done | tee -a $out/replicas.txt
echo "phases, 8 replicas"; TORBI_B200_FLOW_REPLICAS=8 TORBI_B200_LIBRARY=$PWD/tools/_build/libtorbi_b200_prof.so timeout 100 python bench.py --batch 1 --frames 2000 --states 1440 --kernel flow --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>&1 | grep "^cta" | sort | uniq | head -2 | tee -a $out/replicas.txt
