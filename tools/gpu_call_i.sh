#!/bin/bash
# wide kernel with 8 loads in flight per lane on streamed rows; file pipeline with kept pinned buffers
out=gpurun_out/r02i; mkdir -p $out
show='import json,sys
lines=sys.stdin.read().strip().splitlines()
d=json.loads(lines[-1]) if lines else None
print(sys.argv[1], d["value"] if d else "FAILED", d["roofline"]["kernel"] if d else "", d["roofline"]["kernel_ms"] if d else "")'
echo "== wide parity"; timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_round2.py -q -x --timeout 200 -k "wide or config5 or very_wide or banded_hint" > $out/pytest_wide.log 2>&1; echo "rc=$?" >> $out/pytest_wide.log; tail -3 $out/pytest_wide.log
timeout 150 python bench.py --batch 1 --frames 50000 --states 4096 --kernel wide --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "c5/20 wide" | tee $out/wide_timings.txt
for s in 2048 8192; do timeout 100 python bench.py --batch 1 --frames 4000 --states $s --kernel wide --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "states $s wide"; done | tee -a $out/wide_timings.txt
timeout 100 python bench.py --batch 4 --frames 4000 --states 4096 --kernel wide --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-configs --no-strong 2>/dev/null | python -c "$show" "S=4096 batch 4 wide" | tee -a $out/wide_timings.txt
echo "== config 4 leg"; timeout 300 python - <<'PY' 2>&1 | tail -3 | tee $out/c4.txt
import sys, json, types
sys.argv = ['bench.py']
import bench, torch
args = bench.parse_args()
torch.cuda.set_device(0)
barrier = torch.cuda.synchronize
for files in (2000, 6000):
    args.config4_files = files
    r = bench.config4_leg(args, 0, 1, 0, torch.device('cuda:0'), barrier, lambda *v: [float(x) for x in v])
    print(json.dumps({k: r[k] for k in ('value', 'seconds', 'files', 'read_gbs')}))
PY
