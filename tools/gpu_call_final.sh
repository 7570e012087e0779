#!/bin/bash
# final confirmation on one GPU: smoke, whole GPU suite, the contract bench (both arms)
out=gpurun_out/r02final; mkdir -p $out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke > $out/smoke.log 2>&1; echo "rc=$?"; tail -2 $out/smoke.log | cut -c1-400
echo "== GPU suite"; timeout 1200 python -m pytest tests -m gpu -x -q --timeout 400 > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log; tail -3 $out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "rc=$?"; python - <<'PY'
import json
d = json.loads(open('gpurun_out/r02final/bench_n1.json').read().strip().splitlines()[-1])
print('value', d['value'], 'e2e', d['e2e']['value'], 'fp16', d['e2e']['fp16_inputs']['value'], 'frac', d['roofline']['frac'], 'kernel_ms', d['roofline']['kernel_ms'])
for k, c in d['configs'].items():
    print(k, c['value'], c.get('kernel'), c.get('peak_hbm_gb'), c.get('latency_bound', {}).get('us_per_step'), c.get('seconds'))
print('cpu', d['cpu_baseline']['value'], 'launches', d['gpu_launches'], 'clocks', d['clocks'])
PY
echo "== reference arm"; timeout 400 python bench.py --impl reference > $out/bench_reference_n1.json 2> $out/bench_reference_n1.err; cut -c1-200 $out/bench_reference_n1.json
