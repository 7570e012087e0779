#!/bin/bash
# full GPU test suite + batch-size sweep (per-GPU shares of a 512-sequence batch) + config 4
cd "$(dirname "$0")/.."
timeout 500 python -m pytest tests -m gpu -x -q 2>&1 | tail -2 | tee gpurun_out/pytest_gpu.log
rm -f gpurun_out/batch_sweep.jsonl
for b in 16 32 64 128 192 256; do
  python bench.py --no-cpu-baseline --no-e2e --batch $b --steps 3 --warmup 2 2>/dev/null | tail -1 >> gpurun_out/batch_sweep.jsonl
done
python - <<'PY'
import json
for line in open('gpurun_out/batch_sweep.jsonl'):
    d = json.loads(line)
    print(d['config']['batch_per_gpu'], round(d['value']), d['roofline']['kernel'], round(d['roofline']['frac'], 3))
PY
timeout 100 python tools/bench_configs.py --configs 4 --scale-c4 0.15 2>&1 | grep "^{" | cut -c1-330
