#!/bin/bash
# Other BASELINE configs, batch-size sweep (the per-GPU shares of a 512-sequence batch) and sanitizers.
cd "$(dirname "$0")/.."
rm -f gpurun_out/configs.jsonl
python tools/bench_configs.py --configs 1 3 2>&1 | grep "^{"
python tools/bench_configs.py --configs 4 --scale-c4 0.15 2>&1 | grep "^{"
python tools/bench_configs.py --configs 5 --frames-c5 1000000 2>&1 | grep "^{"
for b in 16 64 128 256; do
  python bench.py --no-cpu-baseline --no-e2e --batch $b --steps 3 --warmup 2 2>/dev/null | tail -1 >> gpurun_out/batch_sweep.jsonl
done
python - <<'PY'
import json
for line in open('gpurun_out/batch_sweep.jsonl'):
    d = json.loads(line)
    print(d['config']['batch_per_gpu'], round(d['value']), d['roofline']['kernel'], round(d['roofline']['frac'], 3))
PY
for tool in memcheck racecheck synccheck initcheck; do
  timeout 280 compute-sanitizer --tool $tool python tests/sanitizer_workload.py > gpurun_out/sanitizer_$tool.log 2>&1
  echo "$tool rc=$? ok=$(grep -c '^ok' gpurun_out/sanitizer_$tool.log) $(grep SUMMARY gpurun_out/sanitizer_$tool.log)"
done
