#!/bin/bash
# Round evidence on one GPU: tests, the contract bench (both arms), the ncu launch list and full
# capture of the hot kernel, the other BASELINE configs.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; cut -c1-400 gpurun_out/bench_n1.json
python bench.py --impl reference > gpurun_out/bench_reference_n1.json 2> gpurun_out/bench_reference_n1.err; cut -c1-300 gpurun_out/bench_reference_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_launch_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"forward_tiled_hist|backtrace_recompute" -c 2 -f \
    -o gpurun_out/r01_tiled_hist python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
