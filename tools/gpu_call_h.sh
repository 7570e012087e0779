#!/bin/bash
# round 2 evidence on one GPU: whole GPU suite, the contract bench (both arms), ncu launch list + full captures
out=gpurun_out/r02h; mkdir -p $out
echo "== GPU suite"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 > $out/pytest_gpu.log 2>&1; echo "rc=$?" >> $out/pytest_gpu.log; tail -4 $out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err; echo "rc=$?"; cut -c1-600 $out/bench_n1.json; tail -3 $out/bench_n1.err
echo "== reference arm"; timeout 400 python bench.py --impl reference > $out/bench_reference_n1.json 2> $out/bench_reference_n1.err; cut -c1-300 $out/bench_reference_n1.json
echo "== ncu launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $out/r02_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --skip-config4 --skip-config3 --config5-frames 50000 > $out/ncu_launch_bench.log 2>&1; tail -2 $out/ncu_launch_bench.log; wc -l $out/r02_launches.csv
echo "== ncu full: tiled hist"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"forward_tiled_hist" -c 1 -f \
    -o $out/r02_tiled_hist python bench.py --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/ncu_full_tiled.log 2>&1; tail -1 $out/ncu_full_tiled.log
echo "== ncu full: flow (config 1)"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"forward_flow" -c 1 -f \
    -o $out/r02_flow_c1 python bench.py --batch 1 --frames 2000 --states 1440 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/ncu_full_flow.log 2>&1; tail -1 $out/ncu_full_flow.log
echo "== ncu full: wide (config 5, 20000 frames)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:"forward_wide" -c 1 -f \
    -o $out/r02_wide_c5 python bench.py --batch 1 --frames 20000 --states 4096 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/ncu_full_wide.log 2>&1; tail -1 $out/ncu_full_wide.log
echo "== ncu full: split (64 sequences)"; timeout 400 ncu --set full --clock-control none --import-source on -k regex:"forward_split" -c 1 -f \
    -o $out/r02_split_b64 python bench.py --batch 64 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline --no-configs --no-strong > $out/ncu_full_split.log 2>&1; tail -1 $out/ncu_full_split.log
ls -la $out
