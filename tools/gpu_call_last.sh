#!/bin/bash
out=gpurun_out/r02last; mkdir -p $out
echo "== flow parity (final selection rule)"; timeout 100 python -m pytest tests/test_gpu_round2.py -q -x --timeout 60 -k "flow" > $out/pytest_flow_final.log 2>&1; echo "rc=$?" >> $out/pytest_flow_final.log; tail -2 $out/pytest_flow_final.log
