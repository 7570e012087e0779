#!/bin/bash
# round 2, GPU call E: parity of the round-2 pieces (flow first), then the whole GPU suite with durations
out=gpurun_out/r02e; mkdir -p $out
echo "== flow kernel"; timeout 400 python -m pytest tests/test_gpu_round2.py -q -x --timeout 120 -k "flow" > $out/pytest_flow.log 2>&1; echo "rc=$?" >> $out/pytest_flow.log; tail -4 $out/pytest_flow.log
echo "== round-2 tests (rest)"; timeout 1300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_chunk.py -q --timeout 400 -k "not flow" --durations=12 > $out/pytest_round2.log 2>&1; echo "rc=$?" >> $out/pytest_round2.log; grep "FAILED\|rc=\|Timeout" $out/pytest_round2.log | head -20; tail -18 $out/pytest_round2.log
