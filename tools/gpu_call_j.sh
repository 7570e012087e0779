#!/bin/bash
# the reference's own CUDA operator timed on the B200; compute-sanitizer over every kernel
out=gpurun_out/r02j; mkdir -p $out
echo "== reference CUDA op"; timeout 400 python tools/time_reference_cuda.py > $out/reference_cuda.jsonl 2> $out/reference_cuda.err; cut -c1-330 $out/reference_cuda.jsonl; tail -2 $out/reference_cuda.err | cut -c1-300
echo "== sanitizers"; TOOLS="memcheck synccheck racecheck" timeout 1500 tools/sanitize.sh run > $out/sanitizer.txt 2>&1; cat $out/sanitizer.txt | cut -c1-200
