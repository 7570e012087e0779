#!/bin/bash
out=gpurun_out/r02ws; mkdir -p $out
echo "== phases ws=1"; MODES="--kernel=cluster" BATCH=64 timeout 200 tools/profile_phases.sh run 2>&1 | grep "^cta" | head -8 | tee $out/cluster_ws_phases2.txt
