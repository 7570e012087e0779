#!/bin/bash
out=gpurun_out/r02last; mkdir -p $out
timeout 100 python __graft_entry__.py smoke > $out/smoke_final.log 2>&1; echo "rc=$?"; tail -1 $out/smoke_final.log | cut -c1-300
