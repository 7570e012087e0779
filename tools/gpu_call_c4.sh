#!/bin/bash
out=gpurun_out/r02c4; mkdir -p $out
TORBI_B200_PIPELINE_TIMING=1 timeout 300 python - <<'PY' 2>&1 | grep -v "it/s" | tee $out/c4_timing.txt
import sys, json
sys.argv = ['bench.py']
import bench, torch
args = bench.parse_args()
torch.cuda.set_device(0)
for files in (2000, 2000, 6000):
    args.config4_files = files
    r = bench.config4_leg(args, 0, 1, 0, torch.device('cuda:0'), torch.cuda.synchronize, lambda *v: [float(x) for x in v])
    print(json.dumps({k: r[k] for k in ('value', 'seconds', 'files', 'read_gbs')}), flush=True)
PY
