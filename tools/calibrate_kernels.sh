run() { timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', '->', d['roofline']['kernel'], round(d['value']), 'ts/s', round(d['roofline']['kernel_ms'],2), 'ms')"; }
for B in 1 4 16 64 128; do run --batch $B --kernel wide; run --batch $B --kernel tiled; run --batch $B; done
run --batch 1 --states 4096 --kernel wide; run --batch 1 --states 4096 --kernel tiled; run --batch 4 --states 4096 --kernel wide
