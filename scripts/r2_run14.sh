#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_fastfft.py -q -m gpu -x -k "graph or term_list" 2>&1 | tail -3
for g in 1 0; do
PAD_GRAPHS=$g timeout 600 python bench.py --steps 50 --warmup 8 --no-cpu-baseline --no-denopt 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('graphs $g: ms/step', round(d['ms_per_step'],4), d['launch_detail'], 'e2e', d['e2e']['value'], d['e2e']['repeats']); print('   also', json.dumps(d.get('also')))"
done
