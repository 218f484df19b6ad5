#!/bin/bash
for i in 1 2 3; do
timeout 900 python bench.py --no-cpu-baseline --no-denopt 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step', round(d['ms_per_step'],4), [round(x,4) for x in d['repeats_ms_per_step']], 'warm', d['warmup'], 'host', round(d['launch_detail']['host_enqueue_ms_per_step'],3), d['gpu_launches'], d['clocks']['reasons'])"
done
