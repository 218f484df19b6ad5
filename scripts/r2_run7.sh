#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dbg_geom.py > gpurun_out/r2g_geom.log 2>&1
grep -c call gpurun_out/r2g_geom.log; grep "<<<<" gpurun_out/r2g_geom.log | head -3 | cut -c1-400
timeout 1500 python -m pytest tests -q -m gpu -x --deselect tests/test_gpu_geometry.py > gpurun_out/r2g_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -8 gpurun_out/r2g_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err
tail -c 600 gpurun_out/r2g_bench.err
