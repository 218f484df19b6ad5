#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dbg_geom2.py native > gpurun_out/r2h_geom_native.log 2>&1
timeout 600 python scripts/dbg_geom2.py torch > gpurun_out/r2h_geom_torch.log 2>&1
tail -3 gpurun_out/r2h_geom_native.log gpurun_out/r2h_geom_torch.log
timeout 600 python -m pytest tests/test_gpu_pme.py tests/test_gpu_geometry.py -q -m gpu > gpurun_out/r2h_pytest.log 2>&1
tail -5 gpurun_out/r2h_pytest.log
