#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fastfft.py tests/test_gpu_system.py -q -m gpu -x > gpurun_out/r2n_pytest.log 2>&1
tail -5 gpurun_out/r2n_pytest.log
python scripts/stage_profile.py 256 2>&1 | tail -40
timeout 600 python scripts/denopt_bench.py 2>&1 | tail -8
