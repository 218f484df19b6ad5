#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parallel.py -q -m gpu -x > gpurun_out/r2k_pytest.log 2>&1
tail -25 gpurun_out/r2k_pytest.log
