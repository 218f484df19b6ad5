#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_parallel.py tests/test_gpu_ions.py -q -m gpu -x > gpurun_out/r2o_pytest.log 2>&1
tail -6 gpurun_out/r2o_pytest.log
