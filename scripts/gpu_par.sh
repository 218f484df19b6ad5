mkdir -p gpurun_out
python -m pytest tests/test_gpu_parallel.py -m gpu -x -q > gpurun_out/pytest_par.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_par.log
tail -30 gpurun_out/pytest_par.log
