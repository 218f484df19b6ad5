mkdir -p gpurun_out
python -m pytest tests/test_gpu_system.py tests/test_gpu_ions.py tests/test_gpu_parallel.py -m gpu -x -q 2>&1 | tail -3
python scripts/denopt_bench.py > gpurun_out/denopt_bench.log 2>&1; cut -c1-330 gpurun_out/denopt_bench.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_denopt.csv python scripts/denopt_profile.py 256 6 > gpurun_out/denopt_prof.log 2>&1
tail -3 gpurun_out/denopt_prof.log
python profiles/summarize_launches.py gpurun_out/launches_denopt.csv > gpurun_out/launches_denopt.txt; head -12 gpurun_out/launches_denopt.txt; grep "k_ion\|total" gpurun_out/launches_denopt.txt
