mkdir -p gpurun_out
python scripts/denopt_bench.py > gpurun_out/denopt_bench.log 2>&1; cat gpurun_out/denopt_bench.log | cut -c1-400
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_denopt.csv python scripts/denopt_profile.py 256 6 > gpurun_out/denopt_prof.log 2>&1
cat gpurun_out/denopt_prof.log | tail -5
python profiles/summarize_launches.py gpurun_out/launches_denopt.csv | head -40
