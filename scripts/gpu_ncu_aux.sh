mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'k_lbfgs_pass1|k_lbfgs_pass2|k_ion_spectrum|k_ion_force_partial|k_chi_project' -c 8 -f -o gpurun_out/prof_aux python scripts/denopt_profile.py 256 3 > gpurun_out/ncu_aux.log 2>&1
tail -2 gpurun_out/ncu_aux.log
ncu -i gpurun_out/prof_aux.ncu-rep --page raw --csv > gpurun_out/prof_aux_raw.csv
python profiles/ncu_summary.py gpurun_out/prof_aux_raw.csv
