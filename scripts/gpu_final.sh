mkdir -p gpurun_out
python -m pytest tests/test_gpu_fastfft.py -m gpu -x -q -k "full_size or wt_family" 2>&1 | tail -4
python __graft_entry__.py smoke 2>&1 | tail -2
ncu --set full --clock-control none --import-source on -k regex:'zinv_kernel|zfwd_kernel|xmix_kernel|spass_kernel' -s 15 -c 5 -f -o gpurun_out/prof_wt python - > gpurun_out/ncu_wt.log 2>&1 <<'PY'
import torch, sys
sys.path.insert(0, '.')
import profess_ad_b200.functionals as F
from profess_ad_b200.synthetic import smooth_supercell
dev = torch.device('cuda:0')
box, den = smooth_supercell(256, 4, device=dev)
for _ in range(5):
    x = den.requires_grad_(True); E = F.WangTeter(box, x); torch.autograd.grad(E, x); den.requires_grad_(False)
torch.cuda.synchronize()
PY
tail -1 gpurun_out/ncu_wt.log
ncu -i gpurun_out/prof_wt.ncu-rep --page raw --csv > gpurun_out/prof_wt_raw.csv
python profiles/ncu_summary.py gpurun_out/prof_wt_raw.csv
