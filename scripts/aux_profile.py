"""A few evaluations of the secondary fused routes for a launch list (ncu --metrics gpu__time_duration.sum):
PerdewBurkeErnzerhof, Hartree, WangTeter and the IonElectron + Hartree + WGC99 + PZ term list at 256^3 (CUDA graphs off)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200 import _density_opt as D, _native
from profess_ad_b200.synthetic import smooth_supercell
dev = torch.device('cuda:0')
lib = _native.load_library()
lib.pad_set_option(b'graphs', 0)
box_h, den_h = smooth_supercell(256, 4)
box, den = box_h.to(dev), den_h.to(dev)
v_ext = -0.1 * torch.rand_like(den)
wgc = F.WangGovindCarter99()
T = D.describe_terms([F.IonElectron, F.Hartree, wgc.forward, F.PerdewZunger])
for _ in range(3):
    F.energy_and_potential(box, den, F.PerdewBurkeErnzerhof)
    F.energy_and_potential(box, den, F.Hartree)
    F.energy_and_potential(box, den, F.WangTeter)
    D.eval_total(box, den, v_ext, T)
torch.cuda.synchronize()
print('done')
