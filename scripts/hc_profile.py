"""RevisedHuangCarter E+V at 128^3 (single GPU): a few evaluations for a launch list."""
import os, sys, numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import profess_ad_b200.functionals as F
from profess_ad_b200.synthetic import smooth_supercell
dev = torch.device('cuda:0')
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
box, den = smooth_supercell(n, n // 64, device=dev)
tab = np.load(os.path.join(ROOT, 'tests', 'golden', 'hc_table.npz'))
hc = F.RevisedHuangCarter((0.45, 0.10, 2.0 / 3.0, 1.15), kernel=torch.from_numpy(tab['revhc']))
for _ in range(4):
    x = den.requires_grad_(True); E = hc.forward(box, x); torch.autograd.grad(E, x); den.requires_grad_(False)
torch.cuda.synchronize()
print('nodes', hc.last_n_nodes, 'E', E.item())
