import os, sys
import numpy as np, torch
sys.path.insert(0, '.')
import profess_ad_b200.functionals as F
from profess_ad_b200.system import System
from profess_ad_b200 import ion_utils as IU
import profess_ad_b200.system as SYS
g = np.load('tests/golden/geometry_li2_ions.npz')
terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]

def torch_path(box, coords, charges, Rc, Rd):
    mi, mj, shifts = IU._pair_list(box, coords, Rc)
    rho = torch.sum(charges) / torch.abs(torch.linalg.det(box))
    Zi, Zj = charges[mi], charges[mj]
    Qi = torch.scatter_add(charges, 0, mi, Zj)
    aux = (0.75 / np.pi) * Qi / rho
    Ra = aux.sign() * aux.abs().pow(1 / 3)
    r_ij = (coords[mj] + shifts @ box - coords[mi]).norm(p=2, dim=1)
    E_local = torch.sum(0.5 * Zi * Zj * torch.erfc(r_ij / Rd) / r_ij)
    E_corr = torch.sum(-np.pi * charges * rho * Ra.square() + np.pi * charges * rho * (Ra.square() - 0.5 * Rd * Rd) * torch.erf(Ra / Rd)
                       + np.sqrt(np.pi) * charges * rho * Ra * Rd * torch.exp(-Ra.square() / (Rd * Rd)) - charges.square() / np.sqrt(np.pi) / Rd)
    return E_local + E_corr

calls = [0]
orig = IU.ion_interaction_sum
def both(box, coords, charges, Rc, Rd):
    calls[0] += 1
    En = orig(box, coords, charges, Rc, Rd)
    with torch.enable_grad():
        c2 = coords.detach().clone().requires_grad_(True)
        b2 = box.detach().clone().requires_grad_(True)
        Et = torch_path(b2, c2, charges, Rc, Rd)
        gt = torch.autograd.grad(Et, (b2, c2))
    msg = 'call %d  E native %.12f torch %.12f diff %.2e' % (calls[0], En.item(), Et.item(), En.item() - Et.item())
    if coords.requires_grad or box.requires_grad:
        ins = [t for t in (box, coords) if t.requires_grad]
        gn = torch.autograd.grad(En, ins, retain_graph=True)
        k = 0
        if box.requires_grad:
            msg += '  dbox diff %.2e' % (gn[k] - gt[0]).abs().max().item(); k += 1
        if coords.requires_grad:
            msg += '  dcart diff %.2e (max %.2e)' % ((gn[k] - gt[1]).abs().max().item(), gt[1].abs().max().item())
    if abs(En.item() - Et.item()) > 1e-9 or 'e-0' in msg.split('dcart')[-1][:20]:
        msg += '   <<<<'
        msg += ' coords ' + str(coords.detach().cpu().numpy().tolist()) + ' Rc %r Rd %r' % (float(Rc), float(Rd))
    print(msg, flush=True)
    return En
SYS.ion_interaction_sum = both
ions = [['Li', 'tests/potentials/li.gga.recpot', torch.from_numpy(g['frac0'])]]
s = System(torch.from_numpy(g['box0_A']), tuple(int(n) for n in g['shape']), ions, terms, units='a', coord_type='fractional')
ok = s.optimize_geometry(g_maxiter=14, ntol=1e-9, ftol=0.02, stol=None, g_verbose=True)
print('ok', ok, s.last_geometry_optimization)
