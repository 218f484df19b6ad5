"""Native ion-ion sum (csrc/ionion.cu) against the reference's known answers (tests/test_ion_utils.py:12-147: CASTEP energies
of Al, Si, SiO2, Al2SiO5 and the NaCl Madelung constant, tests/golden/ion_ion_castep.json), and its closed-form derivatives
against autograd through the pair-list restatement (what the reference differentiates)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _cases(golden_dir):
    return json.load(open(os.path.join(golden_dir, 'ion_ion_castep.json')))


def test_castep_energies_and_madelung_constant(golden_dir):
    from profess_ad_b200.ion_utils import ion_interaction_sum
    doc = _cases(golden_dir)
    for c in doc['cases']:
        box = torch.tensor(c['box'], dtype=torch.double)
        cart = (torch.tensor(c['frac'], dtype=torch.double) @ box).to(DEV)
        z = torch.tensor(c['charges'], dtype=torch.double, device=DEV)
        E = ion_interaction_sum(box.to(DEV), cart, z, 12 * c['h_max'], 2 * c['h_max'])
        assert abs(E.item() - c['E']) / len(c['charges']) < 1e-10, (c['name'], E.item(), c['E'])
    m = doc['madelung']
    box = torch.tensor(m['box'], dtype=torch.double, device=DEV)
    Rc, Rd = 12 * m['h_max'], 2 * m['h_max']
    one = lambda k: torch.ones(k, dtype=torch.double, device=DEV)
    E_fcc = ion_interaction_sum(box, torch.tensor(m['fcc_cart'], dtype=torch.double, device=DEV), one(1), Rc, Rd)
    E_2 = ion_interaction_sum(box, torch.tensor(m['pair_cart'], dtype=torch.double, device=DEV), one(2), Rc, Rd)
    assert abs((4 * E_fcc - E_2).item() - m['value']) < 1e-10


@pytest.mark.parametrize('name', ['Si', 'SiO2'])
def test_derivatives_match_autograd_through_the_pair_list(name, golden_dir):
    """dE/dcoords and dE/dbox_vecs (coords = frac @ box: forces and the cell gradient the stress is made of) from the
    kernel's closed forms vs torch.autograd through the pair-list form on the CPU (float64)."""
    from profess_ad_b200 import ion_utils as IU
    c = [x for x in _cases(golden_dir)['cases'] if x['name'] == name][0]
    gen = torch.Generator().manual_seed(3)
    box0 = torch.tensor(c['box'], dtype=torch.double)
    frac = torch.tensor(c['frac'], dtype=torch.double) + 0.01 * torch.rand(len(c['frac']), 3, dtype=torch.double, generator=gen)
    z = torch.tensor(c['charges'], dtype=torch.double)
    Rc, Rd = 6 * c['h_max'], 2 * c['h_max']            # shorter cutoff: the CPU pair list stays small
    out = {}
    for dev in ('cpu', DEV):
        box = box0.clone().to(dev).requires_grad_(True)
        fr = frac.clone().to(dev).requires_grad_(True)
        E = IU.ion_interaction_sum(box, fr @ box, z.to(dev), Rc, Rd)
        gb, gf = torch.autograd.grad(E, (box, fr))
        out[dev] = (E.item(), gb.cpu(), gf.cpu())
    assert abs(out[DEV][0] - out['cpu'][0]) <= 1e-11 * abs(out['cpu'][0])
    for k in (1, 2):
        ref = out['cpu'][k]
        assert ((out[DEV][k] - ref).abs().max() / ref.abs().max()).item() < 1e-9, k


def test_system_energy_forces_stress_with_ion_ion(golden_dir, potentials_dir):
    """System with the IonIon term: forces and stress against the reference's autograd vectors (ions_*.npz)."""
    import profess_ad_b200.functionals as F
    from profess_ad_b200.system import System
    from test_oracle_ions import load_case
    g, box, den, species = load_case('alli_mixed', golden_dir, potentials_dir)
    ions = [[os.path.basename(p)[:2].capitalize(), p, f] for p, f in species]
    s = System(box, tuple(den.shape), ions, [F.IonIon, F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger], units='b',
               coord_type='fractional')
    s.set_density(torch.from_numpy(g['den']))
    assert np.abs(s.forces('Ha/b').cpu().numpy() - g['forces_Ha_b']).max() <= 1e-9 * np.abs(g['forces_Ha_b']).max()
    st = s.stress('Ha/b3').cpu().numpy()
    assert np.abs(st - g['stress_Ha_b3']).max() <= 1e-9 * np.abs(g['stress_Ha_b3']).max()
