"""The analytic, minimum-FFT formulation implemented by the CUDA kernels (tests/analytic_model.py)
equals the autograd oracle on CPU -- including on even grids with skewed cells, where the
reference's positive-Nyquist convention makes the multipliers non-Hermitian."""
import math

import pytest
import torch

import analytic_model as am
from oracle import ofdft_oracle as orc

A98, B98 = (5 + math.sqrt(5)) / 6, (5 - math.sqrt(5)) / 6


@pytest.mark.parametrize('shape', [(12, 10, 14), (9, 11, 7), (8, 9, 6), (6, 6, 6)])
def test_analytic_potentials_equal_autograd(shape):
    box, den = orc.synth_rough(shape, seed=5)
    pairs = {
        'Hartree': (orc.Hartree, lambda: am.hartree(box, den)),
        'ThomasFermi': (orc.ThomasFermi, lambda: am.thomas_fermi(box, den)),
        'Weizsaecker': (orc.Weizsaecker, lambda: am.weizsaecker(box, den)),
        'WangTeter': (orc.WangTeter, lambda: am.wt_family(box, den, 5 / 6, 5 / 6)),
        'WGC98': (orc.WangGovindCarter98, lambda: am.wt_family(box, den, A98, B98)),
        'WGC99': (orc.WangGovindCarter99(),
                  lambda: am.wgc99(box, den, lambda eta: orc.wgc99_kernel(eta, A98, B98, 2.7), A98, B98, 2.7, 1.0)),
        'PerdewZunger': (orc.PerdewZunger, lambda: am.perdew_zunger(box, den)),
        'PBE': (orc.PerdewBurkeErnzerhof, lambda: am.pbe(box, den)),
    }
    for name, (fo, fm) in pairs.items():
        E, V = orc.energy_and_potential(box, den, fo)
        e, v = fm()
        assert abs(E.item() - e) < 1e-12 * max(1.0, abs(e)), name
        assert ((V - v).abs().max() / V.abs().max()).item() < 1e-12, name


def test_chi_projection_equals_autograd():
    """system.py:842-854 vs autograd through n = N chi^2 / int chi^2 (tests/test_den_opt.py:58-75)."""
    box, den = orc.synth_rough((9, 8, 10), seed=2)
    chi = torch.sqrt(den) * 1.3
    n_elec = 7.0
    terms = [orc.Hartree, orc.WangTeter, orc.PerdewZunger]
    E, g_auto, n = orc.chi_gradient(box, chi, n_elec, terms)
    v = sum(orc.energy_and_potential(box, n, f)[1] for f in terms)
    g_proj = am.chi_projection(box, chi, n_elec, v)
    assert ((g_auto - g_proj).abs().max() / g_auto.abs().max()).item() < 1e-12
    g_orc = orc.chi_gradient_from_potential(box, chi, n_elec, v)
    assert ((g_auto - g_orc).abs().max() / g_auto.abs().max()).item() < 1e-12


@pytest.mark.parametrize('shape,seed', [((8, 10, 12), 1), ((9, 7, 11), 2), ((8, 9, 10), 3)])
def test_analytic_stresses_equal_autograd(shape, seed):
    """The stress formulas csrc/stress.cu and csrc/ions.cu implement, against autograd through box_vecs
    (functional_tools.py:73-100) on even, odd and mixed skewed grids."""
    import numpy as np
    import os
    box, den = orc.synth_rough(shape, seed=seed)

    def rel(a, b):
        return float((a - b).abs().max() / b.abs().max())
    E, v = am.thomas_fermi(box, den)
    assert rel(am.stress_local(box, den, E, v), orc.stress(box, den, orc.ThomasFermi)) < 1e-12
    E, v = am.perdew_zunger(box, den)
    assert rel(am.stress_local(box, den, E, v), orc.stress(box, den, orc.PerdewZunger)) < 1e-12
    assert rel(am.stress_hartree(box, den), orc.stress(box, den, orc.Hartree)) < 1e-12
    assert rel(am.stress_weizsaecker(box, den), orc.stress(box, den, orc.Weizsaecker)) < 1e-12
    for a, b in ((5 / 6, 5 / 6), ((5 + 5 ** 0.5) / 6, (5 - 5 ** 0.5) / 6)):
        ref = orc.stress(box, den, lambda bx, d: orc.nonlocal_wt_term(bx, d, a, b))
        assert rel(am.stress_wt_nonlocal(box, den, a, b), ref) < 1e-12
    assert rel(am.stress_pbe(box, den), orc.stress(box, den, orc.PerdewBurkeErnzerhof)) < 1e-12
    frac = torch.rand(3, 3, dtype=torch.double, generator=torch.Generator().manual_seed(seed))
    path = os.path.join(os.path.dirname(__file__), 'potentials', 'al.gga.recpot')
    ks, pot, z = orc.read_recpot(path)
    smooth = pot.copy()
    smooth[1:] += 4 * np.pi * z / ks[1:] ** 2
    assert rel(am.stress_ion_electron(box, den, [(ks, smooth, z, frac)]), orc.ion_electron_stress(box, den, [(path, frac)])) < 1e-11


@pytest.mark.parametrize('shape,seed', [((8, 10, 12), 1), ((9, 7, 11), 2)])
def test_huang_carter_stress_formula_equals_autograd(shape, seed, golden_dir):
    """The analytic Huang-Carter stress (volume / density-scaling term from the potential, strained spectral gradient,
    eta-derivative of the kernel table through the node weights) against autograd through box_vecs -- the formula for
    the kernels that do not exist yet (pad_stress_terms raises for kinetic == 3)."""
    import os
    import numpy as np
    tab = np.load(os.path.join(golden_dir, 'hc_table.npz'))
    box, den = orc.synth_rough(shape, seed=seed)
    for f in (orc.RevisedHuangCarter(0.45, 0.10, 2 / 3, 1.15, kernel=torch.from_numpy(tab['revhc'])),
              orc.HuangCarter(0.01177, 0.7143, 1.2, kernel=torch.from_numpy(tab['hc']))):
        ref = orc.stress(box, den, f)
        E, v = am.thomas_fermi(box, den)
        got = am.stress_huang_carter_nonlocal(box, den, f) + am.stress_local(box, den, E, v) + am.stress_weizsaecker(box, den)
        assert float((got - ref).abs().max() / ref.abs().max()) < 1e-12
