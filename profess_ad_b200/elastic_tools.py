"""Equation-of-state fit for strain scans (src/professad/elastic_tools.py:16-77).  Scalar
post-processing of <= ~11 (V, E) pairs on the host; no hot loop."""
import numpy as np
import torch
from scipy.optimize import curve_fit

m_per_bohr = 5.29177210903e-11
A_per_b = m_per_bohr * 1e10
J_per_Ha = 4.3597447222071e-18
eV_per_Ha = J_per_Ha / 1.602176634e-19
GPa_per_atomic = J_per_Ha / m_per_bohr**3 * 1e-9
GPa_per_Ab3 = GPa_per_atomic / (eV_per_Ha / A_per_b**3)


def murnaghan(v, K0, K0p, E0, V0):
    return E0 + (K0 * v / K0p) * ((V0 / v)**K0p / (K0p - 1) + 1) - K0 * V0 / (K0p - 1)


def birch_murnaghan(v, K0, K0p, E0, V0):
    x = (V0 / v)**(2 / 3) - 1
    return E0 + 9 * V0 * K0 / 16 * (K0p * x**3 + x**2 * (6 - 4 * (V0 / v)**(2 / 3)))


def fit_eos(vol, ene, eos='bm', plot=False):
    """Fit E(V) to the Murnaghan ('m') or Birch-Murnaghan ('bm') equation of state.  Returns
    (K0, K0', E0, V0) and their standard errors, in the units of the inputs.  The starting guess
    is the harmonic solid E = E0 + K0 (V - V0)^2 / (2 V0) from a quadratic fit, K0' = 3.5."""
    if eos not in ('m', 'bm'):
        raise ValueError('Only \'m\' or \'bm\' recognized for \'eos\' argument.')
    if plot:
        raise NotImplementedError('plotting needs matplotlib, which is not part of this build')
    vol, ene = np.asarray(vol, dtype=float), np.asarray(ene, dtype=float)
    a2, a1, a0 = np.polyfit(vol, ene, 2)
    K0 = -a1
    V0 = K0 / (2 * a2)
    E0 = a0 - 0.5 * K0 * V0
    model = murnaghan if eos == 'm' else birch_murnaghan
    params, pcov = curve_fit(model, vol, ene, p0=(K0, 3.5, E0, V0), maxfev=1000)
    return params, np.sqrt(np.diag(pcov))


# ---- polycrystalline averages of a 6 x 6 elastic-constant matrix (elastic_tools.py:80-176) ----------------
def voigt_moduli(C):
    """Voigt (uniform strain) bulk and shear moduli: 9 K = tr(C_nn) + 2 (C12 + C23 + C31),
    15 G = tr(C_nn) - (C12 + C23 + C31) + 3 (C44 + C55 + C66)."""
    normal = C[0, 0] + C[1, 1] + C[2, 2]
    cross = C[0, 1] + C[1, 2] + C[0, 2]
    shear = C[3, 3] + C[4, 4] + C[5, 5]
    return (normal + 2 * cross) / 9, (normal - cross + 3 * shear) / 15


def reuss_moduli(C):
    """Reuss (uniform stress) bulk and shear moduli from the compliance S = C^-1:
    1 / K = tr(S_nn) + 2 (S12 + S23 + S31), 15 / G = 4 tr(S_nn) - 4 (S12 + S23 + S31) + 3 (S44 + S55 + S66)."""
    S = torch.linalg.inv(C) if isinstance(C, torch.Tensor) else np.linalg.inv(C)
    normal = S[0, 0] + S[1, 1] + S[2, 2]
    cross = S[0, 1] + S[1, 2] + S[0, 2]
    shear = S[3, 3] + S[4, 4] + S[5, 5]
    return 1 / (normal + 2 * cross), 15 / (4 * normal - 4 * cross + 3 * shear)


def shear_average(C, mean_type='arithmetic'):
    """Arithmetic (Hill) or geometric mean of the Voigt and Reuss shear moduli."""
    _, gv = voigt_moduli(C)
    _, gr = reuss_moduli(C)
    if mean_type == 'arithmetic':
        return 0.5 * (gv + gr)
    if mean_type == 'geometric':
        return (gv * gr) ** 0.5
    raise ValueError('Only \'arithmetic\' or \'geometric\' recognized for \'mean_type\' argument')


def poissons_ratio(K, G):
    """nu = (1 - 3 G / (3 K + G)) / 2"""
    return 0.5 * (1 - 3 * G / (3 * K + G))


def youngs_modulus(K, G):
    """E = 1 / (1 / (3 G) + 1 / (9 K))"""
    return 1 / (1 / (3 * G) + 1 / (9 * K))
