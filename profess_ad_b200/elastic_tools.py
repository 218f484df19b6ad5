"""Equation-of-state fit for strain scans (src/professad/elastic_tools.py:16-77).  Scalar
post-processing of <= ~11 (V, E) pairs on the host; no hot loop."""
import numpy as np
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
