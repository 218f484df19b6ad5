"""Geometry optimisations by the UNMODIFIED reference (System.optimize_geometry, system.py:937-1068) on small cells.
Run in the build container only:  python tests/golden/make_golden_geometry.py
Outputs geometry_<case>.npz: start geometry, final lattice / fractional coordinates / energy / max force / max stress."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, POT      # noqa: E402


def main():
    F, T, S, C = import_reference()
    torch.set_num_threads(8)
    System = S.System
    # the cell of tests/test_forces.py:14-20 (2 Li atoms, skewed), coarse grid
    box_a = torch.tensor([[3.54, -0.13, 0.25], [-0.33, 3.82, 0.24], [0.55, 0.04, 3.45]], dtype=torch.double)
    frac = torch.tensor([[0, 0, 0], [0.35, 0.65, 0.45]], dtype=torch.double)
    terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    for name, kw in (('li2_ions', dict(ftol=0.02, stol=None)), ('li2_full', dict(ftol=0.02, stol=0.002))):
        shape = System.ecut2shape(500, box_a)
        s = System(box_a, shape, [['Li', os.path.join(POT, 'li.gga.recpot'), frac]], terms, units='a', coord_type='fractional')
        ok = s.optimize_geometry(g_maxiter=60, g_verbose=True, ntol=1e-9, **kw)
        out = dict(box0_A=box_a.numpy(), frac0=frac.numpy(), shape=np.array(shape), converged=ok,
                   box_bohr=s.lattice_vectors('b').detach().numpy(), frac=s.fractional_ionic_coordinates().detach().numpy(),
                   energy_eV=s.energy('eV'), max_force_eV_A=float(s.forces('eV/a').abs().max()),
                   max_stress_eV_A3=float(s.stress('eV/a3').abs().max()), volume_A3=s.volume('a3'))
        np.savez_compressed(os.path.join(HERE, f'geometry_{name}.npz'), **out)
        print(name, ok, out['energy_eV'], out['max_force_eV_A'], out['max_stress_eV_A3'], out['volume_A3'], flush=True)


if __name__ == '__main__':
    main()
