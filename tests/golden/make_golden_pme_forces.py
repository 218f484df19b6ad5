"""Golden fixtures for the particle-mesh Ewald forces / stress of the IonElectron term from the UNMODIFIED reference
(autograd through structure_factor_spline, system.py:913-935 with pme_order set).  Build container only:

    python tests/golden/make_golden_pme_forces.py

Reads the densities and geometries of ions_<case>.npz (make_golden_ions.py); writes ions_pme.npz with
<case>_o<order>_{forces,stress,vext,energy} for the IonElectron term alone."""
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
    out = {}
    for case in ('li2_odd', 'li2_even', 'alli_mixed'):
        g = np.load(os.path.join(HERE, f'ions_{case}.npz'))
        box = torch.from_numpy(g['box_bohr'])
        den = torch.from_numpy(g['den'])
        frac = torch.from_numpy(g['frac'])
        ions, first = [], 0
        for pot, cnt in zip(g['pots'], g['counts']):
            pot = str(pot)
            ions.append([pot[:2].capitalize(), os.path.join(POT, pot), frac[first:first + int(cnt)].clone()])
            first += int(cnt)
        for order in (4, 8):
            s = S.System(box.clone(), tuple(den.shape), ions, [F.IonElectron], units='b', coord_type='fractional', pme_order=order)
            s.set_density(den.clone())
            key = f'{case}_o{order}_'
            out[key + 'forces'] = s.forces('Ha/b').detach().numpy()
            out[key + 'stress'] = s.stress('Ha/b3').detach().numpy()
            out[key + 'vext'] = s.ionic_potential().detach().numpy()
            out[key + 'energy'] = s.energy('Ha')
            print(key, 'E', out[key + 'energy'], 'max|F|', np.abs(out[key + 'forces']).max(), 'diff to exact',
                  np.abs(out[key + 'forces'] - g['forces_IonElectron']).max())
    np.savez_compressed(os.path.join(HERE, 'ions_pme.npz'), **out)


if __name__ == '__main__':
    main()
