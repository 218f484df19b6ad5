"""Golden stresses of the Huang-Carter family from the UNMODIFIED reference's autograd (functional_tools.get_stress,
functional_tools.py:73-100), for the densities of the ions_<case>.npz fixtures.  Build container only:

    python tests/golden/make_golden_hc_stress.py

The omega(eta) table is the one of hc_table.npz, injected into the reference (``hc.kernel`` is a plain attribute,
functionals.py:1230) -- pinned given the table, like the HC energies / potentials.  Output: hc_stress.npz with
<case>_<HC|revHC> (3 x 3, Ha/bohr^3), the energies, and the parameters used."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference      # noqa: E402


def hc_with_table(cls, args, tab):
    """construct without running the (xitorch) kernel integration, then inject the table (as make_golden.py does)"""
    orig = cls.generate_kernel
    cls.generate_kernel = lambda self, *a, **k: None
    obj = cls(args)
    cls.generate_kernel = orig
    obj.kernel = tab.clone()
    obj.debug = False
    return obj


def main():
    F, T, S, C = import_reference()
    torch.set_num_threads(8)
    tab = np.load(os.path.join(HERE, 'hc_table.npz'))
    out = {}
    for case in ('li2_odd', 'li2_even', 'alli_mixed'):
        g = np.load(os.path.join(HERE, f'ions_{case}.npz'))
        box, den = torch.from_numpy(g['box_bohr']), torch.from_numpy(g['den'])
        for name, f in (('HC', hc_with_table(F.HuangCarter, (0.01177, 0.7143, 1.2), torch.from_numpy(tab['hc']))),
                        ('revHC', hc_with_table(F.RevisedHuangCarter, (0.45, 0.10, 2.0 / 3.0, 1.15), torch.from_numpy(tab['revhc'])))):
            st = T.get_stress(box, den, f.forward).detach().numpy()
            E = float(f.forward(box, den))
            out[f'{case}_{name}'] = st
            out[f'{case}_{name}_E'] = np.float64(E)
            print(case, name, 'E =', E, 'stress diag', np.diag(st))
    out['hc_args'] = np.array([0.01177, 0.7143, 1.2])
    out['revhc_args'] = np.array([0.45, 0.10, 2.0 / 3.0, 1.15])
    np.savez_compressed(os.path.join(HERE, 'hc_stress.npz'), **out)


if __name__ == '__main__':
    main()
