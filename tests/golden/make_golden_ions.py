"""Golden fixtures for the ionic rows (v_ext builder, forces, stress) from the UNMODIFIED
reference.  Run in the build container only:

    python tests/golden/make_golden_ions.py

Same import stubs as make_golden.py.  Outputs (fp64):
  ions_<case>.npz : box_bohr, shape, frac (all ions, species order), counts, den, v_ext, forces_Ha_b, stress_Ha_b3,
                    E_ion_Ha and the per-functional stresses used by tests/test_gpu_ions.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import import_reference, POT      # noqa: E402


def stress_of(S, T, F, system_cls, box, den, functional):
    return T.get_stress(box, den, functional).detach().numpy()


def main():
    F, T, S, C = import_reference()
    torch.set_num_threads(8)
    System = S.System
    A = System.A_per_b

    cases = {}
    # the cell of the reference's tests/test_forces.py:14-20, odd grid from ecut2shape and a forced even grid
    box_a = torch.tensor([[3.54, -0.13, 0.25], [-0.33, 3.82, 0.24], [0.55, 0.04, 3.45]], dtype=torch.double)
    frac = torch.tensor([[0, 0, 0], [0.35, 0.65, 0.45]], dtype=torch.double)
    terms = [F.IonIon, F.IonElectron, F.Hartree, F.WangTeter, F.PerdewBurkeErnzerhof]
    for name, shape in (('li2_odd', System.ecut2shape(600, box_a)), ('li2_even', (12, 14, 12))):
        s = System(box_a, shape, [['Li', os.path.join(POT, 'li.gga.recpot'), frac]], terms, units='a', coord_type='fractional')
        s.optimize_density(ntol=1e-8)
        cases[name] = (s, [('li.gga.recpot', 2)], terms)

    # two species, four ions, generic positions, mixed-parity grid; TF + vW + Hartree + PZ
    gen = torch.Generator().manual_seed(7)
    box_b = 7.2 * torch.eye(3, dtype=torch.double) + 0.3 * torch.rand(3, 3, dtype=torch.double, generator=gen)
    f_al = torch.rand(2, 3, dtype=torch.double, generator=gen)
    f_li = torch.rand(2, 3, dtype=torch.double, generator=gen)
    terms_b = [F.IonIon, F.IonElectron, F.Hartree, F.ThomasFermi, F.Weizsaecker, F.PerdewZunger]
    s = System(box_b, (16, 15, 18), [['Al', os.path.join(POT, 'al.gga.recpot'), f_al], ['Li', os.path.join(POT, 'li.gga.recpot'), f_li]],
               terms_b, units='b', coord_type='fractional')
    s.optimize_density(ntol=1e-8)
    cases['alli_mixed'] = (s, [('al.gga.recpot', 2), ('li.gga.recpot', 2)], terms_b)

    for name, (s, species, tms) in cases.items():
        den = s.density().clone()
        box = s.lattice_vectors('b').clone()
        out = dict(box_bohr=box.numpy(), shape=np.array(den.shape), frac=s.fractional_ionic_coordinates().numpy(),
                   counts=np.array([c for _, c in species]), pots=np.array([p for p, _ in species]),
                   den=den.numpy(), v_ext=s.ionic_potential().numpy(), energy_Ha=s.energy('Ha'),
                   forces_Ha_b=s.forces('Ha/b').detach().numpy(), stress_Ha_b3=s.stress('Ha/b3').detach().numpy(),
                   pressure_Ha_b3=s.pressure('Ha/b3'))
        # the ion-electron and ion-ion parts on their own: Systems holding a single term, same density
        ions_arg = []
        first = 0
        for pot, cnt in species:
            ions_arg.append([pot[:2].capitalize(), os.path.join(POT, pot), s.fractional_ionic_coordinates()[first:first + cnt].clone()])
            first += cnt
        for term in (F.IonElectron, F.IonIon):
            part = System(box, tuple(den.shape), ions_arg, [term], units='b', coord_type='fractional')
            part.set_density(den.clone())
            out['forces_' + term.__name__] = part.forces('Ha/b').detach().numpy()
            out['stress_' + term.__name__] = part.stress('Ha/b3').detach().numpy()
        for f in tms:
            if f.__name__ in ('IonIon', 'IonElectron'):
                continue
            out['stress_' + f.__name__] = T.get_stress(box, den, f).detach().numpy()
        # WangGovindCarter99 with a FRESH functional object per call: the kernel is then generated inside the autograd
        # graph and the stress contains the kernel's own dependence on the cell (a kernel cached by an earlier call at the
        # same cell would be reused without its graph, functionals.py:961-966)
        A98, B98 = (5 + 5 ** 0.5) / 6, (5 - 5 ** 0.5) / 6
        out['stressfresh_WGC99'] = T.get_stress(box, den, F.WangGovindCarter99().forward).detach().numpy()
        out['stressfresh_WGC99_gamma05'] = T.get_stress(box, den, F.WangGovindCarter99((A98, B98, 0.5, 1.0)).forward).detach().numpy()
        out['stressfresh_WGC99_kappa12'] = T.get_stress(box, den, F.WangGovindCarter99((A98, B98, 2.7, 1.2)).forward).detach().numpy()
        np.savez_compressed(os.path.join(HERE, f'ions_{name}.npz'), **out)
        print(name, tuple(den.shape), 'E', out['energy_Ha'], 'max|F|', np.abs(out['forces_Ha_b']).max(),
              'stress diag', np.diag(out['stress_Ha_b3']))


if __name__ == '__main__':
    main()
