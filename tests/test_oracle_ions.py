"""The oracle's ionic rows (v_ext builder, ion-electron forces, stress by autograd) against vectors produced by the
unmodified reference (tests/golden/make_golden_ions.py)."""
import os

import numpy as np
import pytest
import torch

CASES = ['li2_odd', 'li2_even', 'alli_mixed']


def load_case(case, golden_dir, potentials_dir):
    g = np.load(os.path.join(golden_dir, f'ions_{case}.npz'))
    box = torch.from_numpy(g['box_bohr'])
    frac = torch.from_numpy(g['frac'])
    species, first = [], 0
    for pot, cnt in zip(g['pots'], g['counts']):
        species.append((os.path.join(potentials_dir, str(pot)), frac[first:first + int(cnt)]))
        first += int(cnt)
    return g, box, torch.from_numpy(g['den']), species


@pytest.mark.parametrize('case', CASES)
def test_oracle_vext_forces_stress_match_reference(case, golden_dir, potentials_dir):
    from oracle import ofdft_oracle as orc
    g, box, den, species = load_case(case, golden_dir, potentials_dir)
    cart = [(p, f @ box) for p, f in species]
    v = orc.ionic_potential(box, den.shape, cart)
    assert np.abs(v.numpy() - g['v_ext']).max() <= 1e-13 * np.abs(g['v_ext']).max()
    F = orc.ion_electron_forces(box, den, cart).numpy()
    assert np.abs(F - g['forces_IonElectron']).max() <= 1e-12 * np.abs(g['forces_IonElectron']).max()
    st = orc.ion_electron_stress(box, den, species).numpy()
    assert np.abs(st - g['stress_IonElectron']).max() <= 1e-12 * np.abs(g['stress_IonElectron']).max()
    for key in g.files:
        if key.startswith('stress_') and key[7:] not in ('Ha_b3', 'IonElectron', 'IonIon'):
            st = orc.stress(box, den, getattr(orc, key[7:])).numpy()
            assert np.abs(st - g[key]).max() <= 1e-12 * np.abs(g[key]).max(), key
