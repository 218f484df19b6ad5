"""Lattice vectors and fractional coordinates of simple crystals (src/professad/crystal_tools.py).
Host-side constructor of a (3,3) and an (N,3) tensor; no hot loop."""
import numpy as np
import torch

_FCC = [[0.0, 0.5, 0.5], [0.5, 0.0, 0.5], [0.5, 0.5, 0.0]]
_BCC = [[-0.5, 0.5, 0.5], [0.5, -0.5, 0.5], [0.5, 0.5, -0.5]]
_FCC_BASIS = [[0.0, 0.0, 0.0], [0.5, 0.5, 0.0], [0.5, 0.0, 0.5], [0.0, 0.5, 0.5]]


def _t(x):
    return torch.tensor(x, dtype=torch.double)


def _cubic(vol_per_atom, atoms_per_conventional_cell):
    return (atoms_per_conventional_cell * vol_per_atom) ** (1 / 3)


def _cell(crystal, vol_per_atom, c_over_a):
    if crystal == 'sc':
        return _cubic(vol_per_atom, 1) * torch.eye(3, dtype=torch.double), torch.zeros((1, 3), dtype=torch.double)
    if crystal == 'bcc':
        return _cubic(vol_per_atom, 2) * _t(_BCC), torch.zeros((1, 3), dtype=torch.double)
    if crystal == 'bcc-c':
        return _cubic(vol_per_atom, 2) * torch.eye(3, dtype=torch.double), _t([[0.0, 0.0, 0.0], [0.5, 0.5, 0.5]])
    if crystal == 'fcc':
        return _cubic(vol_per_atom, 4) * _t(_FCC), torch.zeros((1, 3), dtype=torch.double)
    if crystal == 'fcc-c':
        return _cubic(vol_per_atom, 4) * torch.eye(3, dtype=torch.double), _t(_FCC_BASIS)
    if crystal == 'dc':
        return _cubic(vol_per_atom, 8) * _t(_FCC), _t([[0.0, 0.0, 0.0], [0.25, 0.25, 0.25]])
    if crystal == 'dc-c':
        shifted = [[0.25, 0.25, 0.25], [0.25, 0.75, 0.75], [0.75, 0.75, 0.25], [0.75, 0.25, 0.75]]
        return _cubic(vol_per_atom, 8) * torch.eye(3, dtype=torch.double), _t(_FCC_BASIS + shifted)
    if crystal == 'hcp':
        a = ((2 * vol_per_atom) / (np.sqrt(3) / 2 * c_over_a)) ** (1 / 3)
        lat = a * _t([[1, 0, 0], [-0.5, np.sqrt(3) / 2, 0], [0, 0, c_over_a]])
        return lat, _t([[1 / 3, 2 / 3, 3 / 4], [2 / 3, 1 / 3, 1 / 4]])
    raise ValueError('\'crystal\' argument \'' + crystal + '\' not recognized')


def get_cell(crystal, vol_per_atom, c_over_a=np.sqrt(8 / 3), coord_type='fractional'):
    """Same contract as crystal_tools.py:11-59: 'sc', 'bcc'/'bcc-c', 'fcc'/'fcc-c', 'dc'/'dc-c', 'hcp'."""
    lattice_vectors, frac = _cell(crystal, vol_per_atom, c_over_a)
    if coord_type == 'fractional':
        return lattice_vectors, frac
    if coord_type == 'cartesian':
        return lattice_vectors, frac @ lattice_vectors
    raise ValueError('Only \'fractional\' or \'cartesian\' allowed for argument \'coord_type\'.')


# the reference's per-lattice constructors (crystal_tools.py:62-136), same signatures
def _cell_type(prim, conv, cell_type, vol_per_atom, c_over_a=None):
    if cell_type == 'primitive':
        return _cell(prim, vol_per_atom, c_over_a)
    if cell_type == 'conventional':
        return _cell(conv, vol_per_atom, c_over_a)
    raise ValueError('Only \'primitive\' or \'conventional\' allowed for argument \'cell_type\'.')


def simple_cubic(vol_per_atom):
    return _cell('sc', vol_per_atom, None)


def body_centered_cubic(vol_per_atom, cell_type='conventional'):
    return _cell_type('bcc', 'bcc-c', cell_type, vol_per_atom)


def face_centered_cubic(vol_per_atom, cell_type='primitive'):
    return _cell_type('fcc', 'fcc-c', cell_type, vol_per_atom)


def diamond_cubic(vol_per_atom, cell_type='conventional'):
    return _cell_type('dc', 'dc-c', cell_type, vol_per_atom)


def hexagonal_close_packed(vol_per_atom, c_over_a=1.633):
    return _cell('hcp', vol_per_atom, c_over_a)
