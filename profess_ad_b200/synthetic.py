"""Deterministic synthetic inputs for benchmarks.

No files, no reference code: a smooth Al-like density on a cubic supercell.  bench.py uses this for
the B200 arm; the CPU baseline arm and the tests generate their inputs independently.
"""
import math

import torch


def smooth_supercell(n, side, device='cpu'):
    """(box_vecs, den): cubic cell of side*a (a = 4.05 A in bohr), 4*side^3 Al atoms, 3 e-/atom, n^3 grid.

    den = n0 (1 + 0.3 cos(kX) cos(kY) cos(kZ) + 0.05 cos(2kX) cos(2kY)), k = 2 pi side, x = i / n.
    """
    dt = torch.double
    a = 4.05 / 0.529177210903
    L = side * a
    box = L * torch.eye(3, dtype=dt, device=device)
    x = torch.arange(n, dtype=dt, device=device) / n
    k = 2 * math.pi * side
    c1, c2 = torch.cos(k * x), torch.cos(2 * k * x)
    n0 = 12 * side ** 3 / L ** 3
    den = n0 * (1 + 0.3 * c1[:, None, None] * c1[None, :, None] * c1[None, None, :]
                + 0.05 * (c2[:, None, None] * c2[None, :, None]).expand(n, n, n))
    return box, den.contiguous()
