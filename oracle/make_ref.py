"""Recipe for oracle/_ref/: the UNMODIFIED reference (PROFESS-AD, pure Python) made importable on a box that has
neither /root/reference nor its three uninstalled dependencies.

    python oracle/make_ref.py            # build container only; writes ONLY into oracle/_ref/ (git-ignored, travels
                                         # to the GPU box with gpurun like the built .so)

  oracle/_ref/professad/...   byte-for-byte copy of /root/reference/src/professad (nothing is edited; `diff -r` clean)
  oracle/_ref/_stubs/...      stand-ins for xitorch, torch_nl and matplotlib (not installed in this image):
                              xitorch.solve_ivp / minimize raise (the Huang-Carter kernel table is the only user),
                              torch_nl.compute_neighborlist is a brute-force periodic pair list with the same
                              (mapping, batch, shifts) convention -- used by the ion-ion term only
  oracle/_ref/MANIFEST.json   sha256 of every copied file, source path, date

Test infrastructure: only tests/, __graft_entry__ and bench.py's CPU legs load it (oracle/ref_loader.py)."""
import hashlib
import json
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = '/root/reference/src/professad'
DST = os.path.join(HERE, '_ref')

STUB_XITORCH = '''"""stub: xitorch is not installed in this image"""
'''
STUB_XITORCH_SUB = '''def _no(*a, **k):
    raise NotImplementedError('xitorch is not installed (stub in oracle/_ref/_stubs)')


solve_ivp = _no
minimize = _no
'''
STUB_TORCH_NL = '''"""stub for torch_nl: brute-force periodic pair list, same return convention as compute_neighborlist
(r_ij = pos[j] + shifts @ cell - pos[i])."""
import numpy as np
import torch


def compute_neighborlist(cutoff, pos, cell, pbc, batch, self_interaction=False):
    cutoff = float(cutoff)
    inv = torch.linalg.inv(cell)
    heights = 1.0 / torch.sqrt(torch.sum(inv.T.pow(2), 1))
    reps = [int(np.ceil(cutoff / h.item())) + 1 for h in heights]
    rng = [torch.arange(-r, r + 1, dtype=torch.double) for r in reps]
    S = torch.stack(torch.meshgrid(*rng, indexing='ij'), -1).reshape(-1, 3)
    n = pos.shape[0]
    ii, jj = torch.meshgrid(torch.arange(n), torch.arange(n), indexing='ij')
    ii, jj = ii.reshape(-1), jj.reshape(-1)
    disp = (pos[jj] - pos[ii]).unsqueeze(1) + (S @ cell).unsqueeze(0)
    dist = disp.norm(dim=2)
    ok = dist < cutoff
    if not self_interaction:
        ok &= ~((S.abs().sum(1) == 0).unsqueeze(0) & (ii == jj).unsqueeze(1))
    pair, sh = torch.nonzero(ok, as_tuple=True)
    mapping = torch.stack([ii[pair], jj[pair]])
    return mapping, torch.zeros(mapping.shape[1], dtype=torch.long), S[sh]
'''


def build(verbose=True):
    if not os.path.isdir(SRC):
        if verbose:
            print('oracle/make_ref.py: %s not present (GPU box?): keeping whatever oracle/_ref holds' % SRC)
        return os.path.isdir(os.path.join(DST, 'professad'))
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    shutil.copytree(SRC, os.path.join(DST, 'professad'), ignore=shutil.ignore_patterns('__pycache__', '*.pyc'))
    stubs = os.path.join(DST, '_stubs')
    for pkg, files in (('xitorch', {'__init__.py': STUB_XITORCH, 'integrate.py': STUB_XITORCH_SUB, 'optimize.py': STUB_XITORCH_SUB}),
                       ('torch_nl', {'__init__.py': STUB_TORCH_NL}),
                       ('matplotlib', {'__init__.py': '"""stub"""\n', 'pyplot.py': '"""stub"""\n'})):
        os.makedirs(os.path.join(stubs, pkg), exist_ok=True)
        for name, text in files.items():
            with open(os.path.join(stubs, pkg, name), 'w') as f:
                f.write(text)
    manifest = {'source': SRC, 'when': time.strftime('%Y-%m-%dT%H:%M:%SZ', time.gmtime()), 'files': {}}
    for root, _, names in os.walk(os.path.join(DST, 'professad')):
        for n in sorted(names):
            path = os.path.join(root, n)
            with open(path, 'rb') as f:
                manifest['files'][os.path.relpath(path, DST)] = hashlib.sha256(f.read()).hexdigest()
    with open(os.path.join(DST, 'MANIFEST.json'), 'w') as f:
        json.dump(manifest, f, indent=1)
    if verbose:
        print('oracle/_ref: %d reference files copied unmodified from %s' % (len(manifest['files']), SRC))
    return True


if __name__ == '__main__':
    sys.exit(0 if build() else 1)
