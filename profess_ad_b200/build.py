"""Build the C-ABI shared library (sm_100a only) in-tree with nvcc.

    python -m profess_ad_b200.build          # or: from profess_ad_b200.build import build; build()

The result, profess_ad_b200/libprofessad_b200.so, is git-ignored but travels to the GPU box.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libprofessad_b200.so')

NVCC_FLAGS = [
    '-O3', '-std=c++17', '--extended-lambda', '-lineinfo',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-O3',
    '-cudart', 'shared',
]


def _nvcc():
    exe = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(exe):
        raise RuntimeError('nvcc not found: the B200 library cannot be built (there is no CPU fallback)')
    return exe


def sources():
    return sorted(glob.glob(os.path.join(CSRC, '*.cu')))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, '*.cuh')) + glob.glob(os.path.join(HERE, '..', 'include', '*.h'))
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ into one shared library.  Objects are built in parallel."""
    if not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objdir = os.path.join(HERE, 'build')
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        objs.append(obj)
        cmd = [nvcc, '-c', src, '-o', obj] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else [])
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    failed = False
    for cmd, pr in procs:
        out, _ = pr.communicate()
        if pr.returncode != 0:
            failed = True
            sys.stderr.write(' '.join(cmd) + '\n' + out + '\n')
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc failed')
    tmp = LIB + '.tmp'
    link = [nvcc, '-shared', '-o', tmp] + objs + ['-cudart', 'shared', '-lcufft',
                                                   '-gencode', 'arch=compute_100a,code=sm_100a']
    r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(' '.join(link) + '\n' + r.stdout + '\n')
        raise RuntimeError('link failed')
    os.replace(tmp, LIB)        # atomic: a snapshot of the tree (gpurun) never sees a half-written library
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
