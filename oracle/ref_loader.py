"""Import the unmodified reference from oracle/_ref (built by oracle/make_ref.py).  Test infrastructure only.

    ref = load_reference()        # None if oracle/_ref is absent
    ref.functionals, ref.functional_tools, ref.system, ref.ion_utils

The reference package is called ``professad`` -- the same name as this repository's drop-in alias -- so it is loaded
under the private name ``professad_reference`` with importlib (its intra-package imports are relative or go through
``professad.``; the latter are redirected while the reference modules execute)."""
import importlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')


def available():
    return os.path.isfile(os.path.join(REF, 'professad', 'functionals.py'))


def load_reference():
    if not available():
        return None
    if 'professad_reference' in sys.modules:
        return sys.modules['professad_reference']
    stubs = os.path.join(REF, '_stubs')
    for name in ('xitorch', 'torch_nl', 'matplotlib'):
        try:
            importlib.import_module(name)
        except Exception:      # noqa: BLE001 -- not installed: use the stand-in
            if stubs not in sys.path:
                sys.path.append(stubs)
    # the reference imports itself as `professad.<module>`: park this repo's alias (and anything already imported
    # under that name) while the reference loads, then restore it
    parked = {k: v for k, v in sys.modules.items() if k == 'professad' or k.startswith('professad.')}
    for k in parked:
        del sys.modules[k]
    sys.path.insert(0, REF)
    try:
        pkg = importlib.import_module('professad')
        mods = {}
        for sub in ('functional_tools', 'functionals', 'ion_utils', 'system'):
            mods[sub] = importlib.import_module('professad.' + sub)
        loaded = {k: v for k, v in sys.modules.items() if k == 'professad' or k.startswith('professad.')}
    finally:
        sys.path.remove(REF)
        for k in [k for k in sys.modules if k == 'professad' or k.startswith('professad.')]:
            del sys.modules[k]
        sys.modules.update(parked)
    ref = types.ModuleType('professad_reference')
    ref.package = pkg
    ref._modules = loaded        # keep them alive
    for sub, m in mods.items():
        setattr(ref, sub, m)
    sys.modules['professad_reference'] = ref
    return ref
