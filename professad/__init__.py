"""Drop-in alias: ``import professad`` resolves to the B200-native implementation, so scripts written
against PROFESS-AD (``from professad.system import System``, ``from professad.functionals import ...``)
run unchanged."""
import importlib
import sys

import profess_ad_b200 as _impl

__version__ = _impl.__version__

for _name in ('functionals', 'functional_tools', 'system', 'ion_utils', 'crystal_tools', 'elastic_tools',
              '_optimizers', '_optimizers.lbfgs', '_optimizers.lbfgs.lbfgsnew', '_optimizers.tpgd',
              '_optimizers.tpgd.two_point_gradient_descent'):
    try:
        _mod = importlib.import_module('profess_ad_b200.' + _name)
    except ModuleNotFoundError:
        continue
    sys.modules['professad.' + _name] = _mod
    if '.' not in _name:
        globals()[_name] = _mod
