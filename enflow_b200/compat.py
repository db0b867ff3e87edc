"""Make `import enflow...` resolve to enflow_b200 (drop-in for code written against the reference)."""
import importlib
import sys

_SUBMODULES = ['nn', 'nn.egcl', 'nn.argmax', 'nn.floor', 'flow', 'flow.base', 'flow.dynamics', 'flow.loss',
               'data', 'data.base', 'data.synthetic', 'utils', 'utils.helpers', 'utils.conversion', 'main']


def install():
    import enflow_b200
    sys.modules.setdefault('enflow', enflow_b200)
    for sub in _SUBMODULES:
        try:
            sys.modules.setdefault('enflow.' + sub, importlib.import_module('enflow_b200.' + sub))
        except ModuleNotFoundError:
            pass
