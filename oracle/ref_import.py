"""TEST INFRASTRUCTURE — import the UNMODIFIED reference from /root/reference on top of the timm
shim and the SoccerNet/matplotlib stubs.  Only available in the build container (the GPU box has
no /root/reference); used by oracle/gen_golden.py and by tests marked `needs_reference`.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get('TDEED_REFERENCE_ROOT', '/root/reference')
_HERE = os.path.dirname(os.path.abspath(__file__))


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'model'))


class reference_modules:
    """Context manager: temporarily makes `model.*` / `util.*` resolve to the reference.

    The product ships same-named drop-in packages (t-deed_b200/model, t-deed_b200/util), so any
    already-imported `model*`/`util*`/`timm*` modules are stashed and restored on exit.
    """

    _PREFIXES = ('model', 'util', 'dataset', 'timm', 'SoccerNet', 'matplotlib', 'wandb')

    def __enter__(self):
        if not reference_available():
            raise RuntimeError('reference not present at ' + REFERENCE_ROOT)
        self._saved = {k: v for k, v in sys.modules.items()
                       if k.split('.')[0] in self._PREFIXES}
        for k in self._saved:
            del sys.modules[k]
        self._path = list(sys.path)
        # the reference's model/ and util/ have no __init__.py (namespace packages): any regular package of
        # the same name further down sys.path (the product's drop-in t-deed_b200/model) would win -> hide it
        sys.path[:] = [p for p in sys.path
                       if not os.path.exists(os.path.join(p or '.', 'model', '__init__.py'))]
        sys.path[:0] = [REFERENCE_ROOT, os.path.join(_HERE, 'timm_shim'), os.path.join(_HERE, 'stubs')]
        mods = {}
        for name in ('model.modules', 'model.shift', 'model.model', 'model.impl.gsm',
                     'model.impl.gsf', 'util.eval'):
            mods[name] = importlib.import_module(name)
        self.mods = mods
        return mods

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k.split('.')[0] in self._PREFIXES]:
            del sys.modules[k]
        sys.modules.update(self._saved)
        sys.path[:] = self._path
        return False
