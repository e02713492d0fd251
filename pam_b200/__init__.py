"""Short import alias for the package directory
``part-aware_measurement_for_3d_pose_estimation_and_tracking_b200`` (whose name, fixed by the
build contract, contains a hyphen and therefore cannot appear in an ``import`` statement).

Both spellings work and give the SAME module objects: ``from pam_b200 import tracker`` and
``import pam_b200.tracker`` (a meta-path finder maps every ``pam_b200.x`` to the real package's ``x``, so that e.g.
``pam_b200.tracker.PamError`` is the class the drop-in modules raise)."""
import importlib as _il
import importlib.abc as _abc
import importlib.util as _util
import sys as _sys

_REAL = "part-aware_measurement_for_3d_pose_estimation_and_tracking_b200"
_real = _il.import_module(_REAL)


class _AliasLoader(_abc.Loader):
    def __init__(self, real_name):
        self.real_name = real_name

    def create_module(self, spec):
        return _il.import_module(self.real_name)          # the real module object, not a copy

    def exec_module(self, module):
        pass


class _AliasFinder(_abc.MetaPathFinder):
    def find_spec(self, fullname, path=None, target=None):
        if fullname == __name__ or not fullname.startswith(__name__ + "."):
            return None
        real_name = _REAL + fullname[len(__name__):]
        if _util.find_spec(real_name) is None:
            return None
        return _util.spec_from_loader(fullname, _AliasLoader(real_name))


if not any(isinstance(f, _AliasFinder) for f in _sys.meta_path):
    _sys.meta_path.insert(0, _AliasFinder())
_sys.modules[__name__] = _real
