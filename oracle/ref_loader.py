"""TEST INFRASTRUCTURE ONLY -- loader for the UNMODIFIED reference modules.

Imports the reference's own Python files from ``/root/reference/src`` (read-only,
present in the build container only -- it does NOT exist on the GPU box) so that

  * ``oracle/generic.py`` (the J-parametrised CPU restatement) can be proven
    bit-identical to the unmodified reference at J = 17, and
  * ``tests/golden/make_golden.py`` can generate the committed golden vectors.

Nothing in the product package, ``bench.py`` or the ``-m gpu`` tests may import this
file.  No reference source is copied: the modules are executed where they lie.

Shims applied before import (none touches a reference file):
  * ``np.float`` / ``np.int`` aliases (removed from numpy >= 1.24; used at
    src/utils/calculate.py:27, src/utils/matching.py:246,
    src/tracking/IterativeTracker.py:353, src/tracking/hypothesis.py:35),
  * stub modules ``matplotlib``, ``matplotlib.pyplot`` (hypothesis.py:4), ``cvxopt``
    (binary_integer_programming.py:5), ``easydict`` (ivclabpose.py:27) and the unshipped
    CNN back-ends ``backend.YOLOv3`` / ``backend.HRPose.SimpleHRNet`` (ivclabpose.py:29-30),
  * the flat module names the reference expects on ``sys.path``
    (src/_init_path.py:14-18, src/tracking/__init__.py:8-9, src/utils/__init__.py:8-9).

The reference has a module called ``hypothesis`` which collides with the PyPI package
of the same name that pytest auto-loads; the loader swaps ``sys.modules`` entries while
the reference modules bind their imports and restores them afterwards.
"""
from __future__ import annotations

import importlib
import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PAM_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "src")

_FLAT = {
    # flat module name -> file (relative to src/)
    "calculate": "utils/calculate.py",
    "default_limbs": "utils/default_limbs.py",
    "matching": "utils/matching.py",
    "construction": "utils/construction.py",
    "binary_integer_programming": "tracking/binary_integer_programming.py",
    "OneEuroFilter": "tracking/OneEuroFilter.py",
    "KalmanFilter": "tracking/KalmanFilter.py",
    "hypothesis": "tracking/hypothesis.py",
    "IterativeTracker": "tracking/IterativeTracker.py",
}

_cache = None


class EasyDict(dict):
    """Minimal attribute-dict standing in for ``easydict.EasyDict``."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:  # pragma: no cover
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def available() -> bool:
    return os.path.isdir(_SRC)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def load():
    """Return a namespace with the unmodified reference modules as attributes:
    ``calculate, matching, construction, hypothesis, IterativeTracker, ivclabpose``
    plus ``EasyDict``.  Raises ``RuntimeError`` when /root/reference is absent."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")

    import numpy as np

    if not hasattr(np, "float"):
        np.float = float  # type: ignore[attr-defined]
    if not hasattr(np, "int"):
        np.int = int  # type: ignore[attr-defined]

    stubs = {
        "matplotlib": _stub("matplotlib"),
        "matplotlib.pyplot": _stub("matplotlib.pyplot"),
        "cvxopt": _stub("cvxopt", glpk=None, matrix=None, spmatrix=None),
        "easydict": _stub("easydict", EasyDict=EasyDict),
        "backend": _stub("backend"),
        "backend.YOLOv3": _stub("backend.YOLOv3", YOLOv3=object),
        "backend.HRPose": _stub("backend.HRPose"),
        "backend.HRPose.SimpleHRNet": _stub("backend.HRPose.SimpleHRNet", HRNetPose=object),
    }
    stubs["matplotlib"].pyplot = stubs["matplotlib.pyplot"]

    touched = list(stubs) + list(_FLAT) + ["tracking", "tracking.IterativeTracker", "ivclabpose"]
    saved = {k: sys.modules.get(k) for k in touched}
    saved_path = list(sys.path)
    ns = types.SimpleNamespace(EasyDict=EasyDict)
    try:
        for k in touched:
            sys.modules.pop(k, None)
        for k, m in stubs.items():
            # keep a real installation if there is one (none in this image)
            try:
                if saved[k] is not None:
                    sys.modules[k] = saved[k]
                else:
                    importlib.import_module(k)
            except Exception:
                sys.modules[k] = m
        for name, rel in _FLAT.items():
            spec = importlib.util.spec_from_file_location(name, os.path.join(_SRC, rel))
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            spec.loader.exec_module(mod)
            setattr(ns, name, mod)
        # ivclabpose does `from tracking.IterativeTracker import IterativeTracker`
        pkg = _stub("tracking")
        pkg.__path__ = []  # mark as package
        pkg.IterativeTracker = ns.IterativeTracker
        sys.modules["tracking"] = pkg
        sys.modules["tracking.IterativeTracker"] = ns.IterativeTracker
        import torch.multiprocessing as tmp

        _ssm, _sss = tmp.set_start_method, tmp.set_sharing_strategy
        tmp.set_start_method = lambda *a, **k: None      # ivclabpose.py:20 side effect
        tmp.set_sharing_strategy = lambda *a, **k: None  # ivclabpose.py:21 side effect
        try:
            spec = importlib.util.spec_from_file_location("ivclabpose", os.path.join(_SRC, "ivclabpose.py"))
            mod = importlib.util.module_from_spec(spec)
            sys.modules["ivclabpose"] = mod
            spec.loader.exec_module(mod)
            ns.ivclabpose = mod
        finally:
            tmp.set_start_method, tmp.set_sharing_strategy = _ssm, _sss
    finally:
        for k in touched:
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
        sys.path[:] = saved_path
    _cache = ns
    return ns


def make_cameras(P, K, RT, w=1032, h=776):
    """Build the reference's ``Camera`` list through its own ``GetCameraParameters``
    (src/ivclabpose.py:162-181) without running the CNN-loading constructor."""
    ns = load()
    cls = ns.ivclabpose.ivclabpose
    obj = cls.__new__(cls)
    return obj.GetCameraParameters(dict(P=P, K=K, RT=RT), w, h)


def make_tracker(params: dict):
    """``IterativeTracker`` built from the 16 fields of src/ivclabpose.py:140-156."""
    ns = load()
    return ns.IterativeTracker.IterativeTracker(EasyDict(params))
