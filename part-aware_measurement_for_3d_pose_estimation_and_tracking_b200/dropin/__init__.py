"""Drop-in replacements for the reference's flat modules ``calculate``, ``matching``,
``construction``, ``hypothesis`` and ``IterativeTracker`` (src/utils, src/tracking).

The reference finds its modules through ``sys.path`` injection (src/_init_path.py:14-18,
src/tracking/__init__.py:8-9), so ``install()`` does the same with this directory: afterwards
``from matching import epipolar_affinity_parallel`` or ``from IterativeTracker import
IterativeTracker`` resolve to the GPU-backed versions with the reference's signatures."""
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))


def install():
    for p in (_ROOT, _HERE):
        if p not in sys.path:
            sys.path.insert(0, p)
    return _HERE
