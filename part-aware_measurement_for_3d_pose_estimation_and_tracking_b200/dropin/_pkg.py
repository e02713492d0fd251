"""Resolve the hyphen-named package from the flat drop-in modules."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
pkg = importlib.import_module("part-aware_measurement_for_3d_pose_estimation_and_tracking_b200")
ops = importlib.import_module(pkg.__name__ + ".ops")
camera = importlib.import_module(pkg.__name__ + ".camera")
tracker = importlib.import_module(pkg.__name__ + ".tracker")
capi = importlib.import_module(pkg.__name__ + "._capi")
results = importlib.import_module(pkg.__name__ + ".results")
