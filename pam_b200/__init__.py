"""Short import alias for the package directory
``part-aware_measurement_for_3d_pose_estimation_and_tracking_b200`` (whose name, fixed by the
build contract, contains a hyphen and therefore cannot appear in an ``import`` statement)."""
import importlib as _il
import sys as _sys

_real = _il.import_module("part-aware_measurement_for_3d_pose_estimation_and_tracking_b200")
_sys.modules[__name__] = _real
