"""The drop-in modules keep the reference's call surface: every public function and
method the Iterative path uses exists under the same name with the same parameter names, order and defaults.
Compared against the unmodified reference modules by introspection (build container only: skipped where
/root/reference is absent); nothing is executed, so no GPU is needed."""
import inspect

import pytest

from oracle import ref_loader
from tests import util

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted")

FUNCTIONS = {
    "calculate": ["get_believe", "line2point_distance_3D", "line2line_distance_3D"],
    "matching": ["back_project_ray", "epipolar_distance", "epipolar_affinity", "epipolar_affinity_parallel",
                 "Greedy_matching", "BIP_matching"],
    "construction": ["SVD_pose_kernel", "SVD_pose_kernel_jf", "SVD_pose_kernel_parallel", "top_down_pose_kernel"],
}
METHODS = {
    ("hypothesis", "Hypothesis"): ["__init__", "size", "merge", "calculate_cost", "get_3dpose_jf"],
    ("IterativeTracker", "IterativeTracker"): ["__init__", "track_restart", "tracking"],
    ("OneEuroFilter", "OneEuroFilter"): ["__init__", "__call__"],
    ("KalmanFilter", "KalmanFilter"): ["__init__", "predict"],
}


def _params(fn):
    return [(p.name, p.default if p.default is not inspect.Parameter.empty else "<required>")
            for p in inspect.signature(fn).parameters.values()]


def test_function_signatures_match_reference():
    ref, ours = ref_loader.load(), util.load_dropin()
    for mod, names in FUNCTIONS.items():
        for name in names:
            assert _params(getattr(getattr(ours, mod), name)) == _params(getattr(getattr(ref, mod), name)), (mod, name)


def test_class_surfaces_match_reference():
    ref, ours = ref_loader.load(), util.load_dropin()
    for (mod, cls), names in METHODS.items():
        rc, oc = getattr(getattr(ref, mod), cls), getattr(getattr(ours, mod), cls)
        for name in names:
            assert _params(getattr(oc, name)) == _params(getattr(rc, name)), (cls, name)
    # the track life-cycle constants and read surface of IterTrack (ivclabpose.py:259-287)
    for k in ("Tentative", "Confirmed", "Deleted"):
        assert getattr(ours.IterativeTracker.TrackState, k) == getattr(ref.IterativeTracker.TrackState, k)
    for k in ("is_tentative", "is_confirmed", "is_deleted"):
        assert callable(getattr(ours.IterativeTracker.IterTrack, k))


def test_camera_surface_matches_reference():
    from pam_b200 import camera
    ref = ref_loader.load()
    rp, op = _params(ref.ivclabpose.Camera.__init__), _params(camera.Camera.__init__)
    assert op[: len(rp)] == rp                       # ours only appends optional keyword arguments
    assert all(d != "<required>" for _, d in op[len(rp):])
    for name in ("projectPoints", "projectPoints_parallel", "undistort", "undistort_points"):
        assert [n for n, _ in _params(getattr(camera.Camera, name))] == [n for n, _ in _params(getattr(ref.ivclabpose.Camera, name))], name
