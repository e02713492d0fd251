"""ctypes view of ``include/pam.h`` and the loader for ``libpam.so``.

The product path is the CUDA library: ``load_library()`` raises when it is missing or cannot be
loaded -- there is no CPU fallback anywhere in this package."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libpam.so")


class PamConfig(C.Structure):
    _fields_ = [
        ("num_cameras", C.c_int32), ("num_joints", C.c_int32), ("max_detections", C.c_int32),
        ("max_tracks", C.c_int32), ("n_init", C.c_int32), ("max_age", C.c_int32),
        ("min_valid_joints", C.c_int32), ("stale_window", C.c_int32),
        ("arm_joint_mask", C.c_uint32), ("max_report", C.c_uint32),
        ("conf_threshold", C.c_double), ("epi_threshold", C.c_double), ("init_threshold", C.c_double),
        ("joint_threshold", C.c_double), ("alpha2d", C.c_double), ("lambda_a", C.c_double),
        ("lambda_t", C.c_double), ("sigma", C.c_double), ("arm_sigma", C.c_double),
        ("veto_believe", C.c_double),
    ]


class PamStateLayout(C.Structure):
    _fields_ = [
        ("seq_bytes", C.c_int64), ("off_header", C.c_int64), ("off_meta", C.c_int64),
        ("off_hist", C.c_int64), ("off_view", C.c_int64), ("off_vel", C.c_int64),
        ("off_nviews", C.c_int64), ("off_margin", C.c_int64), ("meta_ints", C.c_int32), ("hist_ring", C.c_int32),
        ("max_views", C.c_int32), ("max_order", C.c_int32), ("header_ints", C.c_int32), ("n_margins", C.c_int32),
    ]


class PamStreamViews(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dets", "counts", "out_count", "out_ids", "out_joints", "out_nviews",
                                          "out_assoc", "out_timing", "out_status", "out_vlist")]


def make_config(params, num_cameras, max_detections, max_tracks, arm_joints=(9, 10), min_valid_joints=10,
                stale_window=3, veto_believe=0.5, max_report=0) -> PamConfig:
    """``params``: mapping / attribute object with the ``iter_args`` fields of
    src/ivclabpose.py:140-156 (``conf_threshold, epi_threshold, init_threshold, joint_threshold,
    num_joints, n_init, max_age, alpha2d, lambda_a, lambda_t, sigma, arm_sigma``)."""
    g = (lambda k: params[k]) if isinstance(params, dict) else (lambda k: getattr(params, k))
    mask = 0
    for j in arm_joints:
        if 0 <= j < g("num_joints"):
            mask |= 1 << j
    return PamConfig(
        num_cameras=num_cameras, num_joints=g("num_joints"), max_detections=max_detections,
        max_tracks=max_tracks, n_init=g("n_init"), max_age=g("max_age"),
        min_valid_joints=min_valid_joints, stale_window=stale_window, arm_joint_mask=mask, max_report=max_report,
        conf_threshold=g("conf_threshold"), epi_threshold=g("epi_threshold"),
        init_threshold=g("init_threshold"), joint_threshold=g("joint_threshold"), alpha2d=g("alpha2d"),
        lambda_a=g("lambda_a"), lambda_t=g("lambda_t"), sigma=g("sigma"), arm_sigma=g("arm_sigma"),
        veto_believe=veto_believe)


_lib = None

# every symbol include/pam.h declares: (name, restype, argtypes)
_P = C.c_void_p
_PROTOTYPES = [
    ("pam_abi_version", C.c_int, []),
    ("pam_status_string", C.c_char_p, [C.c_int]),
    ("pam_last_error", C.c_char_p, [_P]),
    ("pam_create", C.c_int, [C.POINTER(PamConfig), C.c_int, C.POINTER(_P)]),
    ("pam_destroy", C.c_int, [_P]),
    ("pam_set_cameras", C.c_int, [_P, _P, _P, _P, _P]),
    ("pam_get_state_layout", C.c_int, [_P, C.POINTER(PamStateLayout)]),
    ("pam_track_reset", C.c_int, [_P, _P, C.c_int32, _P]),
    ("pam_track_sequences", C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("pam_track_status", C.c_int, [_P, _P, C.c_int32, _P, _P]),
    ("pam_track_margins", C.c_int, [_P, _P, C.c_int32, _P, _P]),
    ("pam_track_launch_info", C.c_int, [_P, C.c_int32, _P]),
    ("pam_sm_clock_khz", C.c_int, [_P, _P]),
    ("pam_track_sequences_host", C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    ("pam_track_host_status", C.c_int, [_P, C.c_int32, _P]),
    ("pam_track_state_to_host", C.c_int, [_P, C.c_int32, _P]),
    ("pam_stream_open", C.c_int, [_P, C.c_int32]),
    ("pam_stream_buffers", C.c_int, [_P, C.POINTER(PamStreamViews)]),
    ("pam_stream_step", C.c_int, [_P, C.c_int32]),
    ("pam_stream_submit", C.c_int, [_P, C.c_int32]),
    ("pam_stream_wait", C.c_int, [_P]),
    ("pam_stream_close", C.c_int, [_P]),
    ("pam_launch_count", C.c_int64, [_P]),
    ("pam_project_points", C.c_int, [_P, _P, C.c_int32, _P, _P]),
    ("pam_assoc_affinity", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    ("pam_assign", C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    ("pam_epipolar_pairs", C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P]),
    ("pam_epipolar_allpairs", C.c_int, [_P, _P, _P, C.c_int32, _P, _P, _P]),
    ("pam_epipolar_distance", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, _P, _P]),
    ("pam_view_filter", C.c_int, [_P, C.c_int32, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    ("pam_triangulate", C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P]),
    ("pam_mean_confidence", C.c_int, [_P, _P, C.c_int32, _P, _P]),
    ("pam_hypothesis_cost", C.c_int, [_P, _P, _P, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    ("pam_eval_pcp", C.c_int, [_P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                               C.c_int32, C.c_double, _P, _P, _P]),
    ("pam_eval_panoptic_match", C.c_int, [_P, _P, _P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    ("pam_ray_distance", C.c_int, [_P, C.c_int32, _P, _P, C.c_int32, _P, _P, _P]),
    ("pam_one_euro", C.c_int, [_P, _P, C.c_int32, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P]),
    ("pam_kalman9", C.c_int, [_P, _P, _P, C.c_int32, C.c_double, _P, _P, _P]),
    ("pam_top_down", C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, _P, _P, _P, _P]),
    ("pam_camera_ingest", C.c_int, [C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
]


def declared_symbols():
    return [p[0] for p in _PROTOTYPES]


def load_library(path: str | None = None):
    """Load ``libpam.so`` and bind every declared symbol.  Raises ``RuntimeError`` if the library
    has not been built (``python -c 'import __graft_entry__ as g; g.build()'``)."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or os.environ.get("PAM_LIBRARY") or LIB_PATH   # PAM_LIBRARY: development builds (phase timing)
    if not os.path.exists(p):
        raise RuntimeError(f"libpam.so not found at {p}: build it with __graft_entry__.build(); "
                           "this package has no CPU fallback")
    lib = C.CDLL(p)
    for name, res, args in _PROTOTYPES:
        fn = getattr(lib, name)   # AttributeError if a declared symbol is missing
        fn.restype = res
        fn.argtypes = args
    if path is None:
        _lib = lib
    return lib


_fast = None


def load_pyfast():
    """The CPython packing helper of the per-frame drop-in path (csrc/pam_pyfast.c), bound to this library's
    pam_stream_submit / pam_stream_wait; None when it has not been built (the drop-in then packs with numpy)."""
    global _fast
    if _fast is None:
        import importlib.util
        import sysconfig
        path = os.path.join(_HERE, "csrc", "pam_pyfast" + (sysconfig.get_config_var("EXT_SUFFIX") or ".so"))
        if not os.path.exists(path):
            _fast = False
            return None
        spec = importlib.util.spec_from_file_location("pam_pyfast", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        lib = load_library()
        mod.bind(C.cast(lib.pam_stream_submit, C.c_void_p).value, C.cast(lib.pam_stream_wait, C.c_void_p).value)
        _fast = mod
    return _fast or None
