"""libpam.so loads on a GPU-less box and exports every symbol include/pam.h declares
(no compute calls here)."""
import ctypes as C
import os
import re

import pytest

from pam_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "pam.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pam_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_capi.LIB_PATH):
        pytest.skip("libpam.so not built (run __graft_entry__.build())")
    lib = C.CDLL(_capi.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in pam.h but not exported"
    # the ctypes prototype table covers the header exactly
    assert sorted(_capi.declared_symbols()) == names


def test_abi_version_and_status_strings_without_gpu():
    if not os.path.exists(_capi.LIB_PATH):
        pytest.skip("libpam.so not built")
    lib = _capi.load_library()
    assert lib.pam_abi_version() == 2
    assert b"capacity" in lib.pam_status_string(-4)


def test_create_fails_loudly_without_device_or_on_bad_config():
    if not os.path.exists(_capi.LIB_PATH):
        pytest.skip("libpam.so not built")
    import torch
    lib = _capi.load_library()
    prm = dict(conf_threshold=.5, epi_threshold=60, init_threshold=30, joint_threshold=60, num_joints=14, n_init=3,
               max_age=10, alpha2d=70, lambda_a=3, lambda_t=5, sigma=.3, arm_sigma=.8)
    h = C.c_void_p()
    bad = _capi.make_config(dict(prm, num_joints=99), 5, 4, 8)
    assert lib.pam_create(C.byref(bad), 0, C.byref(h)) == -1          # PAM_E_INVALID
    if not torch.cuda.is_available():
        ok = _capi.make_config(prm, 5, 4, 8)
        assert lib.pam_create(C.byref(ok), 0, C.byref(h)) == -2       # PAM_E_CUDA: no CPU fallback
        assert b"no usable CUDA device" in lib.pam_last_error(None)


def test_library_is_sm100a_and_stages_frames_with_tma():
    """Static evidence from the built library (no GPU needed): the only embedded cubin is sm_100a, and the
    tracker kernel stages detections with a TMA bulk copy completed on an mbarrier (UBLKCP / SYNCS in SASS)."""
    import shutil
    import subprocess
    from pam_b200 import _capi
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    elfs = subprocess.run([cuobjdump, "-lelf", _capi.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", elfs))
    assert archs == {"sm_100a"}, elfs
    sass = subprocess.run([cuobjdump, "-sass", _capi.LIB_PATH], capture_output=True, text=True).stdout
    # keep the batch tracker kernels only (every launch-shape variant of every capacity class)
    parts = re.split(r"(?=\n\s*Function : )", sass)
    track = [p for p in parts if "Function : _Z17k_track_sequences" in p]
    assert len(track) >= 6
    for p in track:
        assert "UBLKCP" in p and "SYNCS" in p
    sass = "".join(track)
    assert "DFMA" in sass and "MUFU.RSQ64H" in sass        # FP64 path with the short MUFU-seeded reciprocal square root
