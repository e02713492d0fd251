"""Build + load the host-emulation harness (TEST INFRASTRUCTURE ONLY; see hostemu.cpp)."""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libpam_hostemu.so")
_SRC = os.path.join(_HERE, "hostemu.cpp")
_CSRC = os.path.join(_HERE, "..", "..", "part-aware_measurement_for_3d_pose_estimation_and_tracking_b200", "csrc")


def build(force=False):
    deps = [_SRC] + [os.path.join(_CSRC, f) for f in ("pam_core.h", "pam_track.h", "pam_host.h")]
    deps.append(os.path.join(_HERE, "..", "..", "include", "pam.h"))
    if (not force and os.path.exists(_SO)
            and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps)):
        return _SO
    os.makedirs(os.path.dirname(_SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wno-array-bounds",
                           "-o", _SO, _SRC])
    return _SO


def load():
    lib = C.CDLL(build())
    lib.hostemu_track_sequences.restype = C.c_int
    lib.hostemu_state_layout.restype = C.c_int
    return lib
