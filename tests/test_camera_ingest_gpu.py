"""Camera ingest on the device (pam_camera_ingest) against the host ingest
that evaluates the reference's float32 expression (src/ivclabpose.py:35-46,162-181).

The reference's torch / LAPACK kernels and the CUDA kernel round float32 products in different
orders, so the comparison is at float32-rounding level: the device tensor must be as close to the
float64 evaluation of the same formula as the host's is, and a tracker fed with device-ingested
cameras must take the oracle's decisions."""
import numpy as np
import pytest

from tests import util
from pam_b200 import camera, synth, tracker

pytestmark = pytest.mark.gpu


def _f64_truth(K, RT):
    K, RT = K.astype(np.float64), RT.astype(np.float64)
    V = len(K)
    F = np.zeros((V, V, 3, 3))
    for a in range(V):
        Ra, ta = RT[a][:, :3], RT[a][:, 3]
        for b in range(V):
            Rb, tb = RT[b][:, :3], RT[b][:, 3]
            Rab = Ra @ Rb.T
            e = K[b] @ Rb @ Ra.T @ (ta - Rab @ tb)
            ex = np.array([[0, -e[2], e[1]], [e[2], 0, -e[0]], [-e[1], e[0], 0]])
            F[a, b] = np.linalg.inv(K[a]).T @ Rab @ K[b].T @ ex
    return F


@pytest.mark.parametrize("shape", ["campus", "shelf", "panoptic", "dense"])
def test_ingest_matches_host_to_float32_rounding(shape):
    st = synth.make_stream(shape, 1, 2)
    host = camera.GetCameraParameters(st.rig)
    dev = camera.GetCameraParameters(st.rig, device=0)
    V = len(host)
    Fh = np.stack([c.F for c in host]).astype(np.float64)
    Fd = np.stack([c.F for c in dev]).astype(np.float64)
    assert dev[0].F.dtype == np.float32 and dev[0].RK_INV.dtype == np.float32 and dev[0].position.dtype == np.float64
    Ft = _f64_truth(np.stack([c.K for c in host]), np.stack([c.RT for c in host]))
    scale = np.abs(Ft).max(axis=(2, 3), keepdims=True)
    off = ~np.eye(V, dtype=bool)                      # same-camera matrices are rounding noise on both sides and unused
    err_h = (np.abs(Fh - Ft) / scale)[off].max()
    err_d = (np.abs(Fd - Ft) / scale)[off].max()
    print(f"{shape}: F error vs float64 formula: host {err_h:.2e}, device {err_d:.2e}")
    assert err_d <= 3 * err_h + 1e-6
    for h, d in zip(host, dev):
        assert np.abs(d.RK_INV - h.RK_INV).max() <= 2e-5 * np.abs(h.RK_INV).max()
        assert np.abs(d.position - h.position).max() <= 1e-9 * max(1.0, np.abs(h.position).max())


def test_tracker_on_device_ingested_cameras_takes_the_oracle_decisions():
    import torch
    st = synth.make_stream("shelf", 7, 200, miss_prob=0.05, outlier_prob=0.03)
    cams = camera.GetCameraParameters(st.rig, device=0)
    trk = tracker.SequenceTracker(cams, synth.tracker_params(st.shape), num_sequences=1,
                                  max_detections=st.dets.shape[2], max_tracks=12, arm_joints=st.shape.arm_joints)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), nviews=True, assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    oo, oa, _ = util.run_oracle(st)                   # oracle on the host-ingested (reference) constants
    worst = util.compare_with_oracle(out, 0, st, oo, oa)
    assert out["count"].sum() > 0 and worst < 5e-4
