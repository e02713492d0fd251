"""Largest supported shapes on the GPU (> 48 KB of dynamic shared memory per CTA) against the host build of the
same kernel source (development aid; the oracle comparison of these shapes runs on the host build in
tests/test_fuzz_hostemu.py::test_random_large_configuration)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from tests.test_fuzz_hostemu import _large_case
from pam_b200 import _capi, camera, synth, tracker

for seed in (91, 89):
    shape, params, kw, mv = _large_case(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    cfg = _capi.make_config(params, shape.V, st.dets.shape[2], 16, arm_joints=shape.arm_joints, min_valid_joints=mv)
    host = util.run_hostemu([st], cfg)
    trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), params, 1, max_detections=st.dets.shape[2],
                                  max_tracks=16, arm_joints=shape.arm_joints, min_valid_joints=mv)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), nviews=True, assoc=True)
    st_ = trk.check().tolist()
    out = {k: v.cpu().numpy() for k, v in out.items()}
    same = np.array_equal(out["count"], host["count"]) and np.array_equal(out["assoc"], host["assoc"])
    worst = 0.0
    for t in range(st.T):
        k = int(host["count"][0, t])
        same = same and np.array_equal(out["ids"][0, t, :k], host["ids"][0, t, :k]) and np.array_equal(out["nviews"][0, t, :k], host["nviews"][0, t, :k])
        if k:
            worst = max(worst, float(np.abs(out["joints"][0, t, :k] - host["joints"][0, t, :k]).max()))
    print(f"seed {seed}: V{shape.V} D{st.dets.shape[2]} J{shape.J} status {st_} reports {int(out['count'].sum())} decisions identical {same} max |dX| {worst:.2e}")
