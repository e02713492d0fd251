"""Edge cases (empty frames, dead cameras, low / negative confidences, heavy outliers, exact ties)
against the oracle: kernel source on the host (CPU) and the CUDA path through the C ABI (-m gpu)."""
import numpy as np
import pytest

from tests import edge_cases, util
from pam_b200 import camera, synth, tracker


def _oracle(st):
    return util.run_oracle(st)


@pytest.mark.parametrize("name", sorted(edge_cases.CASES))
def test_edge_case_kernel_source_on_host(name):
    st = edge_cases.CASES[name]()
    cfg = util.stream_config(st, max_tracks=12)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = _oracle(st)
    util.compare_with_oracle(out, 0, st, oo, oa)
    if name in ("only_one_camera_alive", "low_confidence_never_initialises"):
        assert out["count"].sum() == 0
    if name == "everyone_leaves_and_returns":
        ids = set(out["ids"][out["ids"] >= 0].tolist())
        assert max(ids) >= 4          # fresh ids after the gap


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(edge_cases.CASES))
def test_edge_case_gpu(name):
    import torch
    st = edge_cases.CASES[name]()
    trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), synth.tracker_params(st.shape), 1,
                                  max_detections=st.dets.shape[2], max_tracks=12, arm_joints=st.shape.arm_joints)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    oo, oa, _ = _oracle(st)
    util.compare_with_oracle(out, 0, st, oo, oa)
