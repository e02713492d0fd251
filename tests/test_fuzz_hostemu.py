"""Randomised shapes / noise levels / thresholds: kernel source on the host vs the oracle (CPU).
Covers joint counts, camera counts and parameter values none of the named configurations use."""
import numpy as np
import pytest

from tests import util
from oracle import generic
from pam_b200 import _capi, synth


def _case(seed):
    rng = np.random.default_rng(seed)
    V = int(rng.integers(2, 7))
    P = int(rng.integers(1, 6))
    J = int(rng.integers(12, 23))
    res = [(360, 288, 440.0, 8.0), (1032, 776, 1060.0, 4.5), (1920, 1080, 1500.0, 4.2)][int(rng.integers(0, 3))]
    arm = tuple(sorted(rng.choice(J, size=int(rng.integers(0, 4)), replace=False).tolist()))
    shape = synth.Shape(f"fuzz{seed}", 50 + seed, V, P, J, 45, res[0], res[1], res[2], res[3], 2.4, 1.1, 0.4, arm)
    params = dict(conf_threshold=float(rng.choice([0.3, 0.5, 0.75])), epi_threshold=float(rng.choice([25, 60, 90.5])),
                  init_threshold=float(rng.choice([15, 30, 50])), joint_threshold=float(rng.choice([8, 15, 60])),
                  n_init=int(rng.integers(1, 5)), max_age=int(rng.integers(2, 11)), alpha2d=float(rng.choice([30, 70, 55.5])),
                  lambda_a=float(rng.choice([1, 3, 5])), lambda_t=float(rng.choice([2, 5, 6])),   # see test_dlt_conditioning.py for why not 10
                  sigma=float(rng.choice([0.3, 0.6, 1.1])), arm_sigma=float(rng.choice([0.5, 0.8, 1.4])),
                  num_joints=J, init_method="GD", w2d=0.4, w3d=0.6, alpha3d=0.15)
    kw = dict(noise_px=float(rng.choice([0.5, 1.0, 3.0])), miss_prob=float(rng.choice([0.0, 0.05, 0.25])),
              outlier_prob=float(rng.choice([0.0, 0.03, 0.15])), enter_stagger=int(rng.choice([0, 7])))
    if P > 1 and rng.random() < 0.5:
        kw["absences"] = [(0, 12, 12 + int(rng.integers(3, 20)))]
    return shape, params, kw, int(rng.integers(8, J - 1))


@pytest.mark.parametrize("seed", range(24))
def test_random_configuration(seed):
    shape, params, kw, min_valid = _case(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    cfg = _capi.make_config(params, shape.V, st.dets.shape[2], 14, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    util.compare_with_oracle(out, 0, st, oo, oa)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_random_configuration_gpu(seed):
    import torch
    from pam_b200 import camera, tracker
    shape, params, kw, min_valid = _case(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), params, 1, max_detections=st.dets.shape[2],
                                  max_tracks=14, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    oo, oa, _ = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    util.compare_with_oracle(out, 0, st, oo, oa)


def _large_case(seed):
    """Upper end of the supported shapes: 6-8 cameras, 5-8 people, 23-32 joints, 16 track slots."""
    rng = np.random.default_rng(100000 + seed)
    V = int(rng.integers(6, 9)); P = int(rng.integers(5, 9)); J = int(rng.integers(23, 33))
    res = [(1032, 776, 1060.0, 4.5), (1920, 1080, 1500.0, 4.2)][int(rng.integers(0, 2))]
    arm = tuple(sorted(rng.choice(J, size=int(rng.integers(0, 4)), replace=False).tolist()))
    shape = synth.Shape(f"big{seed}", 7000 + seed, V, P, J, 40, res[0], res[1], res[2], res[3], 2.4, 1.1, 0.4, arm)
    params = dict(conf_threshold=0.5, epi_threshold=float(rng.choice([40, 60])), init_threshold=float(rng.choice([15, 30])),
                  joint_threshold=float(rng.choice([8, 15])), n_init=int(rng.integers(1, 4)), max_age=int(rng.integers(3, 11)),
                  alpha2d=float(rng.choice([60, 70])), lambda_a=3.0, lambda_t=5.0, sigma=float(rng.choice([0.3, 0.6])),
                  arm_sigma=0.8, num_joints=J, init_method="GD", w2d=0.4, w3d=0.6, alpha3d=0.15)
    kw = dict(noise_px=float(rng.choice([0.5, 1.0, 2.0])), miss_prob=float(rng.choice([0.0, 0.05, 0.15])),
              outlier_prob=float(rng.choice([0.0, 0.02, 0.05])), enter_stagger=int(rng.choice([0, 5])))
    return shape, params, kw, int(rng.integers(10, J - 2))


@pytest.mark.parametrize("seed", [3, 41, 89, 91, 120, 149])     # V6-8, P5-8, J25-32 draws; all 150 seeds of this family were run once on the final source
def test_random_large_configuration(seed):
    shape, params, kw, min_valid = _large_case(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    cfg = _capi.make_config(params, shape.V, st.dets.shape[2], 16, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
    out = util.run_hostemu([st], cfg)
    assert out["status"].tolist() == [0]
    oo, oa, _ = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    util.compare_with_oracle(out, 0, st, oo, oa)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [89, 91])     # V8 D5 J32 and V6 D8 J30: > 48 KB of dynamic shared memory per CTA
def test_random_large_configuration_gpu(seed):
    import torch
    from pam_b200 import camera, tracker
    shape, params, kw, min_valid = _large_case(seed)
    st = synth.make_stream(shape, seed, shape.T, rig=synth.make_rig(shape), **kw)
    trk = tracker.SequenceTracker(camera.GetCameraParameters(st.rig), params, 1, max_detections=st.dets.shape[2],
                                  max_tracks=16, arm_joints=shape.arm_joints, min_valid_joints=min_valid)
    out = trk.run(torch.from_numpy(st.dets[None]).cuda(), torch.from_numpy(st.counts[None]).cuda(), nviews=True, assoc=True)
    assert trk.check().tolist() == [0]
    out = {k: v.cpu().numpy() for k, v in out.items()}
    oo, oa, _ = generic.run_stream(st, params, shape.arm_joints, min_valid, trace=True)
    util.compare_with_oracle(out, 0, st, oo, oa)
