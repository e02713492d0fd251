"""Edge-case streams shared by the host-emulation (CPU) and GPU parity tests."""
import numpy as np

from pam_b200 import synth


def _blank(st, frames=None, cams=None):
    T, V = st.counts.shape
    for t in (frames if frames is not None else range(T)):
        for c in (cams if cams is not None else range(V)):
            st.counts[t, c] = 0
            st.dets[t, c] = 0
            st.person_of_det[t, c] = -1
    return st


def everyone_leaves_and_returns():
    """No detection in any camera for 14 frames (> max_age): every track dies, new ids afterwards."""
    return _blank(synth.make_stream("shelf", 101, 90), frames=range(30, 44))


def starts_empty():
    return _blank(synth.make_stream("campus", 102, 60), frames=range(0, 12))


def one_camera_dead():
    return _blank(synth.make_stream("shelf", 103, 80), cams=[2])


def only_one_camera_alive():
    """A single camera can never start or update a track (needs >= 2 views)."""
    return _blank(synth.make_stream("shelf", 104, 40), cams=[0, 1, 3, 4])


def low_confidence_never_initialises():
    st = synth.make_stream("shelf", 105, 50)
    st.dets[..., 2] *= 0.5          # mean confidence ~0.42 < conf_threshold 0.5: tracks are never started
    return st


def negative_confidences_are_skipped():
    """get_believe ignores joints with conf < 0 (calculate.py:11-13)."""
    st = synth.make_stream("shelf", 106, 50)
    st.dets[:, :, :, ::4, 2] = -1.0
    return st


def heavy_outliers_and_misses():
    return synth.make_stream("shelf", 107, 150, noise_px=3.0, miss_prob=0.3, outlier_prob=0.2, outlier_px=120.0)


def duplicated_person_two_identical_detections():
    """The same detection twice in one camera: exact ties in the affinity / cost matrices."""
    st = synth.make_stream("campus", 108, 60, P=2)
    T, V, D = st.counts.shape[0], st.counts.shape[1], st.dets.shape[2]
    dets = np.zeros((T, V, D + 1) + st.dets.shape[3:], np.float32)
    dets[:, :, :D] = st.dets
    counts = st.counts.copy()
    for t in range(10, T):
        if counts[t, 1] >= 1:
            dets[t, 1, counts[t, 1]] = dets[t, 1, 0]
            counts[t, 1] += 1
    return synth.Stream(st.shape, st.seq_id, st.rig, dets, counts, None, st.gt, st.meta)


def people_cross_close():
    """Two people walk through each other (association ambiguity)."""
    st = synth.make_stream("shelf", 109, 120, P=2, noise_px=2.0)
    return st


CASES = {f.__name__: f for f in [everyone_leaves_and_returns, starts_empty, one_camera_dead, only_one_camera_alive,
                                 low_confidence_never_initialises, negative_confidences_are_skipped,
                                 heavy_outliers_and_misses, duplicated_person_two_identical_detections,
                                 people_cross_close]}
