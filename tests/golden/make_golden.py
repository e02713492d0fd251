"""Generate the committed golden vectors by running the UNMODIFIED reference (J = 17, the only
joint count it supports) in the build container:

    python tests/golden/make_golden.py

Needs /root/reference (read-only); it does not exist on the GPU box, which only reads the .npz
files written here.  Two kinds of vectors:

  streams.npz     whole seeded synthetic sequences through ``IterativeTracker.tracking``: inputs
                  (rig, detections) and, per frame, the reported track ids / 3-D joints / per-joint
                  view counts and the track id every detection was associated with
  functions.npz   one input/output pair per live geometry function of src/utils and
                  src/tracking/hypothesis.py on random inputs
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

import pam_b200  # noqa: E402,F401
from pam_b200 import synth  # noqa: E402
from oracle import ref_loader  # noqa: E402

STREAMS = [
    ("shelf17", 41, 110, dict(enter_stagger=15, miss_prob=0.08, outlier_prob=0.04, absences=[(2, 50, 75)])),
    ("campus17", 42, 110, dict(miss_prob=0.06, outlier_prob=0.03)),
    ("shelf17", 43, 60, dict(noise_px=2.5, miss_prob=0.15, outlier_prob=0.08)),
]


def run_reference_stream(shape, seq, T, kw):
    st = synth.make_stream(shape, seq, T, **kw)
    V, J, MT = st.shape.V, st.shape.J, 12
    cams = ref_loader.make_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    trk = ref_loader.make_tracker(synth.tracker_params(shape))
    count = np.zeros(T, np.int32)
    ids = np.full((T, MT), -1, np.int32)
    joints = np.zeros((T, MT, J, 3))
    views = np.zeros((T, MT, J), np.int32)
    assoc = np.full((T, V, st.dets.shape[2]), -1, np.int32)
    for t in range(T):
        dets = st.frame_detections(t)
        trk.tracking(t, cams, [None] * V, st.frame_boxes(t), dets, "SVD")
        k = 0
        for tr in trk.tracks:
            # which detection went to which track this frame (IterativeTracker.py:155-160)
            for cid, v in tr.poses2d.items():
                if v["time"] == t and tr.poses3d[0]["time"] != t:
                    for d in range(len(dets[cid])):
                        if np.array_equal(dets[cid][d], v["pose"]):
                            assoc[t, cid, d] = tr.track_id
            if tr.time_since_update > 0 or not tr.is_confirmed():     # ivclabpose.py:266
                continue
            ids[t, k] = tr.track_id
            joints[t, k] = tr.poses3d[-1]["pose3d"]
            for nv, js in enumerate(tr.poses3d[-1]["joints_views"]):
                views[t, k, js] = nv + 1
            k += 1
        count[t] = k
    return st, dict(count=count, ids=ids, joints=joints, views=views, assoc=assoc)


def make_streams():
    out = {}
    for n, (shape, seq, T, kw) in enumerate(STREAMS):
        st, res = run_reference_stream(shape, seq, T, kw)
        p = f"s{n}_"
        out[p + "shape"] = np.array(shape)
        out[p + "P"], out[p + "K"], out[p + "RT"] = st.rig["P"], st.rig["K"], st.rig["RT"]
        out[p + "dets"], out[p + "counts"] = st.dets, st.counts
        for k, v in res.items():
            out[p + k] = v
        print(f"stream {n}: {shape} T={T}: {int(res['count'].sum())} reports, ids "
              f"{sorted(set(res['ids'][res['ids'] >= 0].tolist()))}")
    out["n_streams"] = np.array(len(STREAMS))
    np.savez_compressed(os.path.join(HERE, "streams.npz"), **out)


def make_functions():
    ns = ref_loader.load()
    M, C, K, H = ns.matching, ns.construction, ns.calculate, ns.hypothesis
    rng = np.random.default_rng(2024)
    J = 17
    rig = synth.make_rig("shelf17")
    cams = ref_loader.make_cameras(rig["P"], rig["K"], rig["RT"])
    V = len(cams)
    out = dict(P=rig["P"], K=rig["K"], RT=rig["RT"])
    out["cam_F"] = np.stack([c.F for c in cams])
    out["cam_RK_INV"] = np.stack([c.RK_INV for c in cams])
    out["cam_position"] = np.stack([c.position for c in cams])

    def person(noise=1.0):
        X = np.c_[rng.uniform(-1.5, 1.5, (J, 2)) * 0.3 + rng.uniform(-1.2, 1.2, 2), rng.uniform(0.0, 1.8, J)]
        poses = []
        for c in cams:
            pr = c.P.astype(np.float64) @ np.c_[X, np.ones(J)].T
            uv = (pr[:2] / pr[2]).T + rng.normal(0, noise, (J, 2))
            poses.append(np.c_[uv[:, 1], uv[:, 0], rng.uniform(0.5, 1.0, J)].astype(np.float32).astype(np.float64))
        return X, np.array(poses)

    # projectPoints_parallel
    X3 = np.stack([person()[0] for _ in range(6)])
    out["proj_in"] = X3
    out["proj_out"] = np.stack([c.projectPoints_parallel(X3) for c in cams])
    # epipolar_affinity_parallel / epipolar_affinity / epipolar_distance
    Xa, pa = person(1.0)
    _, pb = person(1.0)
    order = [2, 0, 4, 1]
    sub_cams = [cams[i] for i in order]
    pm = np.array([pa[i] for i in order])
    pm[1, 3, :2] += 90.0            # an outlier joint
    a_mean, a_D = M.epipolar_affinity_parallel(sub_cams, np.arange(len(order)), pm, J)
    out["eap_order"], out["eap_pose"], out["eap_mean"], out["eap_D"] = np.array(order), pm, a_mean, a_D
    poses_all = np.concatenate([pa, pb])                       # M = 2V detections, camera index per pose
    cam_idx = np.concatenate([np.arange(V), np.arange(V)])
    f_mean, f_D = M.epipolar_affinity(cams, cam_idx, poses_all, J)
    out["ea_pose"], out["ea_cam"], out["ea_mean"], out["ea_D"] = poses_all, cam_idx, f_mean, f_D
    out["ed_out"] = M.epipolar_distance(cams[1], pa[1], cams[3], pb[3])
    # Greedy_matching, both modes, many random symmetric affinity matrices
    B = 40
    g_A, g_keep_u, g_keep_i, g_uv, g_next = [], [], [], [], []
    for b in range(B):
        n = 4
        A = rng.uniform(-0.6, 1.0, (n, n))
        A = (A + A.T) / 2
        np.fill_diagonal(A, 1.0)
        pose1 = pm[:, b % J].reshape(-1, 1, 3)
        nxt = Xa[b % J] + rng.normal(0, 0.05, 3)
        ml, bl, _ = M.Greedy_matching(sub_cams, pose_mat=pose1, affinity_mat=A, next_pose=nxt)
        ml2, bl2, _ = M.Greedy_matching(sub_cams, affinity_mat=A.astype(np.float32), mode="init")
        g_A.append(A); g_keep_u.append(bl[::2]); g_keep_i.append(bl2[::2]); g_uv.append(pose1[:, 0, :]); g_next.append(nxt)
    out["gm_A"], out["gm_keep_update"], out["gm_keep_init"] = np.array(g_A), np.array(g_keep_u), np.array(g_keep_i)
    out["gm_pose"], out["gm_next"] = np.array(g_uv), np.array(g_next)
    # triangulation kernels
    Ts = [0, 1, 0, 3]
    keep = np.ones((J, len(order) * 2), dtype=int)
    joints_views = [[] for _ in order]
    for j in range(J):
        drop = rng.choice(len(order), size=rng.integers(0, 3), replace=False) if j % 3 == 0 else []
        if j == 5:
            drop = [0, 1, 2]
        for d in drop:
            keep[j, 2 * d:2 * d + 2] = 0
        joints_views[len(order) - len(drop) - 1].append(j)
    nxt = Xa + 0.01
    out["svd_Ts"], out["svd_keep"], out["svd_next"] = np.array(Ts), keep, nxt
    out["svd_jf"] = C.SVD_pose_kernel_jf(sub_cams, Ts, pm, 5, keep, joints_views, nxt)
    out["svd_parallel"] = C.SVD_pose_kernel_parallel(sub_cams, Ts, pm, 5)
    joints = [[pm[v, j] for v in range(len(order))] for j in range(J)]
    remains = [[v for v in range(len(order)) if keep[j, 2 * v]] for j in range(J)]
    out["svd_old"] = np.array(C.SVD_pose_kernel(sub_cams, Ts, joints, remains, 5, nxt), dtype=np.float64)
    # rays
    pts = np.flip(pa[2][:, :2], axis=1)
    dirs = M.back_project_ray(cams[2].RK_INV, cams[2].position, pts)
    out["ray_uv"], out["ray_dirs"] = pts, dirs
    out["ray_dist"] = K.line2point_distance_3D(cams[2].position, dirs, Xa + 0.03)
    out["ray_X"] = Xa + 0.03
    # Hypothesis
    hyp = H.Hypothesis(cams[0], pa[0], 60)
    hyp.merge(cams[2], pa[2])
    c1, v1 = hyp.calculate_cost(cams[3], pa[3])
    c2, v2 = hyp.calculate_cost(cams[3], pb[3])
    out["hyp_cost"] = np.array([c1, c2])
    out["hyp_veto"] = np.array([v1, v2])
    hyp.merge(cams[3], pa[3])
    hyp.merge(cams[4], pa[4])
    _, _, p3d, jv, ok = hyp.get_3dpose_jf(30, 5)
    out["hyp_pose3d"], out["hyp_ok"] = np.array(p3d), np.array(ok)
    nvj = np.zeros(J, np.int32)
    for k, js in enumerate(jv):
        nvj[js] = k + 1
    out["hyp_views"] = nvj
    out["pa"], out["pb"] = pa, pb
    out["believe"] = np.array([K.get_believe(pa[0]), K.get_believe(pb[1])])
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **out)
    print("functions.npz written:", len(out), "arrays")


def make_results():
    """results.pkl.gz: the tuple the reference's facade returns per frame (src/ivclabpose.py:216-287,
    PersonTrack_Project3DPose run UNMODIFIED on top of the unmodified tracker) for the first golden stream: camera ids
    and 2-D poses matched this frame per track (dict-insertion order), person ids, 3-D poses (n, 3, J), joints_views,
    3-D person ids."""
    import pickle
    shape, seq, T, kw = STREAMS[0]
    st = synth.make_stream(shape, seq, T, **kw)
    V = st.shape.V
    ns = ref_loader.load()
    cls = ns.ivclabpose.ivclabpose
    fac = cls.__new__(cls)                                   # skip the CNN-loading constructor
    fac.cameras = ref_loader.make_cameras(st.rig["P"], st.rig["K"], st.rig["RT"])
    fac.tracker = ref_loader.make_tracker(synth.tracker_params(shape))
    frames = []
    for t in range(T):
        dets = st.frame_detections(t)
        person_bbox_list, dump_results = [], []
        for c in range(V):
            # the facade swaps columns 0 and 1 of the pose-net keypoints (ivclabpose.py:238-244): hand it (u, v, .)
            items = [{"bbox": [0, 0, 1, 1], "keypoints": np.stack([d[:, 1], d[:, 0], d[:, 2]], 1).reshape(-1).tolist(),
                      "keypoints_score": d[:, 2].tolist(), "feature": [0.0]} for d in dets[c]]
            person_bbox_list.append([{"data": None}] * len(items))
            dump_results.append(items)
        cam_ids, pts, person_ids, pts3d, jviews, p3ids, _, _, _ = fac.PersonTrack_Project3DPose(t, person_bbox_list, dump_results, "SVD")
        frames.append(dict(camera_ids=[list(map(int, x)) for x in cam_ids], pts=[[np.asarray(p, np.float32) for p in x] for x in pts],   # float32-exact values
                           person_ids=[list(map(int, x)) for x in person_ids], pts3d=np.asarray(pts3d, np.float64),
                           joints_views=[[list(map(int, b)) for b in jv] for jv in jviews],
                           person3d_ids=[int(x) for x in p3ids]))
    import gzip
    with gzip.open(os.path.join(HERE, "results.pkl.gz"), "wb") as f:
        pickle.dump(dict(shape=shape, seq=seq, T=T, kw=kw, frames=frames), f, protocol=4)
    print("results.pkl.gz:", T, "frames,", sum(len(fr["person3d_ids"]) for fr in frames), "reported tracks")


if __name__ == "__main__":
    assert ref_loader.available(), "the golden vectors can only be generated where /root/reference exists"
    import warnings
    warnings.filterwarnings("ignore")
    make_streams()
    make_functions()
    make_results()
