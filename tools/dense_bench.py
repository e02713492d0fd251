"""Dense crowd stress (BASELINE.json configs[3]): 31 cameras, 64 people, 19 joints per frame --
stateless affinity + triangulation + association batch.  Times each kernel with CUDA events and
reports algorithmic bytes / achieved GB/s against the measured HBM peak (development + profiles)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import pam_b200
from pam_b200 import synth, camera, ops

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(_ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(_ROOT, "MEASURED_PEAKS.json")) else 6650.0
st = synth.make_stream("dense", 0, 2, miss_prob=0.0, outlier_prob=0.01)
sh = st.shape
V, P, J = sh.V, sh.P, sh.J
cams = camera.GetCameraParameters(st.rig)
o = ops.GeometryOps(cams, J, synth.tracker_params("dense"))
poses = np.concatenate([st.dets[1, c, :st.counts[1, c]].astype(np.float64) for c in range(V)])
cam_idx = np.concatenate([np.full(st.counts[1, c], c) for c in range(V)])
M = len(poses)
d_pose = torch.from_numpy(poses).cuda()
d_cam = torch.from_numpy(cam_idx.astype(np.int32)).cuda()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()                      # evict L2 between timed iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts))


rows = []
ms = timeit(lambda: o.epipolar_allpairs(d_cam, d_pose, want_dist=False, as_numpy=False))
byts = M * J * 3 * 8 + 4 * M * M
rows.append(("pam_epipolar_allpairs (affinity only)", ms, byts, M * M * J * 2 * 22))
ms = timeit(lambda: o.epipolar_allpairs(d_cam, d_pose, want_dist=True, as_numpy=False), reps=5)
byts = M * J * 3 * 8 + 4 * M * M + 4 * M * M * J
rows.append(("pam_epipolar_allpairs (+ per-joint tensor)", ms, byts, M * M * J * 2 * 22))
# triangulation of all 64 people from all 31 views
pm = np.zeros((P, V, J, 3))
for c in range(V):
    for d in range(st.counts[1, c]):
        pm[st.person_of_det[1, c, d], c] = st.dets[1, c, d]
d_pm = torch.from_numpy(pm).cuda()
d_camv = torch.from_numpy(np.tile(np.arange(V), (P, 1)).astype(np.int32)).cuda()
d_w = torch.ones((P, V), dtype=torch.float64, device="cuda")
d_X = torch.empty((P, J, 3), dtype=torch.float64, device="cuda")
def tri():
    o._call(o.lib.pam_triangulate, o._p(d_pm), o._p(d_camv), o._p(d_w), None, None, P, V, o._p(d_X), o._stream())
ms = timeit(tri)
rows.append(("pam_triangulate (64 people x 19 joints x 31 views)", ms, P * V * J * 3 * 8 + P * J * 3 * 8, P * J * (V * 2 * 30 + 400)))
# association affinity: 64 tracks x 64 detections x 31 cameras
tracks = st.gt[0]
dets = st.frame_detections(1)
d_tr = torch.from_numpy(tracks).cuda(); d_dt = torch.ones(P, dtype=torch.int32, device="cuda")
d_dets = torch.from_numpy(np.stack([np.asarray(d) for d in dets])).cuda(); d_cnt = torch.full((V,), P, dtype=torch.int32, device="cuda")
d_aff = torch.empty((V, P, P), dtype=torch.float64, device="cuda")
def assoc():
    o._call(o.lib.pam_assoc_affinity, o._p(d_tr), o._p(d_dt), o._p(d_dets), o._p(d_cnt), P, P, o._p(d_aff), o._stream())
ms = timeit(assoc)
rows.append(("pam_assoc_affinity (31 cameras x 64 tracks x 64 detections)", ms, V * P * J * 3 * 8 + 8 * V * P * P, V * P * P * J * 30))
aff, counts = o.assoc_affinity(tracks, np.ones(P, int), dets, as_numpy=False)
cost = (-aff).contiguous()
out = torch.empty((V, P), dtype=torch.int32, device="cuda")
import ctypes as C
def assign():
    o._call(o.lib.pam_assign, o._p(cost), V, P, P, 0, o._p(out), o._stream())
ms = timeit(assign)
rows.append(("pam_assign (31 problems of 64 x 64)", ms, 8 * V * P * P, 0))
print(f"dense frame: V={V} P={P} J={J} M={M}")
for name, ms, b, fl in rows:
    print(f"{name:55s} {ms:8.3f} ms  {b/1e6:9.2f} MB  {b/ms/1e6:8.1f} GB/s ({100*b/ms/1e6/PEAK:5.1f}% of measured HBM peak)"
          + (f"  ~{fl/ms/1e9:6.2f} TFLOP/s fp64" if fl else ""))
# camera ingest, 31 cameras -> 961 fundamental matrices: device kernel vs the host's float32 torch loop
Kd = torch.from_numpy(np.stack([c.K for c in cams])).cuda(); RTd = torch.from_numpy(np.stack([c.RT for c in cams])).cuda()
RKd = torch.empty((V, 9), dtype=torch.float32, device="cuda"); posd = torch.empty((V, 3), dtype=torch.float64, device="cuda")
Fd = torch.empty((V, V, 9), dtype=torch.float32, device="cuda")
def ingest():
    rc = o.lib.pam_camera_ingest(0, V, Kd.data_ptr(), RTd.data_ptr(), RKd.data_ptr(), posd.data_ptr(), Fd.data_ptr(), o._stream())
    assert rc == 0
ms = timeit(ingest)
t0 = time.perf_counter(); camera.fundamental_tensor(np.stack([c.K for c in cams]), np.stack([c.RT for c in cams])); host_ms = (time.perf_counter() - t0) * 1e3
print(f"pam_camera_ingest (31 cameras, 961 matrices)            {ms:8.3f} ms on the device; host float32 torch loop {host_ms:.1f} ms")
