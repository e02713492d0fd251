"""GPU-backed ``Hypothesis`` (reference: src/tracking/hypothesis.py:9-77)."""
import numpy as np

from _pkg import ops as _ops
from calculate import get_believe
from construction import SVD_pose_kernel_jf
from matching import epipolar_affinity


class Hypothesis:
    def __init__(self, cam, pts, epi_threshold=40):
        self.joints = len(pts)
        self.pose3d = None
        self.poses = [pts]
        self.cams = [cam]
        self.threshold = epi_threshold

    def size(self):
        return len(self.poses)

    def merge(self, o_cam, o_pose):
        self.cams.append(o_cam)
        self.poses.append(o_pose)

    def calculate_cost(self, o_cam, o_pose):
        """-> (cost, veto) (src/tracking/hypothesis.py:53-68) via ``pam_hypothesis_cost``."""
        cams = list(self.cams) + [o_cam]
        o = _ops.get_ops(cams, self.joints, dict(epi_threshold=self.threshold))
        k = len(self.poses)
        cost, veto = o.hypothesis_cost(np.arange(k), np.asarray(self.poses, dtype=np.float64), k,
                                       np.asarray(o_pose, dtype=np.float64)[None])
        return float(cost[0]), bool(veto[0])

    def get_3dpose_jf(self, init_threshold, lambda_t):
        """-> (cams, poses, pose3d, joints_views, ok) (src/tracking/hypothesis.py:23-44)."""
        n = len(self.cams)
        _, D = epipolar_affinity(self.cams, np.arange(n), self.poses, num_joints=self.joints)
        A = 1 - D / init_threshold                                                 # float32 like the reference
        o = _ops.get_ops(list(self.cams), self.joints)
        keep = o.view_filter(np.arange(n), np.ascontiguousarray(np.transpose(A, (2, 0, 1))), 'init')   # (J, n)
        joints_views = [[] for _ in range(n)]
        for j in range(self.joints):
            cnt = int(keep[j].sum())
            if cnt < 2:
                return [], [], [], [], False
            joints_views[cnt - 1].append(j)
        binary = np.repeat(keep.astype(int), 2, axis=1)
        pose3d = SVD_pose_kernel_jf(self.cams, [0] * n, self.poses, lambda_t, binary, joints_views)
        return self.cams, self.poses, pose3d, joints_views, True
