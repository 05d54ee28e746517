"""Multi-LiDAR rig: the calibration branch of MultiLidarSystem::process() (System/ML_System.hpp:239-323) on device
contexts — host-side control flow only, every point operation is a C-ABI call.

status 0  every LiDAR runs its own tracker (the reference's `omp parallel for`, here one host thread per context);
          the motion increments of LiDAR 0 and LiDAR 1 feed the hand-eye initialisation (lmsf_handeye_*).
status 1  only the primary is tracked; the auxiliary sweep is registered against the PRIMARY's local map from
          primary_pose * extrinsic (lmsf_tracker_register_aux) and the extrinsic is re-derived from the result.
"""
from __future__ import annotations

import threading

import numpy as np

from . import capi


def _qmul(a, b):
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array([aw * bx + ax * bw + ay * bz - az * by, aw * by + ay * bw + az * bx - ax * bz,
                     aw * bz + az * bw + ax * by - ay * bx, aw * bw - ax * bx - ay * by - az * bz])


def _qrot(q, v):
    u = np.asarray(q[:3])
    uv = 2.0 * np.cross(u, v)
    return v + q[3] * uv + np.cross(u, uv)


def pose_mul(a, b):
    """{qx,qy,qz,qw,tx,ty,tz} composition a * b."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    q = _qmul(a[:4], b[:4])
    q /= np.linalg.norm(q)
    return np.concatenate([q, _qrot(a[:4], b[4:]) + a[4:]])


def pose_inv(a):
    a = np.asarray(a, float)
    qi = np.array([-a[0], -a[1], -a[2], a[3]])
    return np.concatenate([qi, -_qrot(qi, a[4:])])


class MultiLidarRig:
    """Two LiDARs (primary = 0, auxiliary = 1) with online extrinsic estimation."""

    def __init__(self, lib: capi.Library, devices=(0, 0), **params):
        self.ctx = [lib.context(d, **params) for d in devices]
        self.handeye = capi.HandEye(lib)
        self.status = 0                       # EXTRINSIC_CALIB_STATUS_
        self.extrinsic = np.array(capi.IDENTITY_POSE, float)    # lidar_lidar_estrinsic_
        self.pose = [np.array(capi.IDENTITY_POSE, float) for _ in devices]   # pose_lidar_cur_
        self.singular_values = np.zeros(4)

    def close(self):
        for c in self.ctx:
            c.close()
        self.handeye.close()

    def process(self, sweeps, stamp: float):
        """One synchronised pair of sweeps (float32 (n, 4) arrays).  Returns the status after the step."""
        if self.status == 0:
            out = [None, None]

            def run(i):
                out[i] = self.ctx[i].tracker_step(sweeps[i], stamp)

            th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            for i in range(2):
                self.pose[i] = pose_mul(self.pose[i], out[i][1])
            if out[0][2]["first"]:
                return self.status             # the initialising sweep has no motion increment
            if self.handeye.add_pose(out[0][1], out[1][1]):
                ok, ext, sv = self.handeye.calibrate()
                self.singular_values = sv
                if ok:
                    self.extrinsic = ext
                    self.status = 1
        else:
            p, d, _ = self.ctx[0].tracker_step(sweeps[0], stamp)
            self.pose[0] = pose_mul(self.pose[0], d)
            sub, _ = self.ctx[0].tracker_register_aux(sweeps[1], pose_mul(p, self.extrinsic))
            self.extrinsic = pose_mul(pose_inv(p), sub)
            self.pose[1] = pose_mul(self.pose[0], self.extrinsic)
        return self.status
