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


class DistributedRig:
    """The same rig with ONE LiDAR PER GPU (BASELINE config 3): rank r of the process group owns LiDAR r and GPU r
    (rank 0 = primary).  Host control flow only; every point operation is a C-ABI call on the rank's own context.

    status 0 (System/ML_System.hpp:248-256)  every rank tracks its own sweep; the 7-double motion increments are
             all-gathered and rank 0 feeds its hand-eye initialisations (one per auxiliary LiDAR); status and extrinsics
             are broadcast, so that every rank switches in the same step.
    status 1 (:296-310)  the primary tracks; an auxiliary rank only EXTRACTS (lmsf_extract_features_to_dev) and ships
             its features — at most a few hundred KB — to the primary's GPU (NCCL send / recv over NVLink);
             the primary registers them against ITS local map (lmsf_tracker_register_aux_features_dev) from
             primary_pose * extrinsic and re-derives the extrinsic.

    world == 1 (tests on one GPU): `peers` holds the other LiDARs' contexts in this process and the features move by a
    device-to-device copy instead of NCCL — the library calls and their order are the same.
    """

    def __init__(self, lib: capi.Library, ctx, rank: int, world: int, peers=None, device=None, max_features: int = 1 << 17):
        import torch
        import torch.distributed as dist

        self.torch, self.dist = torch, dist
        self.lib, self.ctx, self.rank, self.world = lib, ctx, int(rank), int(world)
        self.peers = list(peers) if peers else []
        self.n_lidar = self.world if self.world > 1 else 1 + len(self.peers)
        self.dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        self.stream = torch.cuda.ExternalStream(ctx.stream(), device=self.dev)
        self.cap = int(max_features)
        self.status = 0
        self.extrinsic = [np.array(capi.IDENTITY_POSE, float) for _ in range(self.n_lidar)]   # [0] unused
        self.pose = np.array(capi.IDENTITY_POSE, float)            # this rank's LiDAR in its own odometry frame
        self.aux_pose = {}                                         # rank 0, status 1: pose of every auxiliary LiDAR
        self.handeye = [capi.HandEye(lib) for _ in range(self.n_lidar)] if self.rank == 0 else []
        self.calibrated = [False] * self.n_lidar
        with torch.cuda.stream(self.stream):
            # status 1: the aux rank's features (sender side) / one receive buffer per auxiliary LiDAR (rank 0)
            n_buf = (self.n_lidar - 1) if self.rank == 0 else 1
            self.feat = [torch.empty((self.cap, 4), dtype=torch.float32, device=self.dev) for _ in range(max(1, n_buf))]
            self.inc = torch.zeros(8, dtype=torch.float64, device=self.dev)
            self.inc_all = torch.zeros((max(self.world, 1), 8), dtype=torch.float64, device=self.dev)
            self.ctl = torch.zeros(1 + 7 * self.n_lidar, dtype=torch.float64, device=self.dev)
            self.cnt = torch.zeros(2, dtype=torch.int32, device=self.dev)
            self.cnt_all = torch.zeros((max(self.world, 1), 2), dtype=torch.int32, device=self.dev)
        self.shipped_bytes = 0

    def close(self):
        for h in self.handeye:
            h.close()

    # ---- status 0 on rank 0: increments of LiDAR 0 and LiDAR i -> hand-eye i
    def _calibrate(self, deltas, first):
        if first:
            return
        for i in range(1, self.n_lidar):
            if self.calibrated[i]:
                continue
            if self.handeye[i].add_pose(deltas[0], deltas[i]):
                ok, ext, _ = self.handeye[i].calibrate()
                if ok:
                    self.extrinsic[i] = ext
                    self.calibrated[i] = True
        if all(self.calibrated[1:]):
            self.status = 1

    def process(self, sweep, stamp: float, peer_sweeps=None):
        """One synchronised set of sweeps: `sweep` is this rank's (world > 1) or the primary's with the others in
        peer_sweeps (world == 1).  Returns the status after the step."""
        torch, dist = self.torch, self.dist
        if self.status == 0:
            p, d, st = self.ctx.tracker_step(sweep, stamp)
            self.pose = pose_mul(self.pose, d)
            if self.world > 1:
                with torch.cuda.stream(self.stream):
                    self.inc.copy_(torch.from_numpy(np.concatenate([d, [float(st["first"])]])))
                    dist.all_gather_into_tensor(self.inc_all, self.inc)
                    if self.rank == 0:
                        inc = self.inc_all.cpu().numpy()
                        self._calibrate([inc[i, :7] for i in range(self.n_lidar)], bool(inc[0, 7]))
                        self.ctl.copy_(torch.from_numpy(np.concatenate([[float(self.status)]] +
                                                                       [e for e in self.extrinsic])))
                    dist.broadcast(self.ctl, 0)
                    ctl = self.ctl.cpu().numpy()
                self.status = int(ctl[0])
                self.extrinsic = [ctl[1 + 7 * i: 8 + 7 * i].copy() for i in range(self.n_lidar)]
            else:
                deltas = [d]
                for c, s in zip(self.peers, peer_sweeps):
                    deltas.append(c.tracker_step(s, stamp)[1])
                self._calibrate(deltas, bool(st["first"]))
            return self.status
        # ---- status 1
        if self.world > 1:
            with torch.cuda.stream(self.stream):
                if self.rank == 0:
                    p, d, _ = self.ctx.tracker_step(sweep, stamp)
                    self.pose = pose_mul(self.pose, d)
                    self.cnt.zero_()
                else:
                    ne, ns = self.ctx.extract_features_to_dev(sweep, self.feat[0].data_ptr(), self.cap)
                    self.cnt.copy_(torch.tensor([ne, ns], dtype=torch.int32))
                dist.all_gather_into_tensor(self.cnt_all, self.cnt)
                cnt = self.cnt_all.cpu().numpy()
                if self.rank == 0:
                    ops = [dist.P2POp(dist.irecv, self.feat[i - 1][: int(cnt[i].sum())], i) for i in range(1, self.n_lidar)
                           if int(cnt[i].sum()) > 0]
                    for r in (dist.batch_isend_irecv(ops) if ops else []):
                        r.wait()
                    for i in range(1, self.n_lidar):
                        ne, ns = int(cnt[i, 0]), int(cnt[i, 1])
                        self.shipped_bytes += 16 * (ne + ns)
                        sub, _ = self.ctx.tracker_register_aux_features_dev(self.feat[i - 1].data_ptr(), ne, ns,
                                                                            pose_mul(p, self.extrinsic[i]))
                        self.extrinsic[i] = pose_mul(pose_inv(p), sub)
                        self.aux_pose[i] = pose_mul(self.pose, self.extrinsic[i])
                else:
                    n = int(cnt[self.rank].sum())
                    if n > 0:
                        for r in dist.batch_isend_irecv([dist.P2POp(dist.isend, self.feat[0][:n], 0)]):
                            r.wait()
                    self.shipped_bytes += 16 * n
        else:
            p, d, _ = self.ctx.tracker_step(sweep, stamp)
            self.pose = pose_mul(self.pose, d)
            for i, (c, s) in enumerate(zip(self.peers, peer_sweeps), start=1):
                ne, ns = c.extract_features_to_dev(s, self.feat[i - 1].data_ptr(), self.cap)   # on the aux context
                self.torch.cuda.synchronize()                                                  # (its own stream)
                self.shipped_bytes += 16 * (ne + ns)
                sub, _ = self.ctx.tracker_register_aux_features_dev(self.feat[i - 1].data_ptr(), ne, ns,
                                                                    pose_mul(p, self.extrinsic[i]))
                self.extrinsic[i] = pose_mul(pose_inv(p), sub)
                self.aux_pose[i] = pose_mul(self.pose, self.extrinsic[i])
        return self.status
