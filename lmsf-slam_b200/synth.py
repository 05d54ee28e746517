"""Seeded synthetic LiDAR sweeps for the scan-to-map hot path.

The reference ships no data (its PCD fixtures are absent,
src/MultiSensorFusionEstimator3D/src/test/registration/feature_registration_test.cpp:56-71),
so parity tests, the bench and the CPU baseline all run on this generator.

Scene  : axis-aligned room 60 x 40 x 6 m, sensor 1.8 m above the floor, 24 square
         pillars (0.6 m) on a 10 m grid, 6 wall-mounted boxes. Analytic ray casting.
Sensors: VLP-16  = 16 elevations -15..+15 deg step 2, 1800 azimuths  (28 800 rays)
         HDL-64  = 32 elevations +2..-8.33 step 1/3 and 32 from -8.83..-24.33 step 1/2
                   (the bin centres of LOAMFeatureProcessor_base.hpp:327-330), 2048 azimuths.
Firing order is azimuth-major (all lasers at az0, then az1, ...), float32 XYZI with
I = relative time az/2pi*0.1 (what RotaryLidar_preprocessing.hpp:100-104 leaves in
the intensity channel).  Range noise N(0, sigma) along the ray.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

ROOM_MIN = np.array([-30.0, -20.0, -1.8])
ROOM_MAX = np.array([30.0, 20.0, 4.2])


def _scene_boxes() -> np.ndarray:
    """(B, 2, 3) min/max corners of the solid boxes inside the room."""
    boxes = []
    for px in (-25.0, -15.0, -5.0, 5.0, 15.0, 25.0):
        for py in (-15.0, -5.0, 5.0, 15.0):
            boxes.append(((px - 0.3, py - 0.3, -1.8), (px + 0.3, py + 0.3, 4.2)))
    # wall-mounted boxes: two on each long wall, one on each short wall
    boxes.append(((-12.0, 19.0, -0.8), (-9.0, 20.0, 0.9)))
    boxes.append(((8.0, 19.2, -1.8), (11.0, 20.0, 0.2)))
    boxes.append(((-4.0, -20.0, -1.0), (-1.5, -19.1, 1.2)))
    boxes.append(((17.0, -20.0, -1.8), (19.5, -18.9, 0.5)))
    boxes.append(((29.0, -3.0, -1.2), (30.0, 1.0, 1.0)))
    boxes.append(((-30.0, 4.0, -1.8), (-28.8, 7.0, 0.4)))
    return np.asarray(boxes, dtype=np.float64)


BOXES = _scene_boxes()


@dataclass(frozen=True)
class Sensor:
    name: str
    elev_deg: np.ndarray  # (L,)
    n_az: int

    @property
    def n_scans(self) -> int:
        return int(self.elev_deg.shape[0])

    @property
    def n_rays(self) -> int:
        return self.n_scans * self.n_az


def vlp16() -> Sensor:
    return Sensor("VLP-16", np.arange(-15.0, 15.0 + 1e-9, 2.0), 1800)


def hdl64() -> Sensor:
    upper = 2.0 - np.arange(32) / 3.0
    lower = -8.83 - np.arange(32) / 2.0
    return Sensor("HDL-64", np.concatenate([upper, lower]), 2048)


def sensor_by_name(name: str) -> Sensor:
    return {"vlp16": vlp16, "hdl64": hdl64}[name.lower().replace("-", "")]()


def rot_zyx(yaw: float, pitch: float, roll: float) -> np.ndarray:
    cy, sy = math.cos(yaw), math.sin(yaw)
    cp, sp = math.cos(pitch), math.sin(pitch)
    cr, sr = math.cos(roll), math.sin(roll)
    rz = np.array([[cy, -sy, 0], [sy, cy, 0], [0, 0, 1.0]])
    ry = np.array([[cp, 0, sp], [0, 1.0, 0], [-sp, 0, cp]])
    rx = np.array([[1.0, 0, 0], [0, cr, -sr], [0, sr, cr]])
    return rz @ ry @ rx


def trajectory_pose(k: int, seq: int = 0, dt: float = 0.1):
    """Ground-truth sensor pose (R, t) of sweep k of sequence `seq`.

    S-curve between the pillar rows: v = 1.5 m/s along +x, |yaw rate| <= 0.2 rad/s,
    +-2 deg roll/pitch sinusoids.  `seq` shifts the phase so that the 8 batched
    sequences (BASELINE.json configs[3]) are distinct.
    """
    s = 1.5 * dt * k
    phase = 0.7 * seq
    x = -14.0 + s
    y = 3.0 * math.sin(2.0 * math.pi * s / 30.0 + phase)
    dy_ds = 3.0 * 2.0 * math.pi / 30.0 * math.cos(2.0 * math.pi * s / 30.0 + phase)
    yaw = 0.3 * math.atan(dy_ds)  # yaw rate stays below 0.2 rad/s at 1.5 m/s
    roll = math.radians(2.0) * math.sin(0.9 * s + phase)
    pitch = math.radians(2.0) * math.sin(0.6 * s + 1.0 + phase)
    return rot_zyx(yaw, pitch, roll), np.array([x, y, 0.0])


def ray_dirs(sensor: Sensor) -> np.ndarray:
    """(n_az * L, 3) unit directions in the sensor frame, azimuth-major."""
    az = (np.arange(sensor.n_az) * (2.0 * math.pi / sensor.n_az))[:, None]
    el = np.radians(sensor.elev_deg)[None, :]
    d = np.stack(
        [np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el) * np.ones_like(az)],
        axis=-1,
    )
    return d.reshape(-1, 3)


def _cast(origin: np.ndarray, dirs: np.ndarray) -> np.ndarray:
    """Range to the first surface for rays origin + r*dirs (world frame)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / dirs
        # room: we are inside, take the exit distance
        t1 = (ROOM_MIN - origin) * inv
        t2 = (ROOM_MAX - origin) * inv
        r = np.min(np.maximum(t1, t2), axis=1)
        # boxes: slab entry distance
        bmin = BOXES[:, 0, :][None, :, :]
        bmax = BOXES[:, 1, :][None, :, :]
        o = origin[None, None, :]
        ii = inv[:, None, :]
        ta = (bmin - o) * ii
        tb = (bmax - o) * ii
        tnear = np.max(np.minimum(ta, tb), axis=2)
        tfar = np.min(np.maximum(ta, tb), axis=2)
        hit = (tnear <= tfar) & (tfar > 0.0) & (tnear > 0.0)
        tbox = np.where(hit, tnear, np.inf).min(axis=1)
    return np.minimum(r, tbox)


def make_sweep(
    sensor: Sensor,
    k: int = 0,
    seq: int = 0,
    noise: float = 0.01,
    dropout: float = 0.0,
    pose=None,
    extrinsic=None,
    seed_base: int = 20260001,
) -> np.ndarray:
    """float32 (n, 4) XYZI points of sweep k in the sensor frame, firing order.

    `pose` overrides the trajectory; `extrinsic` = (R, t) of this LiDAR in the
    body frame (multi-LiDAR config).  Rays dropped by `dropout` are removed,
    as `removeNaNFromPointCloud` does upstream (MultiLidarSLAM_node.cpp:126-133).
    """
    rng = np.random.default_rng(seed_base + 1000 * seq + k)
    R, t = trajectory_pose(k, seq) if pose is None else pose
    if extrinsic is not None:
        Re, te = extrinsic
        t = R @ te + t
        R = R @ Re
    d_s = ray_dirs(sensor)
    d_w = d_s @ R.T
    rng_true = _cast(np.asarray(t, dtype=np.float64), d_w)
    rr = rng_true + rng.normal(0.0, noise, size=rng_true.shape) if noise > 0 else rng_true
    pts = d_s * rr[:, None]
    az_idx = np.repeat(np.arange(sensor.n_az), sensor.n_scans)
    inten = az_idx / sensor.n_az * 0.1
    out = np.concatenate([pts, inten[:, None]], axis=1).astype(np.float32)
    if dropout > 0.0:
        keep = rng.random(out.shape[0]) >= dropout
        out = out[keep]
    return np.ascontiguousarray(out)


def pose_to_qt(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """(qx, qy, qz, qw, tx, ty, tz) from a rotation matrix; w >= 0."""
    tr = np.trace(R)
    if tr > 0:
        s = math.sqrt(tr + 1.0) * 2
        w, x, y, z = 0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = [0.0, 0.0, 0.0]
        q[i] = 0.25 * s
        q[j] = (R[j, i] + R[i, j]) / s
        q[k] = (R[k, i] + R[i, k]) / s
        w = (R[k, j] - R[j, k]) / s
        x, y, z = q
    return np.array([x, y, z, w, t[0], t[1], t[2]], dtype=np.float64)


def qt_to_mat(p) -> np.ndarray:
    x, y, z, w = p[0], p[1], p[2], p[3]
    T = np.eye(4)
    T[:3, :3] = [
        [1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
        [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
        [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)],
    ]
    T[:3, 3] = p[4:7]
    return T


def rel_gt_pose(k: int, seq: int = 0) -> np.ndarray:
    """Ground-truth pose of sweep k relative to sweep 0 (the tracker's odom frame)."""
    R0, t0 = trajectory_pose(0, seq)
    Rk, tk = trajectory_pose(k, seq)
    return pose_to_qt(R0.T @ Rk, R0.T @ (tk - t0))
