"""Host-side helpers with the names of the reference's utils module (rocket_simulation/utils.py).

Only what the host side of the boundary needs: attitude conversion when initial conditions are
marshalled (utils.py:129-136), Euler angles for the result series (utils.py:139-144) and the JSON
helpers (utils.py:208-223).  The flight physics itself lives in csrc/ and runs on the GPU only.
"""
from __future__ import annotations

import numpy as np

__all__ = ["euler_to_quaternion", "quaternion_to_euler", "normalize_quaternion", "to_serializable",
           "object_to_serializable_dict"]


def euler_to_quaternion(roll, pitch, yaw):
    """xyz Euler angles -> [w, x, y, z].  Accepts scalars or equal-length arrays (returns (..., 4))."""
    r, p, y = (np.asarray(v, dtype=np.float64) for v in (roll, pitch, yaw))
    cr, sr = np.cos(r / 2), np.sin(r / 2)
    cp, sp = np.cos(p / 2), np.sin(p / 2)
    cy, sy = np.cos(y / 2), np.sin(y / 2)
    qx = sr * cp * cy - cr * sp * sy
    qy = cr * sp * cy + sr * cp * sy
    qz = cr * cp * sy - sr * sp * cy
    qw = cr * cp * cy + sr * sp * sy
    return np.stack([qw, qx, qy, qz], axis=-1)


def quaternion_to_euler(q):
    """[w, x, y, z] (last axis) -> xyz Euler angles (last axis), out-of-range pitch clamped to +-pi/2."""
    q = np.asarray(q, dtype=np.float64)
    w, x, y, z = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    roll = np.arctan2(2 * (w * x + y * z), 1 - 2 * (x * x + y * y))
    sinp = 2 * (w * y - z * x)
    with np.errstate(invalid="ignore"):
        pitch = np.where(np.abs(sinp) >= 1, np.copysign(np.pi / 2, sinp), np.arcsin(np.clip(sinp, -1, 1)))
    yaw = np.arctan2(2 * (w * z + x * y), 1 - 2 * (y * y + z * z))
    return np.stack([roll, pitch, yaw], axis=-1)


def normalize_quaternion(q):
    q = np.asarray(q, dtype=np.float64)
    n = np.linalg.norm(q)
    return q / n if n > 1e-12 else np.array([1.0, 0.0, 0.0, 0.0])


def to_serializable(obj):
    if isinstance(obj, np.ndarray):
        return obj.tolist()
    if isinstance(obj, (np.floating, np.integer)):
        return obj.item()
    if isinstance(obj, dict):
        return {k: to_serializable(v) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return [to_serializable(v) for v in obj]
    return obj


def object_to_serializable_dict(obj):
    return {k: to_serializable(v) for k, v in vars(obj).items() if not k.startswith("_")}
