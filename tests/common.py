"""shared helpers for the parity tests"""
import numpy as np


def rot_err(Ta, Tb):
    R = Ta[:3, :3].astype(np.float64).T @ Tb[:3, :3].astype(np.float64)
    # sin(angle) from the skew part: well conditioned near zero (arccos of the trace is not)
    v = 0.5 * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    s = float(np.linalg.norm(v))
    c = float(np.clip((np.trace(R) - 1) / 2, -1, 1))
    return float(np.arctan2(s, c)) if s > 1e-3 else s


def trans_err(Ta, Tb):
    return float(np.linalg.norm(Ta[:3, 3].astype(np.float64) - Tb[:3, 3].astype(np.float64)))


def perturb(seed, dt=0.3, dr_deg=2.0):
    rng = np.random.default_rng(seed)
    t = rng.uniform(-dt, dt, 3)
    a = np.deg2rad(rng.uniform(-dr_deg, dr_deg, 3))
    cx, sx, cy, sy, cz, sz = np.cos(a[0]), np.sin(a[0]), np.cos(a[1]), np.sin(a[1]), np.cos(a[2]), np.sin(a[2])
    Rx = np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rz = np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]])
    T = np.eye(4)
    T[:3, :3] = Rz @ Ry @ Rx
    T[:3, 3] = t
    return T


def relrel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))
