"""Shared helpers of the parity tests."""
import numpy as np

DEPLOYED = dict(max_correspondence_distance=2.0, transformation_epsilon=0.1)  # launch/ntu_loop2.launch:88-99


def rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(1e-300, np.abs(b).max()))


def pose_err(Ta, Tb):
    """translation (m) and rotation (rad) distance between two 4x4 poses"""
    E = np.asarray(Ta, dtype=np.float64) @ np.linalg.inv(np.asarray(Tb, dtype=np.float64))
    c = np.clip((np.trace(E[:3, :3]) - 1.0) / 2.0, -1.0, 1.0)
    return float(np.linalg.norm(E[:3, 3])), float(np.arccos(c))


def moved_copy(synth, cloud, T):
    """cloud expressed in a frame moved by T: applying T to the copy gives back `cloud`."""
    Ti = np.linalg.inv(T)
    out = cloud.copy()
    out[:, :3] = (cloud[:, :3].astype(np.float64) @ Ti[:3, :3].T + Ti[:3, 3]).astype(np.float32)
    return out
