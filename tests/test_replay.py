"""Config C5: the scan-to-scan odometry front end (go-rio_b200/replay.py mirrors
ScanMatchingOdometryNodelet::matching + KeyframeUpdater::decide). CPU: the call
pattern through the oracle; GPU: the same sequence through the CUDA library gives
the same poses and the same keyframe decisions."""
import importlib

import numpy as np
import pytest

from helpers import DEPLOYED
from oracle_binding import Oracle

replay = importlib.import_module("go-rio_b200.replay")


def test_keyframe_rule():
    """keyframe_updater.hpp:38-63: first frame always; then translation >= 0.5 m or rotation >= 10 deg"""
    u = replay.KeyframeUpdater(0.5, 0.1745)
    P = np.eye(4)
    assert u.decide(P)
    P2 = P.copy(); P2[0, 3] = 0.3
    assert not u.decide(P2)
    P3 = P.copy(); P3[0, 3] = 0.6
    assert u.decide(P3) and abs(u.accum_distance - 0.6) < 1e-12
    c, s = np.cos(0.2), np.sin(0.2)
    P4 = P3.copy(); P4[:2, :2] = [[c, -s], [s, c]]
    assert u.decide(P4)


def test_replay_with_oracle(synth):
    frames = list(synth.drive_frames(5001, 12, 600))
    o = Oracle(search=1)
    o.set_params(**DEPLOYED)
    r = replay.replay(o, frames)
    assert r["poses"].shape == (12, 4, 4) and r["n_keyframes"] >= 3 and r["n_not_converged"] == 0
    assert np.array_equal(r["poses"][0], np.eye(4))
    assert r["final_drift_m"] < 0.5 * r["path_m"]
    o2 = Oracle(search=1)
    o2.set_params(**DEPLOYED)
    assert np.array_equal(replay.replay(o2, frames)["poses"], r["poses"])


@pytest.mark.gpu
def test_replay_gpu_matches_oracle(gorio, synth):
    frames = list(synth.drive_frames(5002, 30, 1000))
    o = Oracle(search=1)
    g = gorio.FastAPDGICP(0)
    for reg in (o, g):
        reg.set_params(**DEPLOYED, maha_fp64=1)
    ro, rg = replay.replay(o, frames), replay.replay(g, frames)
    assert rg["n_keyframes"] == ro["n_keyframes"] and rg["iterations"] == ro["iterations"]
    assert np.abs(rg["poses"] - ro["poses"]).max() < 1e-5  # float final_transformation_ composed over the keyframes
