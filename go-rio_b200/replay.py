"""Odometry replay (config C5): the scan-to-scan front end of Go-RIO with the
registration swapped for the B200 FastAPDGICP.

Mirrors ScanMatchingOdometryNodelet::matching
(4DRadarSLAM/apps/scan_matching_odometry_nodelet.cpp:423-632) for the deployed
configuration of launch/ntu_loop2.launch (enable_scan_to_map = false, use_ego_vel =
false, enable_imu_thresholding = false, enable_transform_thresholding = true) and
KeyframeUpdater::decide (include/radar_graph_slam/keyframe_updater.hpp:38-63).
ROS, IMU fusion and the pose graph are out of scope; this is only the call pattern
the registration sees: setInputTarget(keyframe) ... setInputSource(frame),
align(guess = previous transform), hasConverged, getFinalTransformation, and on a
new keyframe setInputTarget(frame).

`reg` is any object with the Python mirror API of go-rio_b200._binding.Registration
(the CUDA library or, in tests, the CPU oracle).
"""
import numpy as np


def _angle(R):
    return float(np.arccos(np.clip((np.trace(R) - 1.0) / 2.0, -1.0, 1.0)))


class KeyframeUpdater:
    """keyframe_updater.hpp:38-63"""

    def __init__(self, delta_trans=0.5, delta_angle=0.1745):
        self.delta_trans, self.delta_angle = delta_trans, delta_angle
        self.is_first = True
        self.prev_keypose = np.eye(4)
        self.accum_distance = 0.0

    def decide(self, pose):
        if self.is_first:
            self.is_first = False
            self.prev_keypose = pose.copy()
            return True
        delta = np.linalg.inv(self.prev_keypose) @ pose
        dx = float(np.linalg.norm(delta[:3, 3]))
        da = _angle(delta[:3, :3])
        if dx < self.delta_trans and da < self.delta_angle:
            return False
        self.accum_distance += dx
        self.prev_keypose = pose.copy()
        return True


class ScanMatchingOdometry:
    def __init__(self, reg, keyframe_delta_trans=0.5, keyframe_delta_angle=0.1745, max_acceptable_trans=5.0,
                 max_acceptable_angle=3.0, enable_transform_thresholding=True):
        self.reg = reg
        self.updater = KeyframeUpdater(keyframe_delta_trans, keyframe_delta_angle)
        self.max_acceptable_trans = max_acceptable_trans
        self.max_acceptable_angle = max_acceptable_angle  # the reference compares radians with this value as given (:507)
        self.enable_transform_thresholding = enable_transform_thresholding
        self.keyframe_cloud = None
        self.keyframe_pose = np.eye(4)
        self.prev_trans = np.eye(4)
        self.n_keyframes = 0
        self.n_not_converged = 0
        self.n_thresholded = 0
        self.iterations = []
        self.frame_id = 0

    def matching(self, cloud):
        """returns the odometry pose of this frame (:423-632)"""
        # The cache key stands in for the reference's shared_ptr identity (fast_apdgicp_impl.hpp:116,:128): a frame promoted
        # to keyframe below is recognised by it and its grid / covariances are reused, not rebuilt. A frame counter, not the
        # array's address: a streamed sequence frees each frame, and the allocator hands the same address to the next one.
        self.frame_id += 1
        key = self.frame_id
        if self.keyframe_cloud is None:  # :424-438
            self.prev_trans = np.eye(4)
            self.keyframe_pose = np.eye(4)
            self.keyframe_cloud = cloud
            self.reg.set_input_target(cloud, key=key)
            return np.eye(4)
        self.reg.set_input_source(cloud, key=key)  # :442
        guess = self.prev_trans.astype(np.float32)  # :461 (use_ego_vel = false, msf_delta = I)
        r = self.reg.align(guess)  # :465
        self.iterations.append(r["iterations"])
        if not r["converged"]:  # :473-478
            self.n_not_converged += 1
            return self.keyframe_pose @ self.prev_trans
        trans = r["T"].astype(np.float64)  # :479
        odom = self.keyframe_pose @ trans  # :480
        thresholded = False
        if self.enable_transform_thresholding:  # :496-570 with enable_imu_thresholding = false
            radar_delta = np.linalg.inv(self.prev_trans) @ trans
            dx = float(np.linalg.norm(radar_delta[:3, 3]))
            da = _angle(radar_delta[:3, :3])
            if dx > self.max_acceptable_trans or da > self.max_acceptable_angle:
                self.prev_trans = trans
                thresholded = True
                self.n_thresholded += 1
                odom = self.keyframe_pose @ self.prev_trans @ radar_delta
        if not thresholded:  # :578-581
            self.prev_trans = trans
        if self.updater.decide(odom):  # :583-600
            self.keyframe_cloud = cloud
            self.reg.set_input_target(cloud, key=key)
            self.keyframe_pose = odom
            self.prev_trans = np.eye(4)
            self.n_keyframes += 1
        return odom


def replay(reg, frames):
    """runs the front end over an iterable of (index, cloud, gt_pose); returns poses and counters"""
    odo = ScanMatchingOdometry(reg)
    poses, gts = [], []
    for _, cloud, gt in frames:
        poses.append(odo.matching(np.ascontiguousarray(cloud)))
        gts.append(gt)
    poses, gts = np.array(poses), np.array(gts)
    rel_gt = np.linalg.inv(gts[0]) @ gts
    err = np.linalg.norm(poses[:, :3, 3] - rel_gt[:, :3, 3], axis=1)
    return {"poses": poses, "gt": rel_gt, "n_keyframes": odo.n_keyframes, "n_not_converged": odo.n_not_converged,
            "n_thresholded": odo.n_thresholded, "iterations": odo.iterations, "final_drift_m": float(err[-1]),
            "path_m": float(np.linalg.norm(np.diff(rel_gt[:, :3, 3], axis=0), axis=1).sum())}
