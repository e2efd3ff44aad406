"""The two reference callers, mirrored above the C ABI: ScanMatchingOdometryNodelet::matching and LoopDetector::matching.
The same sequences are replayed against the oracle with the reference's control flow restated in Python (test-side)."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg
from common import rot_err, trans_err, perturb

pytestmark = pytest.mark.gpu


def quat_w(R):
    return 0.5 * np.sqrt(max(np.float32(R[0, 0] + R[1, 1] + R[2, 2]) + np.float32(1.0), 0))


def test_odometry_sequence_matches_oracle(synth, oracle):
    frames = [synth.scan("vlp16_16k", frame=k, stride=8) for k in range(8)]
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    odo = pkg.ScanMatchingOdometry(reg, keyframe_delta_trans=1.5, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0, publish_status=True)
    # oracle-side replay of apps/scan_matching_odometry_nodelet.cpp:165-262
    keyframe, keyframe_pose, prev_trans = None, np.eye(4, dtype=np.float32), np.eye(4, dtype=np.float32)
    n_keyframes = 0
    for k, cloud in enumerate(frames):
        st = odo.matching(0.1 * k, cloud)
        if keyframe is None:
            keyframe = cloud
            assert st["keyframe_updated"] and np.array_equal(st["odom"], np.eye(4, dtype=np.float32))
            continue
        o = oracle.gicp_align(cloud, keyframe, prev_trans)
        assert st["converged"] == o["converged"] and st["iterations"] == o["iterations"]
        assert trans_err(st["trans"], o["T"]) < 1e-6 and rot_err(st["trans"], o["T"]) < 1e-6
        fs, fn, fi = oracle.fitness(keyframe, cloud, o["T"])
        assert abs(st["matching_error"] - fs) <= 1e-6 * fs
        assert abs(st["inlier_fraction"] - fi / cloud.shape[0]) < 1e-6
        trans = o["T"]
        odom = keyframe_pose @ trans
        assert np.allclose(st["odom"], odom, atol=1e-5)
        prev_trans = trans
        dt = np.linalg.norm(trans[:3, 3])
        da = np.arccos(min(quat_w(trans[:3, :3]), 1.0))
        if dt > 1.5 or da > 1.0:
            keyframe, keyframe_pose, prev_trans = cloud, odom, np.eye(4, dtype=np.float32)
            n_keyframes += 1
            assert st["keyframe_updated"]
        else:
            assert not st["keyframe_updated"]
    assert n_keyframes >= 2
    # accumulated odometry vs ground truth of the synthetic circuit (7 m travelled)
    gt = np.linalg.inv(synth.pose_matrix(0)) @ synth.pose_matrix(7)
    assert trans_err(st["odom"], gt) < 0.3
    odo.close()
    reg.close()


def test_loop_matching_matches_oracle(synth, oracle):
    new_kf = synth.scan("vlp16_16k", frame=0, stride=8)
    cand_frames = [1, 3, 251, 2]  # 251 ~ one lap later (loop), others nearby
    cands = [synth.scan("vlp16_16k", frame=f, stride=8) for f in cand_frames]
    guesses = []
    for i, f in enumerate(cand_frames):
        rel = np.linalg.inv(synth.pose_matrix(0)) @ synth.pose_matrix(f)
        g = (rel @ perturb(40 + i, 0.3, 2.0)).astype(np.float32)
        g[2, 3] = 0.0  # loop_detector.hpp:142
        guesses.append(g)
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    det = pkg.LoopDetector(reg, fitness_score_max_range=2.5, fitness_score_thresh=0.5)
    best, results = det.matching(cands, new_kf, guesses)
    best_score, best_i = np.finfo(np.float64).max, -1
    for i, (c, g) in enumerate(zip(cands, guesses)):
        o = oracle.gicp_align(c, new_kf, g)
        score, _, _ = oracle.fitness(new_kf, c, o["T"], 2.5)
        r = results[i]
        assert r["converged"] == o["converged"] and r["iterations"] == o["iterations"]
        assert trans_err(r["T"], o["T"]) < 1e-6 and rot_err(r["T"], o["T"]) < 1e-6
        assert abs(r["fitness"] - score) <= 1e-6 * score
        if not o["converged"] or score > best_score:
            continue
        best_score, best_i = score, i
    if best_score > 0.5:
        best_i = -1
    assert best == best_i
    assert det.matching([], new_kf, [])[0] == -1
    reg.close()


def test_prefetch_pipeline_gives_identical_odometry(synth):
    """announcing the next scan (software pipelining on a second stream) must not change a single bit of the results"""
    frames = [synth.scan("vlp16_16k", frame=k, stride=8) for k in range(7)]
    outs = []
    for use_prefetch in (False, True):
        reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
        odo = pkg.ScanMatchingOdometry(reg, keyframe_delta_trans=1.5, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0)
        res = []
        for k, cloud in enumerate(frames):
            if use_prefetch and k + 1 < len(frames):
                nxt = frames[k + 1]
                odo.prefetch_raw(nxt.ctypes.data, nxt.shape[0], nxt.shape[1] * 4)
            st = odo.matching(0.1 * k, cloud)
            res.append((st["odom"].copy(), st["iterations"], st["converged"], st["keyframe_updated"]))
        outs.append(res)
        odo.close()
        reg.close()
    for a, b in zip(*outs):
        assert np.array_equal(a[0], b[0]) and a[1:] == b[1:]
