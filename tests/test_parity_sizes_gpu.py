"""GPU parity AT THE SIZES THE BENCH QUOTES (BASELINE.json configs[1..4]): 65 536-point VLP-16 scans for GICP, 131 072-point
HDL-32e scans for NDT, KITTI-shape 120 000-point scans for voxel grid -> GICP.  Same tolerances as the 16k tests
(DESIGN.md §2): correspondences / counts bit-exact, float64 quantities rel <= 1e-9, poses <= 1e-6 (GICP) / 1e-4 (NDT).
At 64k points the BVH has 64 super-nodes (pass 2 of bvh_group_search loops twice), at 128k 128."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg
from common import rot_err, trans_err, perturb, relrel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair64k(synth):
    return synth.scan("vlp16", frame=1, stride=8), synth.scan("vlp16", frame=0, stride=8)


@pytest.fixture(scope="module")
def reg():
    r = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    yield r
    r.close()


def test_gicp_64k_covariances_linearize_and_correspondences(reg, pair64k, oracle):
    src, tgt = pair64k
    assert src.shape[0] == 65536
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    scov = oracle.gicp_covariances(src, 20)
    tcov = oracle.gicp_covariances(tgt, 20)
    assert np.max(np.abs(reg.getCovariances(0, src.shape[0]) - scov)) < 1e-8
    assert np.max(np.abs(reg.getCovariances(1, tgt.shape[0]) - tcov)) < 1e-8
    for T in (np.eye(4), perturb(1, 0.3, 2.0)):
        H, b, e = reg.gicpLinearizeAt(T)
        o = oracle.gicp_linearize(src, scov, tgt, tcov, T, 2.5)
        assert np.array_equal(reg.getCorrespondences(src.shape[0]), o["corr"])  # seeded-by-firing-order 1-NN, 64 super-nodes
        assert relrel(H, o["H"]) < 1e-9 and relrel(b, o["b"]) < 1e-9 and abs(e - o["err"]) <= 1e-9 * abs(o["err"])


def test_gicp_64k_full_align_and_fitness(reg, pair64k, oracle):
    src, tgt = pair64k
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    for guess in (np.eye(4, dtype=np.float32), perturb(2, 0.3, 2.0).astype(np.float32)):
        reg.align(guess)
        o = oracle.gicp_align(src, tgt, guess)
        assert reg.hasConverged() == o["converged"] and reg.nr_iterations == o["iterations"]
        assert np.array_equal(reg.getCorrespondences(src.shape[0]), o["corr"])  # correspondences of the last (seeded) linearisation
        T = reg.getFinalTransformation()
        assert trans_err(T, o["T"]) < 1e-6 and rot_err(T, o["T"]) < 1e-6
        fs, fn, fi = oracle.fitness(tgt, src, o["T"], 2.5)
        s, used, inl = reg.getFitnessScore(2.5, T=o["T"], full=True)  # explicit T (information_matrix_calculator.cpp:49-80 shape)
        assert used == fn and inl == fi and abs(s - fs) <= 1e-12 * fs


def test_gicp_64k_odometry_chain_with_keyframe_switch(synth, oracle):
    frames = [synth.scan("vlp16", frame=k, stride=8) for k in range(6)]
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    odo = pkg.ScanMatchingOdometry(reg, keyframe_delta_trans=1.5, keyframe_delta_angle=1.0, keyframe_delta_time=10000.0)
    keyframe, prev, switches = None, np.eye(4, dtype=np.float32), 0
    for k, cloud in enumerate(frames):
        st = odo.matching(0.1 * k, cloud)
        if keyframe is None:
            keyframe = cloud
            continue
        o = oracle.gicp_align(cloud, keyframe, prev)
        assert st["converged"] == o["converged"] and st["iterations"] == o["iterations"]
        assert trans_err(st["trans"], o["T"]) < 1e-6 and rot_err(st["trans"], o["T"]) < 1e-6
        if not st["keyframe_updated"]:  # a keyframe switch promotes the source to target: its correspondences are gone
            assert np.array_equal(reg.getCorrespondences(cloud.shape[0]), o["corr"])
        prev = o["T"]
        if np.linalg.norm(prev[:3, 3]) > 1.5:
            keyframe, prev = cloud, np.eye(4, dtype=np.float32)
            switches += 1
            assert st["keyframe_updated"]
    assert switches >= 1
    odo.close()
    reg.close()


def test_gicp_64k_batch_matches_oracle(synth, oracle):
    """configs[3] shape: candidates one lap later against a shared new keyframe, through the batched path"""
    tf = 0
    tgt = synth.scan("vlp16", frame=tf, stride=8)
    lb = pkg.RegistrationBatch(params={"registration_method": "FAST_GICP"})
    it = lb.addCloud(tgt)
    pairs, clouds = [], []
    for j, sf in enumerate((251, 249)):
        c = synth.scan("vlp16", frame=sf, stride=8)
        g = (np.linalg.inv(synth.pose_matrix(tf)) @ synth.pose_matrix(sf) @ perturb(80 + j, 0.5, 3.0)).astype(np.float32)
        g[2, 3] = 0.0
        clouds.append(c)
        pairs.append((lb.addCloud(c), it, g))
    res = lb.align(pairs, True, 2.5)
    for (cid, _, g), c, q in zip(pairs, clouds, res):
        o = oracle.gicp_align(c, tgt, g)
        assert q["converged"] == o["converged"] and q["iterations"] == o["iterations"]
        assert trans_err(q["T"], o["T"]) < 1e-6 and rot_err(q["T"], o["T"]) < 1e-6
        score, _, _ = oracle.fitness(tgt, c, o["T"], 2.5)
        assert abs(q["fitness"] - score) <= 1e-6 * score
    lb.close()


def test_explicit_T_fitness_matches_calc_fitness_score(reg, synth, oracle):
    """InformationMatrixCalculator::calc_fitness_score(cloud1, cloud2, relpose, max_range)
    (src/hdl_graph_slam/information_matrix_calculator.cpp:49-80): kd-tree on cloud1, cloud2 transformed by relpose (float),
    mean of squared NN distances <= max_range.  Served by b2r_fitness with an explicit (column-major) T, non-symmetric T."""
    c1 = synth.scan("vlp16_16k", frame=0, stride=8)
    c2 = synth.scan("vlp16_16k", frame=2, stride=8)
    rel = (np.linalg.inv(synth.pose_matrix(0)) @ synth.pose_matrix(2) @ perturb(9, 0.1, 5.0)).astype(np.float32)
    reg.setInputTarget(c1)
    reg.setInputSource(c2)
    for max_range in (np.finfo(np.float64).max, 2.0, 0.05):
        fs, fn, fi = oracle.fitness(c1, c2, rel, max_range)
        s, used, inl = reg.getFitnessScore(max_range, T=rel, full=True)
        assert used == fn and inl == fi
        assert abs(s - fs) <= 1e-12 * abs(fs)
    # a transposed T would give another answer: guards the column-major -> row-major branch of b2r_fitness
    s_wrong = reg.getFitnessScore(2.0, T=rel.T)
    assert abs(s_wrong - oracle.fitness(c1, c2, rel, 2.0)[0]) > 1e-6


def test_ndt_128k_voxel_map_derivatives_and_fixed_30_iterations(synth, oracle):
    tgt = synth.scan("hdl32e", frame=0, stride=8)
    src = synth.scan("hdl32e", frame=1, stride=8)
    assert src.shape[0] == 131072
    cfg = pkg.default_config(pkg.B2R_METHOD_NDT)
    cfg.ndt_resolution = 1.0
    cfg.ndt_fixed_iterations = 30
    reg = pkg.Registration(cfg)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    m = oracle.NdtMap(tgt, 1.0)
    want = m.dump()
    got = reg.ndtGetVoxels()
    assert np.array_equal(got["keys"], want["keys"]) and np.array_equal(got["npts"], want["npts"])
    valid = want["npts"] >= 6
    assert np.array_equal(got["mean"][valid], want["mean"][valid])
    assert relrel(got["icov"][valid], want["icov"][valid]) < 1e-9
    for p in (np.zeros(6), np.array([0.3, -0.2, 0.05, 0.01, -0.02, 0.03])):
        score, g, H, npairs = reg.ndtDerivativesAt(p)
        o = m.derivatives(src, p)
        assert npairs == o["n_pairs"]
        assert abs(score - o["score"]) <= 1e-9 * abs(o["score"]) and relrel(g, o["g"]) < 1e-9 and relrel(H, o["H"]) < 1e-9
    reg.align(np.eye(4, dtype=np.float32))
    o = m.align(src, np.eye(4, dtype=np.float32), fixed_iterations=30)
    assert reg.nr_iterations == o["iterations"] == 30
    T = reg.getFinalTransformation()
    assert trans_err(T, o["T"]) < 1e-4 and rot_err(T, o["T"]) < 1e-4
    reg.close()


def test_kitti_120k_voxelgrid_then_gicp(synth, oracle):
    """configs[4] front end: 120 000-point scan -> voxel grid 0.25 m (launch/hdl_graph_slam_kitti.launch:28) -> GICP"""
    raw0 = synth.scan("kitti", frame=0, stride=8)
    raw1 = synth.scan("kitti", frame=1, stride=8)
    assert raw0.shape[0] == 120000
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    ds = []
    for raw in (raw0, raw1):
        out, keys, counts, rc = reg.voxelGridFilter(raw, 0.25, with_keys=True)
        oo, ok, oc, orc_ = oracle.voxelgrid(raw, 0.25)
        assert rc == orc_ and np.array_equal(keys, ok) and np.array_equal(counts, oc)
        assert np.array_equal(out[:, :3], oo[:, :3]) and np.array_equal(out[:, 4], oo[:, 3])  # centroid xyz + averaged intensity (PointXYZI offset 16)
        ds.append(out)
    reg.setInputTarget(ds[0])
    reg.setInputSource(ds[1])
    reg.align(np.eye(4, dtype=np.float32))
    o = oracle.gicp_align(ds[1], ds[0], np.eye(4, dtype=np.float32))
    assert reg.hasConverged() == o["converged"] and reg.nr_iterations == o["iterations"]
    assert np.array_equal(reg.getCorrespondences(ds[1].shape[0]), o["corr"])
    T = reg.getFinalTransformation()
    assert trans_err(T, o["T"]) < 1e-6 and rot_err(T, o["T"]) < 1e-6
    reg.close()
