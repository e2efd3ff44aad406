"""GPU parity of the GICP path against the CPU oracle (fast_gicp semantics), through the C ABI.

Tolerances: correspondences / NN indices / inlier counts bit-exact; float32 NN distances bit-exact;
float64 covariances, H, b, cost rel <= 1e-9; final pose <= 1e-6 m / 1e-6 rad; same iteration count and convergence flag."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg
from common import rot_err, trans_err, perturb, relrel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(synth):
    tgt = synth.scan("vlp16_16k", frame=0, stride=8)
    src = synth.scan("vlp16_16k", frame=1, stride=8)
    return src, tgt


@pytest.fixture(scope="module")
def reg():
    r = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    yield r
    r.close()


def test_factory_defaults(reg):
    c = reg.config
    assert c.method == pkg.B2R_METHOD_GICP and c.max_iterations == 64 and c.k_correspondences == 20
    assert c.transformation_epsilon == 0.01 and c.max_correspondence_distance == 2.5 and c.rotation_epsilon == 2e-3


def test_nearest_bit_exact(reg, pair, oracle):
    src, tgt = pair
    reg.setInputTarget(tgt)
    idx, d2 = reg.nearestKSearch(src)
    oi, od = oracle.knn(tgt, src, 1)
    assert np.array_equal(idx, oi[:, 0])
    assert np.array_equal(d2, od[:, 0])


def test_covariances(reg, pair, oracle):
    src, tgt = pair
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    for which, cloud in ((0, src), (1, tgt)):
        got = reg.getCovariances(which, cloud.shape[0])
        want = oracle.gicp_covariances(cloud, 20)
        assert np.max(np.abs(got - want)) < 1e-8
        assert np.max(np.abs(got - np.transpose(got, (0, 2, 1)))) == 0.0


@pytest.mark.parametrize("k", [5, 12, 33])
def test_covariances_other_k(pair, oracle, k):
    """reg_correspondence_randomness != 20 takes the shared-memory list kernel (k_knn_cov) instead of the register one"""
    src, _ = pair
    r = pkg.select_registration_method({"registration_method": "FAST_GICP", "reg_correspondence_randomness": k})
    try:
        r.setInputSource(src)
        got = r.getCovariances(0, src.shape[0])
        want = oracle.gicp_covariances(src, k)
        assert np.max(np.abs(got - want)) < 1e-8
    finally:
        r.close()


def test_linearize_and_error(reg, pair, oracle):
    src, tgt = pair
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    scov = oracle.gicp_covariances(src, 20)
    tcov = oracle.gicp_covariances(tgt, 20)
    for seed in (0, 1):
        T = perturb(seed) if seed else np.eye(4)
        H, b, e = reg.gicpLinearizeAt(T)
        o = oracle.gicp_linearize(src, scov, tgt, tcov, T, 2.5)
        corr = reg.getCorrespondences(src.shape[0])
        assert np.array_equal(corr, o["corr"])
        assert (corr >= 0).sum() > 0.5 * src.shape[0]
        assert relrel(H, o["H"]) < 1e-9 and relrel(b, o["b"]) < 1e-9 and abs(e - o["err"]) <= 1e-9 * abs(o["err"])
        T2 = perturb(seed + 10, 0.05, 0.3) @ T
        e2 = reg.gicpErrorAt(T2)
        oe2 = oracle.gicp_error(src, tgt, o["corr"], o["mahal"], T2)
        assert abs(e2 - oe2) <= 1e-9 * abs(oe2)


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_align_matches_oracle(reg, pair, oracle, seed):
    src, tgt = pair
    guess = (perturb(100 + seed, 0.4, 2.5) if seed else np.eye(4)).astype(np.float32)
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    reg.align(guess)
    o = oracle.gicp_align(src, tgt, guess)
    T = reg.getFinalTransformation()
    assert reg.hasConverged() == o["converged"]
    assert reg.nr_iterations == o["iterations"]
    assert trans_err(T, o["T"]) < 1e-6 and rot_err(T, o["T"]) < 1e-6
    assert np.array_equal(reg.getCorrespondences(src.shape[0]), o["corr"])
    # the synthetic pair is one metre apart along the circuit: the estimate must be close to the ground truth motion
    assert 0.5 < np.linalg.norm(T[:3, 3]) < 1.5


def test_fitness_and_inliers(reg, pair, oracle):
    src, tgt = pair
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    reg.align(np.eye(4, dtype=np.float32))
    T = reg.getFinalTransformation()
    for max_range in (np.finfo(np.float64).max, 2.5, 0.01):
        score, used, inl = reg.getFitnessScore(max_range, full=True)
        os_, on, oi = oracle.fitness(tgt, src, T, max_range, 0.25)
        assert used == on and inl == oi
        assert abs(score - os_) <= 1e-12 * max(abs(os_), 1e-300) or score == os_
    aligned = reg.getAligned()
    want = (src[:, :3].astype(np.float32) @ T[:3, :3].T.astype(np.float32))
    assert np.allclose(aligned[:, :3], want + T[:3, 3], atol=1e-4)
    assert np.array_equal(aligned[:, 4:], src[:, 4:])


def test_promote_source_equals_fresh_target(reg, pair, oracle):
    src, tgt = pair
    src2 = tgt  # align frame0 against frame1 after promoting frame1 from source to target
    reg.setInputTarget(tgt)
    reg.setInputSource(src)
    reg.align(np.eye(4, dtype=np.float32))
    reg.promoteSourceToTarget()
    reg.setInputSource(src2)
    reg.align(np.eye(4, dtype=np.float32))
    T_promoted = reg.getFinalTransformation()
    reg.setInputTarget(src)
    reg.setInputSource(src2)
    reg.align(np.eye(4, dtype=np.float32))
    assert np.array_equal(T_promoted, reg.getFinalTransformation())


def test_empty_and_tiny_clouds(reg, pair):
    src, tgt = pair
    reg.setInputTarget(tgt)
    reg.setInputSource(np.zeros((0, 8), np.float32))
    reg.align(np.eye(4, dtype=np.float32))
    assert not reg.hasConverged()
    reg.setInputTarget(np.zeros((0, 8), np.float32))
    reg.setInputSource(src)
    reg.align(np.eye(4, dtype=np.float32))
    assert not reg.hasConverged()
    score = reg.getFitnessScore()
    assert score == np.finfo(np.float64).max
