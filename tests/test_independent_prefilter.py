"""CPU-only: an INDEPENDENT second derivation of the oracle's prefilter-chain and voxel-grid functions (rows a5 / f-2 of SURVEY.md §8), written
from the in-tree reference code where it is in the tree (apps/prefiltering_nodelet.cpp: distance_filter :164-180, deskewing :182-243) and from
the published PCL algorithms where it is not (VoxelGrid, RadiusOutlierRemoval, StatisticalOutlierRemoval), with numpy / scipy only — no code
shared with oracle/ or the engine.  The GPU path is checked against the oracle elsewhere (tests/test_prefilter_gpu.py, tests/test_voxelgrid_gpu.py);
this file is what keeps the oracle itself honest for these rows."""
import numpy as np
import pytest
from scipy.spatial import cKDTree


@pytest.fixture(scope="module")
def cloud(synth):
    c = synth.scan("vlp16_16k", frame=3, stride=8).copy()
    rng = np.random.default_rng(9)
    c[::997, 0] = np.nan          # a few non-finite points, as a real driver produces
    c[5::1013, 2] = np.inf
    c[:, 4] = rng.uniform(0, 255, c.shape[0]).astype(np.float32)
    return c


def test_distance_filter_is_the_in_tree_lambda(oracle, cloud):
    """prefiltering_nodelet.cpp:168-171: d = p.getVector3fMap().norm() (float), kept iff d > near && d < far"""
    xyz = cloud[:, :3]
    with np.errstate(invalid="ignore", over="ignore"):
        d = np.sqrt((xyz[:, 0] * xyz[:, 0] + xyz[:, 1] * xyz[:, 1]).astype(np.float32) + (xyz[:, 2] * xyz[:, 2]).astype(np.float32)).astype(np.float64)
    for near, far in ((1.0, 100.0), (0.1, 50.0), (5.0, 20.0)):
        want = (d > near) & (d < far)  # NaN / inf compare false
        got = oracle.distance_filter(cloud, near, far)
        assert np.array_equal(got, want)
    # the thresholds themselves are excluded on both sides
    edge = np.zeros((3, 8), np.float32)
    edge[0, 0], edge[1, 0], edge[2, 0] = 1.0, 2.0, 3.0
    assert list(oracle.distance_filter(edge, 1.0, 3.0)) == [False, True, False]


def test_deskew_is_the_in_tree_loop(oracle, cloud):
    """prefiltering_nodelet.cpp:216-240: ang_v *= -1; delta_t = scan_period * i / size; delta_q = Quaternionf(1, dt/2 w); pt' = delta_q.inverse() * pt
    (Eigen: inverse = conjugate / squaredNorm — NOT normalised —, q * v = v + w (2 u x v) + u x (2 u x v))"""
    c = cloud[np.isfinite(cloud[:, :3]).all(axis=1)][:6000].copy()
    n = c.shape[0]
    w_imu = np.array([0.3, -0.2, 1.1], np.float32)
    scan_period = 0.1
    ang_v = (-w_imu).astype(np.float32)
    dt = scan_period * np.arange(n, dtype=np.float64) / n
    q = np.stack([np.ones(n, np.float32)] + [(dt / 2.0 * float(ang_v[k])).astype(np.float32) for k in range(3)], axis=1)  # w x y z, float32
    n2 = (q.astype(np.float64) ** 2).sum(axis=1)
    inv = np.concatenate([q[:, :1], -q[:, 1:]], axis=1).astype(np.float64) / n2[:, None]
    u, w = inv[:, 1:], inv[:, 0]
    v = c[:, :3].astype(np.float64)
    uv = 2.0 * np.cross(u, v)
    want = v + w[:, None] * uv + np.cross(u, uv)
    got = oracle.deskew(c, scan_period, w_imu)
    assert np.array_equal(got[:, 3:], c[:, 3:])          # only xyz change (:238-239)
    assert np.max(np.abs(got[:, :3] - want)) < 2e-5       # float32 evaluation vs float64 restatement at ranges up to 100 m
    assert np.array_equal(got[0, :3], c[0, :3])           # i = 0: identity quaternion
    # the correction is a rotation about -w by about |w| dt (small angle, slightly scaled because the quaternion is not normalised)
    ang = np.linalg.norm(w_imu) * dt[-1]
    rot = np.linalg.norm(got[-1, :3] - c[-1, :3]) / np.linalg.norm(np.cross(w_imu / np.linalg.norm(w_imu), c[-1, :3]))
    assert abs(rot - ang) < 0.1 * ang


@pytest.mark.parametrize("leaf", [0.1, 0.25, 1.0])
def test_voxelgrid_independent(oracle, cloud, leaf):
    """pcl::VoxelGrid::applyFilter: finite points only; min_b = floor(min * inv_leaf), ijk = floor(p * inv_leaf) - min_b, key = i + j div_x + k div_x div_y,
    one centroid (xyz and intensity) per occupied voxel, output in ascending key order"""
    fin = np.isfinite(cloud[:, :3]).all(axis=1)
    p = cloud[fin]
    inv = np.float32(1.0) / np.float32(leaf)
    lo = np.floor(p[:, :3].min(axis=0) * inv).astype(np.int64)
    hi = np.floor(p[:, :3].max(axis=0) * inv).astype(np.int64)
    div = hi - lo + 1
    ijk = np.floor(p[:, :3] * inv).astype(np.int64) - lo
    key = ijk[:, 0] + ijk[:, 1] * div[0] + ijk[:, 2] * div[0] * div[1]
    uk, inverse, counts = np.unique(key, return_inverse=True, return_counts=True)
    cent = np.zeros((uk.size, 4), np.float64)
    np.add.at(cent, inverse, np.concatenate([p[:, :3], p[:, 4:5]], axis=1).astype(np.float64))
    cent /= counts[:, None]
    out, keys, cnt, rc = oracle.voxelgrid(cloud, leaf)
    assert rc == 0
    assert np.array_equal(keys, uk) and np.array_equal(cnt, counts)
    assert np.max(np.abs(out.astype(np.float64) - cent)) < 2e-4   # float32 running sums of up to ~100 points at 100 m vs float64 means
    assert out.shape[0] < p.shape[0] or leaf < 0.2


@pytest.mark.parametrize("radius,min_neighbors", [(0.5, 2), (1.0, 5), (0.3, 1)])
def test_radius_outlier_independent(oracle, cloud, radius, min_neighbors):
    """pcl::RadiusOutlierRemoval (dense path): a point stays iff its (min_neighbors + 1)-th nearest neighbour — itself included — lies within the radius,
    i.e. iff at least min_neighbors OTHER points are within the radius"""
    fin = np.isfinite(cloud[:, :3]).all(axis=1)
    pts = cloud[fin, :3].astype(np.float64)
    tree = cKDTree(pts)
    cnt = np.array([len(x) for x in tree.query_ball_point(pts, radius)]) - 1   # the point itself does not count
    want = np.zeros(cloud.shape[0], bool)
    want[fin] = cnt >= min_neighbors
    got = oracle.radius_outlier(cloud, radius, min_neighbors)
    # float32 squared distances vs a float64 tree: a pair exactly at the radius may fall on either side
    dd, _ = tree.query(pts, k=min_neighbors + 1)
    unsure = np.zeros(cloud.shape[0], bool)
    unsure[fin] = np.abs(dd[:, -1] - radius) < 1e-5 * radius
    assert np.array_equal(got[~unsure], want[~unsure])
    assert want.sum() > 100 and (~want[fin]).sum() > 10


@pytest.mark.parametrize("mean_k,stddev_mul", [(20, 1.0), (8, 2.0)])
def test_statistical_outlier_independent(oracle, cloud, mean_k, stddev_mul):
    """pcl::StatisticalOutlierRemoval: mean distance to the mean_k nearest OTHER points; threshold = mean + mul * stddev over the finite points with the
    (N - 1) variance; a point stays iff its mean distance is not above the threshold"""
    fin = np.isfinite(cloud[:, :3]).all(axis=1)
    pts = cloud[fin, :3].astype(np.float64)
    dd, _ = cKDTree(pts).query(pts, k=mean_k + 1)
    md = dd[:, 1:].mean(axis=1)
    thresh = md.mean() + stddev_mul * md.std(ddof=1)
    got, dist = oracle.statistical_outlier(cloud, mean_k, stddev_mul)
    assert np.max(np.abs(dist[fin] - md)) < 1e-4
    sure = np.abs(md - thresh) > 1e-4
    assert np.array_equal(got[fin][sure], (md <= thresh)[sure])
    assert got[fin].sum() > 0.8 * fin.sum() and (~got[fin]).sum() > 10
