"""GPU parity of the prefilter chain ("next" rows) against the oracle: kept sets bit-exact, records copied whole,
input order preserved."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reg():
    r = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    yield r
    r.close()


def test_distance_filter(reg, synth, oracle):
    cloud = synth.scan("vlp16", frame=5, stride=8)
    for near, far in ((1.0, 100.0), (0.1, 1000.0), (8.0, 30.0)):
        out = reg.distanceFilter(cloud, near, far)
        keep = oracle.distance_filter(cloud, near, far)
        assert np.array_equal(out, cloud[keep])
    assert 0 < reg.distanceFilter(cloud, 8.0, 30.0).shape[0] < cloud.shape[0]


@pytest.mark.parametrize("radius,min_nb", [(0.5, 2), (0.8, 2), (0.3, 5)])
def test_radius_outlier_removal(reg, synth, oracle, radius, min_nb):
    cloud = reg.voxelGridFilter(synth.scan("vlp16", frame=6, stride=8), 0.1)   # the reference filters the downsampled cloud
    out = reg.radiusOutlierRemoval(cloud, radius, min_nb)
    keep = oracle.radius_outlier(cloud, radius, min_nb)
    assert np.array_equal(out, cloud[keep])
    assert 0 < out.shape[0] < cloud.shape[0]


@pytest.mark.parametrize("mean_k,mul", [(20, 1.0), (10, 0.5)])
def test_statistical_outlier_removal(reg, synth, oracle, mean_k, mul):
    cloud = reg.voxelGridFilter(synth.scan("vlp16_16k", frame=6, stride=8), 0.1)
    out = reg.statisticalOutlierRemoval(cloud, mean_k, mul)
    keep, dist = oracle.statistical_outlier(cloud, mean_k, mul)
    assert np.array_equal(out, cloud[keep])
    assert 0 < out.shape[0] < cloud.shape[0]


def test_prefilter_edge_cases(reg, oracle):
    empty = np.zeros((0, 8), np.float32)
    assert reg.distanceFilter(empty, 1, 100).shape[0] == 0 and reg.radiusOutlierRemoval(empty, 0.5, 2).shape[0] == 0
    few = np.array([[0, 0, 0, 1, 1, 0, 0, 0], [0.1, 0, 0, 1, 2, 0, 0, 0], [5, 5, 5, 1, 3, 0, 0, 0]], np.float32)
    out = reg.radiusOutlierRemoval(few, 0.5, 1)
    assert np.array_equal(out, few[oracle.radius_outlier(few, 0.5, 1)]) and out.shape[0] == 2
    assert reg.radiusOutlierRemoval(few, 0.5, 5).shape[0] == 0   # fewer points than min_neighbors + 1: everything is an outlier
