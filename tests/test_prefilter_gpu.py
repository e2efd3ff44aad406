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


def test_deskew_matches_oracle(reg, synth, oracle):
    """PrefilteringNodelet::deskewing (apps/prefiltering_nodelet.cpp:182-243): bit-exact float32 quaternion arithmetic"""
    cloud = synth.scan("vlp16", frame=4, stride=8)
    for w in ((0.0, 0.0, 0.0), (0.02, -0.01, 0.35), (1.5, 2.0, -3.0)):
        got = reg.deskew(cloud, 0.1, w)
        want = oracle.deskew(cloud, 0.1, w)
        assert np.array_equal(got, want)
    assert np.array_equal(reg.deskew(cloud, 0.1, (0, 0, 0))[:, :3], cloud[:, :3])  # no rotation: points unchanged
    assert np.max(np.abs(reg.deskew(cloud, 0.1, (0, 0, 0.5))[:, :3] - cloud[:, :3])) > 1e-3


@pytest.mark.parametrize("outlier", ["STATISTICAL", "RADIUS", "NONE"])
def test_prefilter_chain_equals_stagewise_oracle(reg, synth, oracle, outlier):
    """b2r_prefilter = PrefilteringNodelet::cloud_callback (apps/prefiltering_nodelet.cpp:106-136) with the cloud resident in HBM between
    the stages: deskew -> distance filter -> voxel grid -> outlier removal == the oracle's stages composed on the host"""
    raw = synth.scan("vlp16", frame=9, stride=8)
    w = (0.01, -0.02, 0.3)
    method = {"NONE": 0, "STATISTICAL": 1, "RADIUS": 2}[outlier]
    out, dptr, m = reg.prefilter(raw, deskewing=1, scan_period=0.1, angular_velocity=w, distance_near_thresh=1.0, distance_far_thresh=60.0,
                                 downsample_resolution=0.1, outlier_removal_method=method, radius_radius=0.5, radius_min_neighbors=2)
    c = oracle.deskew(raw, 0.1, w)
    c = c[oracle.distance_filter(c, 1.0, 60.0)]
    xyzi, _, _, rc = oracle.voxelgrid(c, 0.1)
    ds = np.zeros((xyzi.shape[0], 8), np.float32)
    ds[:, :3], ds[:, 3], ds[:, 4] = xyzi[:, :3], 1.0, xyzi[:, 3]
    if outlier == "STATISTICAL":
        ds = ds[oracle.statistical_outlier(ds, 20, 1.0)[0]]
    elif outlier == "RADIUS":
        ds = ds[oracle.radius_outlier(ds, 0.5, 2)]
    assert m == ds.shape[0] and dptr
    assert np.array_equal(out, ds)
    # the device-resident result feeds the registration directly
    reg.setInputTargetDevice(dptr, m, raw.shape[1] * 4)
    idx, d2 = reg.nearestKSearch(ds[:100])
    assert np.array_equal(idx, np.arange(100)) and np.all(d2 == 0)


def test_statistical_outlier_keeps_nonfinite_points_like_pcl(reg, synth, oracle):
    cloud = reg.voxelGridFilter(synth.scan("vlp16_16k", frame=6, stride=8), 0.1)[:3000].copy()
    cloud[10, 0] = np.nan
    cloud[20, 2] = np.inf
    out = reg.statisticalOutlierRemoval(cloud, 20, 1.0)
    keep, _ = oracle.statistical_outlier(cloud, 20, 1.0)
    assert keep[10] and keep[20]  # pcl: a non-finite point has distance 0 and passes the threshold
    assert np.array_equal(np.nan_to_num(out, nan=-7.0, posinf=-8.0), np.nan_to_num(cloud[keep], nan=-7.0, posinf=-8.0))


def test_kitti_chain_in_hbm_equals_the_host_hand_off(synth, oracle):
    """BASELINE configs[4], per-scan chain (bench.py --workload kitti_pipeline): raw 120k-point scan -> b2r_prefilter (distance, voxel grid
    0.25, radius outlier removal) -> b2r_odometry_matching_device on the filtered cloud WHERE IT LIES in HBM, against the reference's hand-off
    (filtered cloud back on the host, then matching): same filtered clouds (vs the stage-wise oracle), same poses, same keyframe decisions."""
    import torch
    pre = dict(use_distance_filter=1, distance_near_thresh=0.1, distance_far_thresh=100.0, downsample_method=1, downsample_resolution=0.25,
               outlier_removal_method=2, radius_radius=0.5, radius_min_neighbors=2)
    par = {"registration_method": "FAST_GICP", "reg_transformation_epsilon": 0.1, "reg_max_correspondence_distance": 2.0}
    raws = [synth.scan("kitti", frame=f, stride=8) for f in range(4)]
    # host hand-off
    reg = pkg.select_registration_method(dict(par))
    odo = pkg.ScanMatchingOdometry(reg, keyframe_delta_trans=5.0, keyframe_delta_angle=2.0, keyframe_delta_time=10000.0)
    host_out, host_clouds = [], []
    for k, raw in enumerate(raws):
        filt, _, m = reg.prefilter(raw, **pre)
        host_clouds.append(filt.copy())
        host_out.append(odo.matching(0.1 * k, filt))
    odo.close()
    reg.close()
    # stage-wise oracle of the first scan
    c = raws[0][oracle.distance_filter(raws[0], 0.1, 100.0)]
    v = oracle.voxelgrid(c, 0.25)[0]
    full = np.zeros((v.shape[0], 8), np.float32)
    full[:, :3], full[:, 3], full[:, 4] = v[:, :3], 1.0, v[:, 3]
    full = full[oracle.radius_outlier(full, 0.5, 2)]
    assert np.array_equal(host_clouds[0][:, :3], full[:, :3]) and np.array_equal(host_clouds[0][:, 4], full[:, 4])
    # device chain
    reg = pkg.select_registration_method(dict(par))
    odo = pkg.ScanMatchingOdometry(reg, keyframe_delta_trans=5.0, keyframe_delta_angle=2.0, keyframe_delta_time=10000.0)
    for k, raw in enumerate(raws):
        d = torch.from_numpy(raw).cuda()
        torch.cuda.synchronize()
        _, dptr, m = reg.prefilter_raw(d.data_ptr(), d.shape[0], d.shape[1] * 4, device=True, **pre)
        assert m == host_clouds[k].shape[0]
        st = odo.matching_raw(0.1 * k, dptr, m, raw.shape[1] * 4, device=True)
        assert np.array_equal(st["odom"], host_out[k]["odom"]) and st["iterations"] == host_out[k]["iterations"]
        assert st["keyframe_updated"] == host_out[k]["keyframe_updated"] and st["converged"] == host_out[k]["converged"]
    odo.close()
    reg.close()
