"""GPU parity of the voxel-grid downsample (pcl::VoxelGrid semantics): keys, counts, ordering and centroids
bit-exact against the oracle (the oracle pins the within-voxel summation order to ascending point index)."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def reg():
    r = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    yield r
    r.close()


@pytest.mark.parametrize("sensor,leaf", [("vlp16_16k", 0.1), ("vlp16", 0.1), ("kitti", 0.25), ("vlp16_16k", 1.0)])
def test_voxelgrid_matches_oracle(reg, synth, oracle, sensor, leaf):
    cloud = synth.scan(sensor, frame=3, stride=8)
    out, keys, counts, rc = reg.voxelGridFilter(cloud, leaf, with_keys=True)
    o_xyzi, o_keys, o_counts, o_rc = oracle.voxelgrid(cloud, leaf)
    assert rc == o_rc == 0
    assert out.shape[0] == o_xyzi.shape[0] and 0 < out.shape[0] < cloud.shape[0]
    assert np.array_equal(keys, o_keys) and np.all(np.diff(keys) > 0)
    assert np.array_equal(counts, o_counts) and counts.sum() == cloud.shape[0]
    assert np.array_equal(out[:, :3], o_xyzi[:, :3])
    assert np.array_equal(out[:, 4], o_xyzi[:, 3])
    assert np.all(out[:, 3] == 1.0)


def test_voxelgrid_edge_cases(reg, oracle):
    empty = np.zeros((0, 8), np.float32)
    assert reg.voxelGridFilter(empty, 0.1).shape[0] == 0
    one = np.array([[1, 2, 3, 1, 9, 0, 0, 0]], np.float32)
    out = reg.voxelGridFilter(one, 0.1)
    assert out.shape[0] == 1 and np.array_equal(out[0, :3], one[0, :3]) and out[0, 4] == 9
    # non-finite points are dropped
    pts = np.array([[0, 0, 0, 1, 1, 0, 0, 0], [np.nan, 0, 0, 1, 1, 0, 0, 0], [0.01, 0.01, 0.01, 1, 3, 0, 0, 0], [5, 5, 5, 1, 7, 0, 0, 0]], np.float32)
    out, keys, counts, rc = reg.voxelGridFilter(pts, 0.5, with_keys=True)
    o_xyzi, o_keys, o_counts, _ = oracle.voxelgrid(pts, 0.5)
    assert np.array_equal(keys, o_keys) and np.array_equal(counts, o_counts) and np.array_equal(out[:, :3], o_xyzi[:, :3])
    # leaf too small for int32 indices: PCL warns and passes the input through
    far = np.array([[0, 0, 0, 1, 0, 0, 0, 0], [3000, 3000, 3000, 1, 0, 0, 0, 0]], np.float32)
    out, keys, counts, rc = reg.voxelGridFilter(far, 0.001, with_keys=True)
    _, _, _, o_rc = oracle.voxelgrid(far, 0.001)
    assert rc == o_rc == 1 and np.array_equal(out, far)


def test_voxelgrid_properties_full_size(reg, synth):
    """size-independent properties at the BASELINE KITTI shape: idempotence on the voxel structure, mass conservation"""
    cloud = synth.scan("kitti", frame=7, stride=8)
    out, keys, counts, rc = reg.voxelGridFilter(cloud, 0.25, with_keys=True)
    assert counts.sum() == cloud.shape[0] and np.all(np.diff(keys) > 0)
    w = counts[:, None].astype(np.float64)
    assert np.allclose((out[:, :3].astype(np.float64) * w).sum(0) / cloud.shape[0], cloud[:, :3].astype(np.float64).mean(0), atol=1e-3)
    out2, keys2, counts2, _ = reg.voxelGridFilter(out, 0.25, with_keys=True)
    assert out2.shape[0] <= out.shape[0] and counts2.sum() == out.shape[0]


@pytest.mark.parametrize("sensor,leaf,n", [("hdl32e", 0.1, 131072), ("vlp16", 0.1, 40000), ("vlp16_16k", 0.25, 16384), ("vlp16_16k", 0.5, 777)])
def test_voxelgrid_cluster_sizes_match_oracle(reg, synth, oracle, sensor, leaf, n):
    """every cluster size of the one-kernel path (1, 4 [40 000 points leave the last CTA mostly padding], 8 CTAs), with runs that
    continue across CTA boundaries, non-finite points and a ragged tail"""
    cloud = synth.scan(sensor, frame=5, stride=8)[:n].copy()
    if n > 100:
        cloud[7, 1] = np.nan
        cloud[n - 3, 0] = np.inf
    out, keys, counts, rc = reg.voxelGridFilter(cloud, leaf, with_keys=True)
    o_xyzi, o_keys, o_counts, o_rc = oracle.voxelgrid(cloud, leaf)
    assert rc == o_rc == 0
    assert np.array_equal(keys, o_keys) and np.array_equal(counts, o_counts)
    assert np.array_equal(out[:, :3], o_xyzi[:, :3]) and np.array_equal(out[:, 4], o_xyzi[:, 3])


def test_voxelgrid_device_feeds_registration_without_leaving_hbm(reg, synth, oracle):
    """b2r_voxelgrid_device: raw scan resident in HBM -> downsampled cloud in HBM -> b2r_set_source/target_device -> align; identical
    to the host round trip (downsample to host, upload again)"""
    import torch
    raws = [synth.scan("kitti", frame=f, stride=8) for f in (0, 1)]
    host = [reg.voxelGridFilter(r, 0.25) for r in raws]
    reg.setInputTarget(host[0])
    reg.setInputSource(host[1])
    reg.align(np.eye(4, dtype=np.float32))
    T_host, it_host = reg.getFinalTransformation(), reg.nr_iterations
    dev = [torch.from_numpy(r).cuda() for r in raws]
    keep = []
    for k, d in enumerate(dev):
        ptr, m, rc = reg.voxelGridFilterDevice(d.data_ptr(), d.shape[0], d.shape[1] * 4, 0.25)
        assert rc == 0 and m == host[k].shape[0]
        copy = torch.empty((m, d.shape[1]), dtype=torch.float32, device="cuda")   # the engine's buffer is reused by the next call
        torch.cuda.synchronize()
        from cuda.bindings import runtime as cudart
        err, = cudart.cudaMemcpy(copy.data_ptr(), ptr, m * d.shape[1] * 4, cudart.cudaMemcpyKind.cudaMemcpyDeviceToDevice)
        assert int(err) == 0
        assert np.array_equal(copy.cpu().numpy(), host[k])
        keep.append(copy)
    reg.setInputTargetDevice(keep[0].data_ptr(), keep[0].shape[0], keep[0].shape[1] * 4)
    reg.setInputSourceDevice(keep[1].data_ptr(), keep[1].shape[0], keep[1].shape[1] * 4)
    reg.align(np.eye(4, dtype=np.float32))
    assert np.array_equal(reg.getFinalTransformation(), T_host) and reg.nr_iterations == it_host
