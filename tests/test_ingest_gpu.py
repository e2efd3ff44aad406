"""GPU: wire / disk formats at the seam ("next" row f-4): a sensor_msgs/PointCloud2 blob or a binary PCD body is uploaded unconverted
and unpacked on the device into pcl::PointXYZI records — what pcl::fromROSMsg (apps/scan_matching_odometry_nodelet.cpp:118-119) and
pcl::io::loadPCDFile (src/hdl_graph_slam/keyframe.cpp:141) do on the host.  Checked by registering the device cloud as a target and
reading it back through the exact 1-NN search (distance 0, index i), and against a registration of the host-converted cloud."""
import numpy as np
import pytest
import hdl_graph_slam_b200 as pkg

pytestmark = pytest.mark.gpu


def velodyne_blob(cloud, big_endian=False):
    """the usual velodyne_pointcloud layout: x y z (0,4,8) intensity FLOAT32 @16 ring UINT16 @20, point_step 32"""
    n = cloud.shape[0]
    dt = np.dtype({"names": ["x", "y", "z", "intensity", "ring"], "formats": [">f4" if big_endian else "<f4"] * 4 + [">u2" if big_endian else "<u2"],
                   "offsets": [0, 4, 8, 16, 20], "itemsize": 32})
    rec = np.zeros(n, dt)
    rec["x"], rec["y"], rec["z"], rec["intensity"], rec["ring"] = cloud[:, 0], cloud[:, 1], cloud[:, 2], cloud[:, 4], np.arange(n) % 16
    return rec.tobytes(), dict(point_step=32, off_x=0, off_y=4, off_z=8, off_intensity=16, intensity_datatype=7, is_bigendian=big_endian)


def ouster_blob(cloud):
    """odd offsets, UINT16 intensity, 48-byte step, x y z not at the front"""
    n = cloud.shape[0]
    dt = np.dtype({"names": ["t", "x", "y", "z", "reflectivity", "intensity"], "formats": ["<u4", "<f4", "<f4", "<f4", "<u2", "<u2"],
                   "offsets": [0, 6, 10, 14, 20, 26], "itemsize": 48})
    rec = np.zeros(n, dt)
    rec["x"], rec["y"], rec["z"] = cloud[:, 0], cloud[:, 1], cloud[:, 2]
    rec["intensity"] = np.clip(cloud[:, 4], 0, 65535).astype(np.uint16)
    return rec.tobytes(), dict(point_step=48, off_x=6, off_y=10, off_z=14, off_intensity=26, intensity_datatype=4, is_bigendian=False)


@pytest.mark.parametrize("maker", ["velodyne", "velodyne_be", "ouster"])
def test_pointcloud2_unpacked_on_device(synth, maker):
    cloud = synth.scan("vlp16_16k", frame=2, stride=8)
    blob, lay = {"velodyne": lambda: velodyne_blob(cloud), "velodyne_be": lambda: velodyne_blob(cloud, True), "ouster": lambda: ouster_blob(cloud)}[maker]()
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    dptr = reg.ingestPointCloud2(blob, cloud.shape[0], **lay)
    reg.setInputTargetDevice(dptr, cloud.shape[0], 32)
    idx, d2 = reg.nearestKSearch(cloud)
    assert np.array_equal(idx, np.arange(cloud.shape[0])) and np.all(d2 == 0)   # every point is where the host conversion puts it
    # and the registration of a source against the ingested target equals the one against the host-converted cloud
    src = synth.scan("vlp16_16k", frame=3, stride=8)
    reg.setInputSource(src)
    reg.align(np.eye(4, dtype=np.float32))
    T1, it1 = reg.getFinalTransformation(), reg.nr_iterations
    reg.setInputTarget(cloud)
    reg.setInputSource(src)
    reg.align(np.eye(4, dtype=np.float32))
    assert np.array_equal(T1, reg.getFinalTransformation()) and it1 == reg.nr_iterations
    reg.close()


def write_binary_pcd(path, cloud):
    """what pcl::io::savePCDFileBinary writes for PointXYZI (KeyFrame::save, keyframe.cpp:57): packed x y z intensity"""
    n = cloud.shape[0]
    hdr = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\nCOUNT 1 1 1 1\n"
           f"WIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA binary\n")
    body = np.ascontiguousarray(cloud[:, [0, 1, 2, 4]], np.float32).tobytes()
    with open(path, "wb") as f:
        f.write(hdr.encode())
        f.write(body)


def test_binary_pcd_ingest(synth, tmp_path):
    cloud = synth.scan("vlp16_16k", frame=4, stride=8)
    path = tmp_path / "cloud.pcd"
    write_binary_pcd(path, cloud)
    reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
    dptr, n = reg.ingestPcd(path)
    assert n == cloud.shape[0]
    reg.setInputTargetDevice(dptr, n, 32)
    idx, d2 = reg.nearestKSearch(cloud)
    assert np.array_equal(idx, np.arange(n)) and np.all(d2 == 0)
    with pytest.raises(pkg.B2RError):
        reg.ingestPcd(tmp_path / "missing.pcd")
    (tmp_path / "ascii.pcd").write_text("VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 1\nHEIGHT 1\nPOINTS 1\nDATA ascii\n0 0 0\n")
    with pytest.raises(pkg.B2RError):
        reg.ingestPcd(tmp_path / "ascii.pcd")
    reg.close()
