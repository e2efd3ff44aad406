"""CPU-only: the pcl::Registration adapter shipped for the reference (hdl_graph_slam_b200/adapter/b200_registration.hpp)
compiles against the C ABI header (with a test-only stand-in for the PCL/Eigen declarations it uses) and links to the
library: a tiny program instantiates it through the base-class pointer the reference's factory returns."""
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PROG = r'''
#include "pcl_shim.hpp"
#include "b200_registration.hpp"
#include <cstdio>
int main() {
  using PointT = pcl::PointXYZI;
  b2r_config cfg;
  b2r_config_default(&cfg, B2R_METHOD_GICP);
  try {
    pcl::Registration<PointT, PointT>::Ptr reg(new hdl_graph_slam::B200Registration<PointT>(cfg));  // the factory's return type
    auto cloud = std::make_shared<pcl::PointCloud<PointT>>();
    cloud->points.resize(4);
    reg->setInputTarget(cloud);
    reg->setInputSource(cloud);
    pcl::PointCloud<PointT> out;
    Eigen::Matrix4f guess{};
    for (int i = 0; i < 4; i++) guess.m[i * 5] = 1.f;
    reg->align(out, guess);
    std::printf("aligned converged=%d\n", (int)reg->hasConverged());
  } catch (const std::exception& e) {
    std::printf("no device: %s\n", e.what());  // expected in the CPU container: b2r_create -> B2R_ENODEVICE, loudly
  }
  return 0;
}
'''


def test_adapter_compiles_and_links(tmp_path):
    from hdl_graph_slam_b200 import build
    build.build_engine()
    src = tmp_path / "adapter_main.cpp"
    src.write_text(PROG)
    exe = tmp_path / "adapter_main"
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    lib = os.path.join(ROOT, "hdl_graph_slam_b200", "_lib")
    subprocess.check_call([cxx, "-std=c++14", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "hdl_graph_slam_b200", "adapter"),
                           "-o", str(exe), str(src), "-L", lib, "-lb200reg", "-Wl,-rpath," + lib])
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, out.stderr
    assert "no device" in out.stdout or "aligned" in out.stdout
