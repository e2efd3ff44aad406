"""CPU-only: the engine's exact BVH search logic (bvh.cuh compiled as HOST code by tests/host_harness.cu) returns the
oracle's (d2, index) results bit-for-bit, including lattice ties, range-limited search and queries far outside the cloud."""
import os
import subprocess
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(oracle):
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    exe = os.path.join(ROOT, "build", "host_harness")
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
                           "-Xcompiler", "-ffp-contract=off", "-o", exe, os.path.join(ROOT, "tests", "host_harness.cu"),
                           "-L" + os.path.join(ROOT, "oracle"), "-loracle", "-Xlinker", "-rpath=" + os.path.join(ROOT, "oracle")])
    return exe


@pytest.mark.parametrize("mode,n", [(0, 20000), (1, 20000), (2, 20000), (0, 1500)])
def test_bvh_search_equals_kdtree(harness, mode, n):
    out = subprocess.run([harness, str(n), "4000", str(mode)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "1nn_mismatch=0" in out.stdout and "knn_mismatch=0" in out.stdout and "limited_mismatch=0" in out.stdout


def test_min_eigenvector_solver_matches_jacobi(tmp_path):
    exe = tmp_path / "eig_harness"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-w", "-ccbin", "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++",
                           "-o", str(exe), os.path.join(ROOT, "tests", "eig_harness.cu")])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
