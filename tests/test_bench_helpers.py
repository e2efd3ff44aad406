"""CPU-only: arithmetic helpers of bench.py (no GPU, no engine)."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bench():
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_algorithmic_bytes_match_design_table():
    b = _bench()
    n = 65536
    assert b.algorithmic_bytes("knn_covariance", n, n, 32) == n * 64
    assert b.algorithmic_bytes("gicp_correspondences", n, n, 32) == 16 * 2 * n + 16 * n
    assert b.algorithmic_bytes("gicp_linearize", n, n, 32) == 248 * n + 232
    assert b.algorithmic_bytes("no_such_class", n, n, 32) is None


def test_whole_step_algorithmic_throughput():
    b = _bench()
    n = 65536
    calls = {"knn_covariance": 200, "gicp_correspondences": 974, "gicp_linearize": 974, "misc": 7, "nn_fitness": 0}
    gbs, per_step, unknown = b.whole_step_algorithmic_gbs(calls, n, n, 32, 92.6, 200)
    want = 200 * n * 64 + 974 * (48 * n) + 974 * (248 * n + 232)
    assert per_step == want / 200 and unknown == ["misc"]
    assert abs(gbs - want / 92.6e-3 / 1e9) < 1e-9
    assert b.whole_step_algorithmic_gbs(calls, n, n, 32, 0.0, 200) is None
