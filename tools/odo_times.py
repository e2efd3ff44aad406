"""per-frame kernel-class times of the odometry chain (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
odo = pkg.ScanMatchingOdometry(reg, 1.0, 1.0, 10000.0)
reg.setProfiling(True)
quiet = len(sys.argv) > 2
corr = []
for k in range(nf):
    st = odo.matching(0.1 * k, synth.scan("vlp16", frame=k))
    reg.synchronize()
    s = reg.getStats(reset=True)
    per = {n: round(s["ms"][n] / max(s["calls"][n], 1) * 1e3, 1) for n in s["ms"] if s["calls"][n]}
    corr.append(per.get("gicp_correspondences", 0))
    if not quiet:
        print(k, "it", st["iterations"], "kf", st["keyframe_updated"], per, flush=True)
print("correspondences us/call per frame:", corr[1:], "mean", round(float(np.mean(corr[1:])), 1))
