"""instrumented-variant driver: a few odometry frames, per-warp profiles of the LAST align are left in gpurun_out/ (development aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
odo = pkg.ScanMatchingOdometry(reg, 1.0, 1.0, 10000.0)
for k in range(nf):
    st = odo.matching(0.1 * k, synth.scan("vlp16", frame=k))
    print(k, st["iterations"], st["keyframe_updated"], flush=True)
cp = np.zeros(65536, np.int32)
