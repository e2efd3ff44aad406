"""tiny driver for ncu captures: one GICP odometry frame pair at 64k"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
f0, f1 = synth.scan("vlp16", frame=0), synth.scan("vlp16", frame=1)
for rep in range(3):
    reg.setInputTarget(f0); reg.setInputSource(f1)
    g = np.eye(4, dtype=np.float32); g[0, 3] = 0.9
    reg.align(g)
print("iters", reg.nr_iterations, reg.hasConverged())
