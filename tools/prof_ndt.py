"""tiny driver for ncu captures of the NDT path: one 131 072-point HDL-32e pair, a few fixed iterations"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth
cfg = pkg.default_config(pkg.B2R_METHOD_NDT)
cfg.ndt_resolution = 1.0
cfg.ndt_fixed_iterations = 6
reg = pkg.Registration(cfg)
f0, f1 = synth.scan("hdl32e", frame=0), synth.scan("hdl32e", frame=1)
for rep in range(2):
    reg.setInputTarget(f0); reg.setInputSource(f1)
    reg.align(np.eye(4, dtype=np.float32))
print("iters", reg.nr_iterations)
