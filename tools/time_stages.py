"""quick wall-clock stage timing on the GPU box (development aid, not the bench)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth

def t(f, n=1):
    t0 = time.perf_counter()
    for _ in range(n):
        f()
    return (time.perf_counter() - t0) / n * 1e3

for sensor, method in (("vlp16", "FAST_GICP"), ("hdl32e", "NDT_OMP")):
    reg = pkg.select_registration_method({"registration_method": method, "reg_resolution": 1.0})
    frames = [synth.scan(sensor, frame=k) for k in range(6)]
    reg.setInputTarget(frames[0]); reg.synchronize()
    for rep in range(3):
        ms_t = t(lambda: (reg.setInputTarget(frames[0]), reg.synchronize()))
        ms_s = t(lambda: (reg.setInputSource(frames[1]), reg.synchronize()))
        g = np.eye(4, dtype=np.float32); g[0, 3] = 0.9
        ms_a = t(lambda: reg.align(g), 5)
        ms_f = t(lambda: reg.getFitnessScore(), 5)
        print(f"{method} {sensor} n={frames[0].shape[0]}: set_target {ms_t:.3f} ms  set_source {ms_s:.3f} ms  align {ms_a:.3f} ms ({reg.nr_iterations} it, conv={reg.hasConverged()})  fitness {ms_f:.3f} ms", flush=True)
    if method == "FAST_GICP":
        odo = pkg.ScanMatchingOdometry(reg, 1.0, 1.0, 10000.0)
        odo.matching(0, frames[0])
        t0 = time.perf_counter()
        for k in range(1, 6):
            st = odo.matching(0.1 * k, frames[k])
        print(f"odometry e2e: {(time.perf_counter() - t0) / 5 * 1e3:.3f} ms/frame, last iterations {st['iterations']} kf={st['keyframe_updated']}")
    out = t(lambda: reg.voxelGridFilter(frames[2], 0.1), 3)
    print(f"voxelgrid 0.1 on {frames[2].shape[0]} pts: {out:.3f} ms -> {reg.voxelGridFilter(frames[2], 0.1).shape[0]} voxels")
    reg.close()
