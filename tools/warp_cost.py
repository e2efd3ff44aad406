"""offline tail analysis of the search kernels on the CPU warp emulator (development aid): python tools/warp_cost.py [kf src [copies]]"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
from hdl_graph_slam_b200 import synth

kf = int(sys.argv[1]) if len(sys.argv) > 1 else 5
src = int(sys.argv[2]) if len(sys.argv) > 2 else 7
copies = sys.argv[3] if len(sys.argv) > 3 else "4"
stride = sys.argv[4] if len(sys.argv) > 4 else "1"
out = os.path.join(ROOT, "build")
os.makedirs(out, exist_ok=True)
exe = os.path.join(out, "warp_cost")
srcs = [os.path.join(ROOT, "tools", "warp_cost.cpp"), os.path.join(ROOT, "hdl_graph_slam_b200", "csrc", "bvh.cuh"), os.path.join(ROOT, "tests", "warp_emu.hpp")]
if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in srcs):
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-w", "-ffp-contract=off", "-I/usr/local/cuda/include", "-o", exe, srcs[0]])
t = synth.scan("vlp16", frame=kf, stride=4)
s = synth.scan("vlp16", frame=src, stride=4)
t.tofile(os.path.join(out, "wc_t.f32")); s.tofile(os.path.join(out, "wc_s.f32"))
# guess = relative pose of the frame BEFORE the source (what the odometry's motion model would hand in), then the true one
for name, f in (("guess from previous frame", src - 1), ("true pose", src)):
    rel = np.linalg.inv(synth.pose_matrix(kf)) @ synth.pose_matrix(f)
    np.savetxt(os.path.join(out, "wc_pose.txt"), rel[:3, :].astype(np.float32).reshape(1, 12))
    for mode in (0, 1):
        print(f"== keyframe {kf} source {src}, {name}, copies {copies}")
        subprocess.check_call([exe, os.path.join(out, "wc_t.f32"), os.path.join(out, "wc_s.f32"), str(t.shape[0]), os.path.join(out, "wc_pose.txt"), copies, str(mode), stride])
print("== 20-NN self")
subprocess.check_call([exe, os.path.join(out, "wc_t.f32"), os.path.join(out, "wc_s.f32"), str(t.shape[0]), os.path.join(out, "wc_pose.txt"), copies, "2", stride])
