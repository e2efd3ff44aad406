"""standalone per-class kernel times of the preprocessing chain (development aid): python tools/knn_time.py [frames]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hdl_graph_slam_b200 as pkg
from hdl_graph_slam_b200 import synth

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 12
reg = pkg.select_registration_method({"registration_method": "FAST_GICP"})
frames = [synth.scan("vlp16", frame=7 * k) for k in range(nf)]
reg.setInputSource(frames[0]); reg.synchronize()
reg.getStats(reset=True)
reg.setProfiling(True)
for f in frames[1:]:
    reg.setInputSource(f)
    reg.synchronize()
st = reg.getStats(reset=True)
print("budget", os.environ.get("B2R_KNN_BUDGET", "default"), {k: round(v / (nf - 1) * 1e3, 1) for k, v in st["ms"].items() if v > 0}, "us per frame")
reg.close()
