"""Generates tests/golden/golden_v1.npz — frozen outputs of the CPU oracle on seeded synthetic inputs.

The reference (koide3/hdl_graph_slam) ships no tests or golden vectors for this path and its third-party math cannot be
built or imported here: these fixtures therefore freeze the ORACLE (after it passed the analytic
known-answer tests of tests/test_oracle_kat.py), so that later changes to the oracle or the engine are caught.
Run:  python tests/golden/gen_golden.py     (CPU only, a few seconds)
"""
import hashlib
import os
import sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from hdl_graph_slam_b200 import synth  # noqa: E402
from oracle import oracle as orc  # noqa: E402
from common import perturb  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    orc.build()
    tgt = synth.scan("vlp16_16k", frame=0, stride=8)
    src = synth.scan("vlp16_16k", frame=1, stride=8)
    G = {"input_sha_tgt": sha(tgt), "input_sha_src": sha(src)}
    # --- exact NN
    idx, d2 = orc.knn(tgt, src, 1)
    G["nn_idx_sha"], G["nn_d2_sha"] = sha(idx), sha(d2)
    # --- GICP
    sc, tc = orc.gicp_covariances(src, 20), orc.gicp_covariances(tgt, 20)
    G["cov_src_sample"] = sc[::1024].copy()
    T1 = perturb(7, 0.3, 2.0)
    for name, T in (("I", np.eye(4)), ("P", T1)):
        o = orc.gicp_linearize(src, sc, tgt, tc, T, 2.5)
        G[f"lin_{name}_corr_sha"] = sha(o["corr"])
        G[f"lin_{name}_H"], G[f"lin_{name}_b"], G[f"lin_{name}_err"] = o["H"], o["b"], np.float64(o["err"])
        G[f"lin_{name}_nvalid"] = np.int64((o["corr"] >= 0).sum())
    G["lin_P_T"] = T1
    for name, guess in (("I", np.eye(4, dtype=np.float32)), ("P", perturb(8, 0.4, 2.5).astype(np.float32))):
        r = orc.gicp_align(src, tgt, guess)
        G[f"gicp_{name}_guess"], G[f"gicp_{name}_T"] = guess, r["T"]
        G[f"gicp_{name}_iters"], G[f"gicp_{name}_conv"], G[f"gicp_{name}_corr_sha"] = np.int64(r["iterations"]), np.int64(r["converged"]), sha(r["corr"])
        s, n, i = orc.fitness(tgt, src, r["T"], 2.5, 0.25)
        G[f"fit_{name}"] = np.array([s, n, i], np.float64)
    # --- NDT
    for res in (1.0, 0.5):
        m = orc.NdtMap(tgt, res)
        d = m.dump()
        tag = f"ndt{int(res * 10):02d}"
        G[f"{tag}_keys_sha"], G[f"{tag}_npts_sha"], G[f"{tag}_nvox"] = sha(d["keys"]), sha(d["npts"]), np.int64(len(d["keys"]))
        G[f"{tag}_mean_sample"], G[f"{tag}_icov_sample"] = d["mean"][::97].copy(), d["icov"][::97].copy()
        p = np.array([0.9, 0.05, -0.02, 0.01, -0.015, 0.03])
        o = m.derivatives(src, p)
        G[f"{tag}_p"], G[f"{tag}_score"], G[f"{tag}_g"], G[f"{tag}_H"], G[f"{tag}_pairs"] = p, np.float64(o["score"]), o["g"], o["H"], np.int64(o["n_pairs"])
        G[f"{tag}_cells_sha"] = sha(o["cells"])
        guess = perturb(9, 0.3, 1.5).astype(np.float32)
        r = m.align(src, guess)
        G[f"{tag}_guess"], G[f"{tag}_T"], G[f"{tag}_iters"], G[f"{tag}_conv"] = guess, r["T"], np.int64(r["iterations"]), np.int64(r["converged"])
    # --- voxel grid
    out, keys, counts, rc = orc.voxelgrid(tgt, 0.1)
    G["vg_n"], G["vg_keys_sha"], G["vg_counts_sha"], G["vg_xyzi_sha"] = np.int64(len(keys)), sha(keys), sha(counts), sha(out)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.npz")
    np.savez_compressed(path, **{k: np.asarray(v) for k, v in G.items()})
    print("wrote", path, os.path.getsize(path), "bytes,", len(G), "entries")


if __name__ == "__main__":
    main()
