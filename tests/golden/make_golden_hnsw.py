#!/usr/bin/env python
"""Generates tests/golden/hnsw_mid/: REFERENCE-built HNSW indices with realistic dimensions + reference search results.

Why: the reference's own fixture index (hnsw_toy) has feat_dim = 2, so the 16-lane main loop of the distance kernel, the
permuted vector layout, the 4-wide remainder / scalar tail and the bulk-copy ring never run on it.  These fixtures make
those paths part of the committed, reference-pinned goldens (no oracle/_ref needed at test time):

  ip_d768   N =  600, d = 768 (48 full 16-lane blocks, no tail)          M = 16, efC = 60, metric ip
  l2_d128   N = 1500, d = 128                                            M = 12, efC = 60, metric l2
  ip_d70    N = 1500, d =  70 (4 blocks + 4-wide remainder + 2 scalars)  M =  8, efC = 40, metric ip
  l2_dup    N =  900, d =  20, every point 6 times (exact distance ties) M =  8, efC = 50, metric l2

Runs HERE only (CPU container): needs oracle/_ref/libpecos_float32.so (``make -C oracle``; the unmodified reference
library).  Indices are trained single-threaded (deterministic) by the reference's c_ann_hnsw_train_*, saved by its
c_ann_hnsw_save_*, and searched by its c_ann_hnsw_predict_* for efS in {10, 200, 600} x topk in {10, 100}; ids and distance
BITS are recorded (this container's CPU selects the avx512f clone of the reference's distance kernels -- recorded in
provenance.json -- which is the summation order the CUDA kernel restates).
"""
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

CASES = [
    dict(name="ip_d768", N=600, d=768, M=16, efC=60, metric="ip", dup=1, seed=101, nq=48),
    dict(name="l2_d128", N=1500, d=128, M=12, efC=60, metric="l2", dup=1, seed=102, nq=64),
    dict(name="ip_d70", N=1500, d=70, M=8, efC=40, metric="ip", dup=1, seed=103, nq=64),
    dict(name="l2_dup", N=150, d=20, M=8, efC=50, metric="l2", dup=6, seed=104, nq=48),
]
GRID = [(10, 10), (200, 10), (600, 10), (200, 100), (5, 40)]


def main():
    import oracle
    from oracle import ref

    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = os.path.join(HERE, "hnsw_mid")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    expected, index = {}, []
    for c in CASES:
        rng = np.random.default_rng(c["seed"])
        base = rng.standard_normal((c["N"], c["d"])).astype(np.float32)
        base /= np.linalg.norm(base, axis=1, keepdims=True)
        X = np.concatenate([base] * c["dup"], axis=0)
        if c["dup"] > 1:
            Q = base[: c["nq"]].copy()  # queries ON the duplicated points: every distance is tied six-fold
        else:
            Q = rng.standard_normal((c["nq"], c["d"])).astype(np.float32)
            Q /= np.linalg.norm(Q, axis=1, keepdims=True)
        folder = os.path.join(out, c["name"])
        os.makedirs(folder)
        r = ref.RefHNSW.train(X, M=c["M"], efC=c["efC"], metric=c["metric"], threads=1)
        r.save(os.path.join(folder, "c_model"))
        json.dump({"model": "HNSW", "data_type": "drm", "metric_type": c["metric"], "num_item": int(X.shape[0]),
                   "feat_dim": int(X.shape[1]), "pred_kwargs": {"efS": 50, "topk": 10, "threads": 1}},
                  open(os.path.join(folder, "param.json"), "w"))
        np.save(os.path.join(folder, "Q.npy"), Q)
        r2 = ref.RefHNSW.load(os.path.join(folder, "c_model"), c["metric"])  # search the SAVED index
        for efS, topk in GRID:
            idx, dist = r2.predict(Q, efS, topk, threads=1)
            key = f"{c['name']}|{efS}|{topk}"
            expected[key + "|idx"] = idx.astype(np.uint32)
            expected[key + "|dist"] = dist.astype(np.float32)
            index.append({"key": key, "model": c["name"], "efS": efS, "topk": topk})
    np.savez_compressed(os.path.join(out, "expected.npz"), **expected)
    json.dump(index, open(os.path.join(out, "expected_index.json"), "w"), indent=1)
    flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
    isa = "avx512f" if " avx512f" in flags else ("avx" if " avx " in flags else "sse")
    json.dump({"distance_isa_clone": isa, "cases": CASES, "grid": GRID,
               "note": "indices trained (threads=1), saved and searched by the unmodified reference library"},
              open(os.path.join(out, "provenance.json"), "w"), indent=1)
    print("written", out)


if __name__ == "__main__":
    main()
