#!/usr/bin/env python
"""Generates tests/golden/hnsw_sparse/: REFERENCE-built sparse (csr) HNSW indices + reference search results.

  fixture_ip   the reference's own prebuilt sparse index (test/tst-data/ann/hnsw-model-sparse: 90 x 2, ip, M = 24) with its
               test queries (X.tst.l2-normalized.npy as csr), the case of test/pecos/ann/test_hnsw.py:86-124
  ip_tfidf     N = 3000, D = 20000, ~60 stored entries per row drawn from a Zipf-like popularity law, rows L2-normalised,
               every 97th row EMPTY                                                       M = 12, efC = 60, metric ip
  l2_tfidf     N = 2000, D = 5000, ~40 entries per row (the reference's sparse "l2" is -2<x,y>: feat_vectors.hpp:186-192)
                                                                                          M = 8,  efC = 40, metric l2
  ip_short     N = 1500, D = 64, 1..9 entries per row (rows shorter than the 4-wide intersection blocks, many exact ties)
                                                                                          M = 8,  efC = 40, metric ip

Runs HERE only (CPU container): needs oracle/_ref/libpecos_float32.so.  Indices are trained single-threaded by the reference's
c_ann_hnsw_train_csr_*, saved by c_ann_hnsw_save_csr_*, re-loaded and searched by c_ann_hnsw_predict_csr_*; ids and distance
BITS are recorded.  Query batches contain empty rows and one row longer than any stored row.
"""
import json
import os
import shutil
import sys

import numpy as np
import scipy.sparse as smat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REFERENCE = os.environ.get("PECOS_REFERENCE", "/root/reference")

CASES = [
    dict(name="ip_tfidf", N=3000, D=20000, nnz=60, M=12, efC=60, metric="ip", seed=201, nq=64, empty_every=97),
    dict(name="l2_tfidf", N=2000, D=5000, nnz=40, M=8, efC=40, metric="l2", seed=202, nq=64, empty_every=0),
    dict(name="ip_short", N=1500, D=64, nnz=5, M=8, efC=40, metric="ip", seed=203, nq=48, empty_every=0),
]
GRID = [(10, 10), (200, 10), (400, 10), (100, 50), (5, 40)]


def make_rows(seed, n, D, nnz, empty_every, normalise=True, long_row=None):
    rng = np.random.default_rng(seed)
    p = 1.0 / (np.arange(D) + 20.0)
    p /= p.sum()
    indptr, idx, val = [0], [], []
    for i in range(n):
        k = int(rng.integers(max(1, nnz // 2), nnz * 2)) if nnz > 8 else int(rng.integers(1, 2 * nnz))
        if empty_every and i % empty_every == 5:
            k = 0
        if long_row is not None and i == long_row[0]:
            k = long_row[1]
        k = min(k, D)
        c = np.sort(rng.choice(D, size=k, replace=False, p=p)).astype(np.uint32)
        v = np.abs(rng.standard_normal(k)).astype(np.float32) + np.float32(0.01)
        if nnz <= 8:  # coarse values: exact distance ties
            v = np.round(v * 2).astype(np.float32) / 2 + np.float32(0.5)
        if normalise and k:
            v = (v / np.linalg.norm(v)).astype(np.float32)
        idx.append(c)
        val.append(v)
        indptr.append(indptr[-1] + k)
    return smat.csr_matrix((np.concatenate(val), np.concatenate(idx), np.array(indptr)), shape=(n, D), dtype=np.float32)


def record(folder, X_index, Q, metric, expected, index, name, r=None, M=0, efC=0):
    from oracle import ref

    if r is None:
        r = ref.RefHNSW.train(X_index, M=M, efC=efC, metric=metric, threads=1)
        r.save(os.path.join(folder, "c_model"))
        json.dump({"model": "HNSW", "data_type": "csr", "metric_type": metric, "num_item": int(X_index.shape[0]),
                   "feat_dim": int(X_index.shape[1]), "pred_kwargs": {"efS": 50, "topk": 10, "threads": 1}},
                  open(os.path.join(folder, "param.json"), "w"))
    smat.save_npz(os.path.join(folder, "Q.npz"), Q)
    r2 = ref.RefHNSW.load(os.path.join(folder, "c_model"), metric, data_type="csr")  # search the SAVED index
    for efS, topk in GRID:
        idx, dist = r2.predict(Q, efS, topk, threads=1)
        key = f"{name}|{efS}|{topk}"
        expected[key + "|idx"] = idx.astype(np.uint32)
        expected[key + "|dist"] = dist.astype(np.float32)
        index.append({"key": key, "model": name, "efS": efS, "topk": topk})


def main():
    import oracle

    assert oracle.have_ref(), "build oracle/_ref first (make -C oracle)"
    out = os.path.join(HERE, "hnsw_sparse")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    expected, index = {}, []
    # the reference's own sparse fixture index, copied as data (config.json + index.mmap_store + param.json)
    src = os.path.join(REFERENCE, "test", "tst-data", "ann")
    folder = os.path.join(out, "fixture_ip")
    os.makedirs(os.path.join(folder, "c_model"))
    shutil.copy(os.path.join(src, "hnsw-model-sparse", "param.json"), os.path.join(folder, "param.json"))
    for f in ("config.json", "index.mmap_store"):
        shutil.copy(os.path.join(src, "hnsw-model-sparse", "c_model", f), os.path.join(folder, "c_model", f))
    Q = smat.csr_matrix(np.load(os.path.join(src, "X.tst.l2-normalized.npy")).astype(np.float32))
    Q.sort_indices()
    record(folder, None, Q, "ip", expected, index, "fixture_ip", r=True)
    for c in CASES:
        X = make_rows(c["seed"], c["N"], c["D"], c["nnz"], c["empty_every"])
        Q = make_rows(c["seed"] + 1000, c["nq"], c["D"], c["nnz"], 13, long_row=(7, min(c["D"], 6 * c["nnz"] + 3)))
        folder = os.path.join(out, c["name"])
        os.makedirs(folder)
        record(folder, X, Q, c["metric"], expected, index, c["name"], M=c["M"], efC=c["efC"])
    np.savez_compressed(os.path.join(out, "expected.npz"), **expected)
    json.dump(index, open(os.path.join(out, "expected_index.json"), "w"), indent=1)
    json.dump({"cases": CASES, "grid": GRID,
               "note": "indices trained (threads=1), saved and searched by the unmodified reference library; fixture_ip is the "
                       "reference's own prebuilt test index"},
              open(os.path.join(out, "provenance.json"), "w"), indent=1)
    print("written", out)


if __name__ == "__main__":
    main()
