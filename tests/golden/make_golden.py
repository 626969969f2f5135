#!/usr/bin/env python
"""Generates the committed golden fixtures under tests/golden/ by running the REFERENCE here (CPU container only).

Needs /root/reference (read-only checkout) and oracle/_ref/libpecos_float32.so (``make -C oracle``).  The reference's
Python package is copied to a scratch directory (never into this repo), the reference library is dropped into it, and
one compatibility shim is applied for scipy >= 1.14 / numpy >= 2 (``pecos.utils.smat_util.cs_matrix`` uses removed
APIs).  Nothing here runs on the GPU box: the outputs are small files committed next to this script.

Fixtures written:
  xlinear_toy/model/...        XLinearModel trained by the reference on test/tst-data/xmc/xlinear/{X,Y}.npz with the
                               reference test's settings (test_xlinear.py:346-352: --max-leaf-size 10)
  xlinear_toy/Xt.npz           the reference's test queries (copied data fixture)
  xlinear_toy/expected.npz     reference C++ predictions for a grid of (post_processor, beam, topk), csr and dense;
                               the default-parameter prediction is also checked here against the reference's own golden
                               test/tst-data/xmc/xlinear/Yt_pred.npz (abs 1e-6), which pins the regenerated model.
  xlinear_toy/model_mmap/...   the same model compiled to the mmap format by the reference (inference.hpp:2575-2595)
  xlinear_toy/model_splits{2,4}/...  models of test_predict_consistency_between_python_and_cpp (:132-147)
  hnsw_toy/model_{ip,l2}/...   the reference's prebuilt dense index (test/tst-data/ann/hnsw-model-dense, ip) and an l2
                               index trained single-threaded by the reference on X.trn
  hnsw_toy/X.tst.npy, X.trn.npy, expected.npz   reference search results for efS in {10, 50, 75, 100}, topk in {1, 10}
"""
import json
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as smat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REFERENCE = os.environ.get("REFERENCE", "/root/reference")
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libpecos_float32.so")


def stage_reference_python(scratch):
    dst = os.path.join(scratch, "pecos")
    shutil.copytree(os.path.join(REFERENCE, "pecos"), dst)
    subprocess.run(["chmod", "-R", "u+w", scratch], check=True)
    shutil.copy(REF_LIB, os.path.join(dst, "core", "libpecos_float32.so"))
    # compatibility shim for scipy>=1.14 / numpy>=2 (the reference pins numpy<2)
    p = os.path.join(dst, "utils", "smat_util.py")
    s = open(p).read()
    s = s.replace("smat.sputils.get_index_dtype", "smat._sputils.get_index_dtype")
    s = s.replace("copy=False", "copy=None")
    open(p, "w").write(s)
    return scratch


def save_csr(path, M):
    M = smat.csr_matrix(M, dtype=np.float32)
    M.sort_indices()
    smat.save_npz(path, M, compressed=False)


def main():
    assert os.path.isdir(REFERENCE), "reference checkout not found"
    assert os.path.exists(REF_LIB), "build oracle/_ref first (make -C oracle)"
    scratch = tempfile.mkdtemp(prefix="refpy_")
    stage_reference_python(scratch)
    sys.path.insert(0, scratch)
    env = dict(os.environ, PYTHONPATH=scratch + os.pathsep + os.environ.get("PYTHONPATH", ""))

    from pecos.ann.hnsw import HNSW
    from pecos.xmc.xlinear.model import XLinearModel

    tst = os.path.join(REFERENCE, "test", "tst-data")

    # ------------------------------------------------------------------ XR-Linear toy
    out = os.path.join(HERE, "xlinear_toy")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    X_file, Y_file = os.path.join(tst, "xmc/xlinear/X.npz"), os.path.join(tst, "xmc/xlinear/Y.npz")
    Xt_file = os.path.join(tst, "xmc/xlinear/Xt.npz")

    def train(folder, extra):
        cmd = [sys.executable, "-m", "pecos.xmc.xlinear.train", "-x", X_file, "-y", Y_file, "-m", folder] + extra
        r = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=scratch)
        assert r.returncode == 0, r.stderr

    train(os.path.join(out, "model"), ["--max-leaf-size", "10"])
    train(os.path.join(out, "model_splits2"), ["--nr-splits", "2", "--max-leaf-size", "2"])
    train(os.path.join(out, "model_splits4"), ["--nr-splits", "4", "--max-leaf-size", "2"])
    Xt = smat.load_npz(Xt_file).tocsr().astype(np.float32)
    Xt.sort_indices()
    save_csr(os.path.join(out, "Xt.npz"), Xt)

    expected = {}
    pps = ["noop", "sigmoid", "log-sigmoid", "l1-hinge", "l2-hinge", "l3-hinge", "l4-hinge", "log-l1-hinge",
           "log-l2-hinge", "log-l3-hinge", "log-l4-hinge"]
    index = []
    for name in ["model", "model_splits2", "model_splits4"]:
        m = XLinearModel.load(os.path.join(out, name), is_predict_only=True)
        grid = [(None, None, None)] + [(pp, 2, None) for pp in pps] + [("l3-hinge", 1, 1), ("l3-hinge", 10, 20), (None, 3, 5)]
        for gi, (pp, beam, topk) in enumerate(grid):
            kw = {}
            if pp:
                kw["post_processor"] = pp
            if beam:
                kw["beam_size"] = beam
            if topk:
                kw["only_topk"] = topk
            for kind, Xq in (("csr", Xt), ("drm", np.ascontiguousarray(Xt.toarray()))):
                P = m.predict(Xq, **kw).tocsr()
                key = f"{name}|{gi}|{kind}"
                expected[key + "|indptr"] = P.indptr.astype(np.int64)
                expected[key + "|indices"] = P.indices.astype(np.int64)
                expected[key + "|data"] = P.data.astype(np.float32)
                index.append({"key": key, "model": name, "kind": kind, "post_processor": pp, "beam_size": beam,
                              "only_topk": topk, "shape": list(P.shape)})
        if name == "model":
            gold = smat.load_npz(os.path.join(tst, "xmc/xlinear/Yt_pred.npz")).toarray()
            mine = m.predict(Xt).toarray()
            assert np.abs(gold - mine).max() <= 1e-6, "regenerated toy model does not reproduce the reference golden Yt_pred.npz"
            np.save(os.path.join(out, "Yt_pred_reference_golden.npy"), gold.astype(np.float32))
    np.savez(os.path.join(out, "expected.npz"), **expected)
    json.dump(index, open(os.path.join(out, "expected_index.json"), "w"), indent=1)
    # compiled (mmap) copy of the toy model, written by the reference's c_xlinear_compile_mmap_model
    XLinearModel.compile_mmap_model(os.path.join(out, "model"), os.path.join(out, "model_mmap"))

    # ------------------------------------------------------------------ HNSW toy
    out = os.path.join(HERE, "hnsw_toy")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    X_trn = np.load(os.path.join(tst, "ann/X.trn.l2-normalized.npy")).astype(np.float32)
    X_tst = np.load(os.path.join(tst, "ann/X.tst.l2-normalized.npy")).astype(np.float32)
    np.save(os.path.join(out, "X.trn.npy"), X_trn)
    np.save(os.path.join(out, "X.tst.npy"), X_tst)
    src = os.path.join(tst, "ann/hnsw-model-dense")
    dst = os.path.join(out, "model_ip")
    os.makedirs(os.path.join(dst, "c_model"))
    shutil.copy(os.path.join(src, "param.json"), dst)
    for f in ("config.json", "index.mmap_store"):  # the legacy index.bin is ignored by the v2.0 loader (hnsw.hpp:537-551)
        shutil.copy(os.path.join(src, "c_model", f), os.path.join(dst, "c_model", f))
    subprocess.run(["chmod", "-R", "u+w", dst], check=True)
    train_params = HNSW.TrainParams(M=8, efC=40, metric_type="l2", threads=1)
    m_l2 = HNSW.train(X_trn, train_params=train_params, pred_params=HNSW.PredParams(efS=50, topk=10, threads=1))
    m_l2.save(os.path.join(out, "model_l2"))
    train_params = HNSW.TrainParams(M=6, efC=30, metric_type="ip", threads=1, max_level_upper_bound=3)
    m_ip2 = HNSW.train(X_trn, train_params=train_params, pred_params=HNSW.PredParams(efS=50, topk=10, threads=1))
    m_ip2.save(os.path.join(out, "model_ip_small_m"))
    expected, index = {}, []
    for name in ["model_ip", "model_l2", "model_ip_small_m"]:
        m = HNSW.load(os.path.join(out, name))
        for efS in (10, 50, 75, 100):
            for topk in (1, 10, 20):
                pp = HNSW.PredParams(efS=efS, topk=topk, threads=1)
                searchers = m.searchers_create(1)
                idx, dist = m.predict(X_tst, pred_params=pp, searchers=searchers, ret_csr=False)
                key = f"{name}|{efS}|{topk}"
                expected[key + "|idx"] = idx.astype(np.uint32)
                expected[key + "|dist"] = dist.astype(np.float32)
                index.append({"key": key, "model": name, "efS": efS, "topk": topk})
    np.savez(os.path.join(out, "expected.npz"), **expected)
    json.dump(index, open(os.path.join(out, "expected_index.json"), "w"), indent=1)
    import platform

    flags = [ln for ln in open("/proc/cpuinfo") if ln.startswith("flags")][0]
    isa = "avx512f" if " avx512f" in flags else ("avx" if " avx " in flags else "sse")
    json.dump({"generated_on": platform.processor() or platform.machine(), "distance_isa_clone": isa,
               "note": "HNSW distances depend on the SIMD clone the reference selects at run time (distance_impl/x86.hpp)"},
              open(os.path.join(out, "provenance.json"), "w"), indent=1)
    shutil.rmtree(scratch, ignore_errors=True)
    print("golden fixtures written under", HERE)


if __name__ == "__main__":
    main()
