#!/usr/bin/env python
"""Generates tests/golden/mlmodel_toy/: a single-layer mmap model written by the REFERENCE's c_mlmodel_compile_mmap_model from
the committed toy model's leaf layer (tests/golden/xlinear_toy/model/ranker/<last>.model) and the reference's
c_mlmodel_predict_* / c_mlmodel_predict_on_selected_outputs_* results on the toy queries (pecos/core/libpecos.cpp:37-113).
Runs HERE only (needs oracle/_ref); the outputs are small files committed next to this script."""
import json
import os
import shutil
import sys

import numpy as np
import scipy.sparse as smat

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref

    toy = os.path.join(HERE, "xlinear_toy")
    depth = json.load(open(os.path.join(toy, "model", "ranker", "param.json")))["depth"]
    out = os.path.join(HERE, "mlmodel_toy")
    shutil.rmtree(out, ignore_errors=True)
    os.makedirs(out)
    ref.compile_mlmodel_mmap(os.path.join(toy, "model", "ranker", f"{depth - 1}.model"), os.path.join(out, "layer_mmap"))
    m = ref.MLModelHandle(os.path.join(out, "layer_mmap"))
    Xt = smat.load_npz(os.path.join(toy, "Xt.npz")).tocsr().astype(np.float32)
    Xt.sort_indices()
    nr_codes, nr_labels = m.attr("nr_codes"), m.attr("nr_labels")
    rng = np.random.default_rng(3)
    sel = smat.csr_matrix((rng.random((Xt.shape[0], nr_labels)) < 0.4).astype(np.float32))
    # csr_codes as the reference's own callers build it (pecos/xmc/base.py:1771-1780): the parents of the selected labels
    # (selected x C) -- a selected label whose parent is missing from csr_codes is outside the reference's contract -- plus a
    # few extra codes, with random positive values
    C = smat.load_npz(os.path.join(toy, "model", "ranker", f"{depth - 1}.model", "C.npz")).tocsr()
    pattern = ((sel @ C) + smat.csr_matrix((rng.random((Xt.shape[0], nr_codes)) < 0.3).astype(np.float32))).tocsr()
    pattern.sort_indices()
    codes = smat.csr_matrix((0.05 + rng.random(pattern.nnz).astype(np.float32), pattern.indices, pattern.indptr), shape=pattern.shape)
    smat.save_npz(os.path.join(out, "codes.npz"), codes, compressed=False)
    smat.save_npz(os.path.join(out, "selected.npz"), sel, compressed=False)
    E, index = {}, []

    def rec(key, P, **info):
        P = P.tocsr()
        E[key + "|indptr"], E[key + "|indices"], E[key + "|data"] = P.indptr.astype(np.int64), P.indices.astype(np.int64), P.data.astype(np.float32)
        index.append(dict(key=key, shape=list(P.shape), **info))

    for kind, Xq in (("csr", Xt), ("drm", np.ascontiguousarray(Xt.toarray()))):
        for ci, cc in (("codes", codes), ("nocodes", None)):
            for pp in (None, "sigmoid", "log-l2-hinge", "noop"):
                for topk in (0, 3):
                    rec(f"predict|{kind}|{ci}|{pp}|{topk}", m.predict(Xq, cc, pp, topk), op="predict", kind=kind, codes=ci, post_processor=pp, only_topk=topk)
                rec(f"selected|{kind}|{ci}|{pp}", m.predict_on_selected_outputs(Xq, sel, cc, pp), op="selected", kind=kind, codes=ci, post_processor=pp)
    np.savez(os.path.join(out, "expected.npz"), **E)
    json.dump({"attrs": {"nr_codes": nr_codes, "nr_labels": nr_labels, "nr_features": m.attr("nr_features")}, "entries": index},
              open(os.path.join(out, "expected_index.json"), "w"), indent=1)
    print("written", out, len(index), "entries")


if __name__ == "__main__":
    main()
