"""Synthetic XR-Linear models / queries in the reference's on-disk format (no reference code needed to write them).

On-disk layout written here == what ``XLinearModel.save`` produces (pecos/xmc/xlinear/model.py:92-103,
pecos/xmc/base.py:807-830, :1371-1395): ``param.json`` + ``ranker/param.json`` + ``ranker/{d}.model/{param.json,W.npz,C.npz}``
with the matrices stored by ``scipy.sparse.save_npz(compressed=False)`` as CSC.

Workload generators follow SURVEY.md section 8(d) (configs E and S); seeds are fixed so that every engine sees the
same files.
"""
import json
import os

import numpy as np
import scipy.sparse as smat


# ------------------------------------------------------------------------------------------------ model writer
def save_xlinear_model(folder, layers, bias=1.0, only_topk=20, post_processor="l3-hinge", skip_root_C=False):
    """layers: list of (W, C) scipy sparse matrices, W: (D [+1]) x n_l, C: n_l x n_{l-1}; any format (stored as CSC)."""
    os.makedirs(folder, exist_ok=True)
    ranker = os.path.join(folder, "ranker")
    os.makedirs(ranker, exist_ok=True)
    depth = len(layers)
    W_last, C_last = layers[-1]
    nr_features = int(W_last.shape[0] - (1 if bias > 0 else 0))
    with open(os.path.join(folder, "param.json"), "w", encoding="utf-8") as f:
        f.write(json.dumps({"__meta__": {"class_fullname": "pecos.xmc.xlinear.model###XLinearModel"}, "model": "XLinearModel"}, indent=True))
    with open(os.path.join(ranker, "param.json"), "w", encoding="utf-8") as f:
        f.write(json.dumps({
            "__meta__": {"class_fullname": "pecos.xmc.base###HierarchicalMLModel"},
            "model": "HierarchicalMLModel",
            "depth": depth,
            "nr_features": nr_features,
            "nr_codes": int(C_last.shape[1]),
            "nr_labels": int(W_last.shape[1]),
        }, indent=True))
    for d, (W, C) in enumerate(layers):
        sub = os.path.join(ranker, f"{d}.model")
        os.makedirs(sub, exist_ok=True)
        pk = only_topk[d] if isinstance(only_topk, (list, tuple)) else only_topk
        pp = post_processor[d] if isinstance(post_processor, (list, tuple)) else post_processor
        with open(os.path.join(sub, "param.json"), "w", encoding="utf-8") as f:
            f.write(json.dumps({
                "__meta__": {"class_fullname": "pecos.xmc.base###MLModel"},
                "model": "MLModel",
                "nr_labels": int(W.shape[1]),
                "nr_features": nr_features,
                "nr_codes": int(C.shape[1]),
                "bias": float(bias),
                "pred_kwargs": {
                    "__meta__": {"class_fullname": "pecos.xmc.base###MLModel.PredParams"},
                    "only_topk": int(pk),
                    "post_processor": str(pp),
                },
            }, indent=True))
        Wc = smat.csc_matrix(W, dtype=np.float32)
        Wc.sort_indices()
        with open(os.path.join(sub, "W.npz"), "wb") as f:
            smat.save_npz(f, Wc, compressed=False)
        if not (d == 0 and skip_root_C):
            Cc = smat.csc_matrix(C, dtype=np.float32)
            with open(os.path.join(sub, "C.npz"), "wb") as f:
                smat.save_npz(f, Cc, compressed=False)
    return folder


# ------------------------------------------------------------------------------------------------ helpers
def _contiguous_codes(sizes):
    """C (n_child x n_parent) with children of parent p occupying a contiguous, ordered block."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n_child = int(sizes.sum())
    cols = np.repeat(np.arange(len(sizes)), sizes)
    return smat.csc_matrix((np.ones(n_child, dtype=np.float32), (np.arange(n_child), cols)), shape=(n_child, len(sizes)))


def _split_sizes(rng, total, parts, even):
    if even:
        base = total // parts
        sizes = np.full(parts, base, dtype=np.int64)
        sizes[: total - base * parts] += 1
        return sizes
    sizes = 1 + rng.multinomial(total - parts, np.full(parts, 1.0 / parts))
    return sizes.astype(np.int64)


def _sample_rows_uniform(rng, n_cols, n_rows, k):
    """k distinct row ids per column, uniform on [0, n_rows) -> (n_cols, k) sorted."""
    k = min(k, n_rows)
    if n_rows <= 4 * k or n_rows * n_cols <= 50_000_000:
        out = np.empty((n_cols, k), dtype=np.int64)
        step = max(1, 20_000_000 // max(n_rows, 1))
        for s in range(0, n_cols, step):
            e = min(n_cols, s + step)
            r = rng.random((e - s, n_rows), dtype=np.float32)
            out[s:e] = np.argpartition(r, k - 1, axis=1)[:, :k]
        out.sort(axis=1)
        return out
    return _sample_rows_weighted(rng, n_cols, None, n_rows, k)


def _sample_rows_weighted(rng, n_cols, cdf, n_rows, k):
    """k distinct ids per column drawn (approximately) from the popularity law given by `cdf` (None = uniform)."""
    out = np.empty((n_cols, k), dtype=np.int64)
    step = max(1, 4_000_000 // k)
    for s in range(0, n_cols, step):
        e = min(n_cols, s + step)
        n = e - s
        m = int(k * 1.6) + 16
        while True:
            u = rng.random((n, m))
            ids = (u * n_rows).astype(np.int64) if cdf is None else np.searchsorted(cdf, u).astype(np.int64)
            np.clip(ids, 0, n_rows - 1, out=ids)
            ids.sort(axis=1)
            dup = np.zeros_like(ids, dtype=bool)
            dup[:, 1:] = ids[:, 1:] == ids[:, :-1]
            n_unique = m - dup.sum(axis=1)
            if (n_unique >= k).all():
                break
            m = int(m * 1.5)
        # keep a random k-subset of the distinct ids of each column
        key = rng.random((n, m))
        key[dup] = 2.0
        sel = np.argpartition(key, k - 1, axis=1)[:, :k]
        picked = np.take_along_axis(ids, sel, axis=1)
        picked.sort(axis=1)
        out[s:e] = picked
    return out


def _random_weight_matrix(rng, n_cols, D, nnz_per_col, bias, popularity_cdf=None, w_std=0.5, b_std=0.1):
    k = min(nnz_per_col, D)
    rows = _sample_rows_uniform(rng, n_cols, D, k) if popularity_cdf is None else _sample_rows_weighted(rng, n_cols, popularity_cdf, D, k)
    vals = (rng.standard_normal((n_cols, k)) * w_std).astype(np.float32)
    vals[vals == 0] = np.float32(w_std)
    if bias > 0:
        rows = np.concatenate([rows, np.full((n_cols, 1), D, dtype=np.int64)], axis=1)
        bvals = (rng.standard_normal((n_cols, 1)) * b_std).astype(np.float32)
        bvals[bvals == 0] = np.float32(b_std)
        vals = np.concatenate([vals, bvals], axis=1)
    per = rows.shape[1]
    indptr = np.arange(0, (n_cols + 1) * per, per, dtype=np.int64)
    return smat.csc_matrix((vals.ravel(), rows.ravel(), indptr), shape=(D + (1 if bias > 0 else 0), n_cols))


def make_tree_model(seed, layer_sizes, D, nnz_per_col, bias=1.0, even=False, popularity_cdf=None):
    """Random-weight label tree with the given node count per layer (layer_sizes[-1] = number of labels)."""
    rng = np.random.default_rng(seed)
    layers = []
    prev = 1
    for d, n in enumerate(layer_sizes):
        sizes = _split_sizes(rng, n, prev, even)
        C = _contiguous_codes(sizes)
        W = _random_weight_matrix(np.random.default_rng(seed + 1 + d), n, D, nnz_per_col, bias, popularity_cdf)
        layers.append((W, C))
        prev = n
    return layers


def make_queries(seed, Q, D, nnz_per_row, popularity_cdf=None):
    rng = np.random.default_rng(seed)
    k = min(nnz_per_row, D)
    cols = _sample_rows_uniform(rng, Q, D, k) if popularity_cdf is None else _sample_rows_weighted(rng, Q, popularity_cdf, D, k)
    vals = np.abs(rng.standard_normal((Q, k))).astype(np.float32)
    vals[vals == 0] = np.float32(1.0)
    vals /= np.sqrt((vals.astype(np.float64) ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
    indptr = np.arange(0, (Q + 1) * k, k, dtype=np.int64)
    X = smat.csr_matrix((vals.ravel(), cols.ravel(), indptr), shape=(Q, D), dtype=np.float32)
    X.has_sorted_indices = True
    return X


def zipf_cdf(D, shift=100.0):
    p = 1.0 / (np.arange(D, dtype=np.float64) + shift)
    c = np.cumsum(p)
    return c / c[-1]


# ------------------------------------------------------------------------------------------------ named workloads
WORKLOADS = {
    # eurlex-4k-shaped (BASELINE.json configs[1])
    "eurlex-4k": dict(layer_sizes=[4, 64, 3956], D=5000, nnz_per_col=250, Q=15449, nnz_per_row=250, beam_size=10,
                      only_topk=10, zipf=False, model_seed=0, query_seed=2),
    # amazon-3M-scale synthetic tree (BASELINE.json configs[2])
    "synthetic-3m": dict(layer_sizes=[8, 64, 512, 4096, 32768, 3000000], D=500000, nnz_per_col=128, Q=100000,
                         nnz_per_row=128, beam_size=20, only_topk=10, zipf=True, model_seed=10, query_seed=20),
    # reduced copy of S for quick functional runs
    "synthetic-small": dict(layer_sizes=[8, 64, 512, 4096, 65536], D=50000, nnz_per_col=64, Q=4096, nnz_per_row=64,
                            beam_size=20, only_topk=10, zipf=True, model_seed=10, query_seed=20),
}


def build_workload(name, folder, scale_queries=None, only_topk_stored=20, post_processor="l3-hinge"):
    """Writes model `folder` (if absent) and returns (model_folder, X, cfg)."""
    cfg = dict(WORKLOADS[name])
    cdf = zipf_cdf(cfg["D"]) if cfg["zipf"] else None
    marker = os.path.join(folder, "ranker", "param.json")
    if not os.path.exists(marker):
        layers = make_tree_model(cfg["model_seed"], cfg["layer_sizes"], cfg["D"], cfg["nnz_per_col"], bias=1.0,
                                 even=cfg["zipf"], popularity_cdf=cdf)
        save_xlinear_model(folder, layers, bias=1.0, only_topk=only_topk_stored, post_processor=post_processor)
    Q = scale_queries if scale_queries else cfg["Q"]
    X = make_queries(cfg["query_seed"], Q, cfg["D"], cfg["nnz_per_row"], cdf)
    return folder, X, cfg
