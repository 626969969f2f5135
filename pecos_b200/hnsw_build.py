"""GPU construction of an HNSW index in the reference's on-disk format.

What it replaces: ``HNSW::train`` (pecos/core/ann/hnsw.hpp:677-846), the reference's incremental, lock-based CPU build
(44 s for 100k x 768, 797 s for 1M x 768 on 128 host threads) -- the step that kept BASELINE.json's 10M-vector configuration
out of reach.  What it produces: ``<folder>/param.json`` + ``<folder>/c_model/{config.json, index.mmap_store}``, byte-compatible
with what ``HNSW.save`` writes (hnsw.hpp:490-532, GraphL0 :104-120, GraphL1 :188-219, container pecos/core/utils/mmap_util.hpp),
so BOTH the reference library and pecos_b200 load and search it.

Parity contract: the BUILD is recall-level -- a batch construction cannot reproduce the insertion-order
dependent graph of the incremental algorithm (the reference itself is nondeterministic with threads > 1, hnsw.hpp:804-809);
SEARCH on the saved file is bit-level: the reference and the CUDA engine return identical ids / distance bits on it
(tests/test_hnsw_build_gpu.py).

Algorithm (B200-first: the distance work is dense GEMMs on the tensor cores, cuBLAS through torch -- a plain library GEMM --
instead of 10^9 dependent single-vector distance calls):

1. node levels as the reference draws them: ``floor(-ln(U) / ln(M))`` (hnsw.hpp:785-793), entry point = first node of the top level;
2. for every level l and node i: the EXACT k = efC nearest among the nodes present at l that the incremental algorithm would
   already have inserted (ids < i; hnsw.hpp:804-809 inserts in id order), by tiled brute force (``X_tile @ X^T`` over the lower
   triangle + running top-k) -- the prefix constraint gives early nodes their long-range links, i.e. the navigability of the
   incrementally built graph;
3. the reference's neighbour-selection heuristic (hnsw.hpp:556-592: keep a candidate iff it is closer to the node than to
   every neighbour kept so far; at most M; fewer than M candidates are kept whole), evaluated for a whole tile of nodes at
   once from the candidates' pairwise distance matrix (one batched GEMM);
4. reverse links: every selected edge u -> v also offers u to v; a node whose selected + offered set exceeds the level's
   capacity (maxM0 = 2M at level 0, maxM above) is pruned with the same heuristic (hnsw.hpp:628-652);
5. neighbour lists sorted by ascending distance (hnsw.hpp:823-845), records laid out as GraphL0 / GraphL1 store them.

torch is used for device memory and the GEMM / top-k primitives; nothing here is on the search path.
"""
import json
import math
import os

import numpy as np

_HNSW_T = {
    "ip": "pecos::ann::HNSW<float, pecos::ann::FeatVecDenseIPSimd<float>>",
    "l2": "pecos::ann::HNSW<float, pecos::ann::FeatVecDenseL2Simd<float>>",
}


# ------------------------------------------------------------------------------------------------ container writer
def write_mmap_store(path, blocks):
    """PECOS MmapStore container (pecos/core/utils/mmap_util.hpp:54-184): data blocks, each padded to 16-byte alignment,
    then the metadata ``[n_blocks u64][(offset u64, size u64) x n]``, then the 16-byte signature
    ``0x93 'PECOS' | '<' | version 1 | metadata offset u64``.  ``blocks``: list of bytes-like / numpy arrays."""
    info = []
    with open(path, "wb") as f:
        off = 0
        for b in blocks:
            raw = b.tobytes() if isinstance(b, np.ndarray) else bytes(b)
            info.append((off, len(raw)))
            f.write(raw)
            off += len(raw)
            pad = (-off) % 16
            if pad:
                f.write(b"\0" * pad)
                off += pad
        meta = np.array([len(info)] + [v for pair in info for v in pair], dtype="<u8").tobytes()
        f.write(meta)
        f.write(b"\x93PECOS" + b"<" + bytes([1]) + np.array([off], dtype="<u8").tobytes())


def _scalar(v, dtype="<u4"):
    return np.array([v], dtype=dtype)


def _vector(arr):
    """MmapableVector = two blocks: size u64, then the elements (mmap_util.hpp:526-537)."""
    arr = np.ascontiguousarray(arr)
    return [np.array([arr.size], dtype="<u8"), arr]


# ------------------------------------------------------------------------------------------------ distance helpers
def _pairwise(torch, A, B, metric, a_sq=None, b_sq=None):
    """distance(A_i, B_j): ip -> 1 - <a, b>; l2 -> |a|^2 + |b|^2 - 2<a, b> (clamped at 0)."""
    G = A @ B.transpose(-1, -2)
    if metric == "ip":
        return 1.0 - G
    if a_sq is None:
        a_sq = (A * A).sum(-1)
    if b_sq is None:
        b_sq = (B * B).sum(-1)
    return (a_sq.unsqueeze(-1) + b_sq.unsqueeze(-2) - 2.0 * G).clamp_min_(0.0)


def _exact_knn(torch, X, ids, k, metric, q_tile, c_tile):
    """For every node of `ids` (LongTensor, ascending = the reference's insertion order): its k nearest EARLIER nodes of `ids`
    -- what an incremental insertion can link a new node to (hnsw.hpp:742-760: the graph only holds the nodes inserted so
    far).  This prefix constraint is what makes the graph navigable: early nodes get long-range links, exactly as in the
    incremental algorithm (an unconstrained kNN graph has none: recall 0.89 at N = 20k, measured).  Returns (nbr positions
    into `ids` [n, k], distances [n, k]) ascending; missing slots hold -1 / inf."""
    n = ids.numel()
    k = min(k, max(n - 1, 0))
    dev = X.device
    out_i = torch.full((n, max(k, 1)), -1, dtype=torch.long, device=dev)
    out_d = torch.full((n, max(k, 1)), float("inf"), dtype=torch.float32, device=dev)
    if k == 0:
        return out_i[:, :0], out_d[:, :0]
    sq = (X * X).sum(-1) if metric == "l2" else None
    for q0 in range(0, n, q_tile):
        q1 = min(n, q0 + q_tile)
        qi = ids[q0:q1]
        A = X[qi]
        best_d = torch.full((q1 - q0, k), float("inf"), dtype=torch.float32, device=dev)
        best_i = torch.full((q1 - q0, k), -1, dtype=torch.long, device=dev)
        for c0 in range(0, q1, c_tile):                      # only earlier nodes can be candidates
            c1 = min(q1, c0 + c_tile)
            ci = ids[c0:c1]
            D = _pairwise(torch, A, X[ci], metric, None if sq is None else sq[qi], None if sq is None else sq[ci])
            if c1 > q0:  # the tile reaches into the query rows' own range: keep strictly earlier positions only
                rows = torch.arange(q0, q1, device=dev).unsqueeze(1)
                cols = torch.arange(c0, c1, device=dev).unsqueeze(0)
                D = torch.where(cols < rows, D, torch.full_like(D, float("inf")))
            cat_d = torch.cat([best_d, D], dim=1)
            cat_i = torch.cat([best_i, torch.arange(c0, c1, device=dev).expand(q1 - q0, -1)], dim=1)
            best_d, sel = torch.topk(cat_d, k, dim=1, largest=False, sorted=True)
            best_i = torch.gather(cat_i, 1, sel)
        out_d[q0:q1, :k] = best_d
        out_i[q0:q1, :k] = torch.where(torch.isinf(best_d), torch.full_like(best_i, -1), best_i)
    return out_i[:, :k], out_d[:, :k]


def _heuristic(torch, X, node_pos, cand, cand_d, cap, metric, tile):
    """The reference's get_neighbors_heuristic (hnsw.hpp:556-592) for many nodes at once.
    cand [n, C]: candidate ids (global), ascending by cand_d (distance to the node), -1 = empty.  Returns kept mask [n, C]:
    a node with fewer than `cap` candidates keeps all of them; else candidates are visited in order and kept iff no
    already-kept candidate is strictly closer to them than the node is, until `cap` are kept."""
    n, C = cand.shape
    dev = X.device
    kept_all = torch.zeros((n, C), dtype=torch.bool, device=dev)
    valid_all = cand >= 0
    for t0 in range(0, n, tile):
        t1 = min(n, t0 + tile)
        c = cand[t0:t1]
        valid = valid_all[t0:t1]
        dq = cand_d[t0:t1]
        V = X[c.clamp_min(0)]                                   # [T, C, d]
        D = _pairwise(torch, V, V, metric)                      # [T, C, C] candidate-to-candidate distances
        few = valid.sum(1) < cap
        kept = torch.zeros_like(valid)
        count = torch.zeros(t1 - t0, dtype=torch.long, device=dev)
        for j in range(C):
            bad = ((D[:, :, j] < dq[:, j:j + 1]) & kept).any(dim=1)
            ok = valid[:, j] & ~bad & (count < cap)
            kept[:, j] = ok
            count += ok.long()
        kept_all[t0:t1] = torch.where(few.unsqueeze(1), valid, kept)
    return kept_all


def _build_level(torch, X, ids, M, cap, efC, metric, q_tile, c_tile, h_tile):
    """Neighbour lists (global ids, ascending distance, <= cap each) of the nodes `ids` on one level."""
    n = ids.numel()
    dev = X.device
    lists = torch.full((n, cap), -1, dtype=torch.long, device=dev)
    if n <= 1:
        return lists, torch.zeros(n, dtype=torch.long, device=dev)
    pos, dist = _exact_knn(torch, X, ids, efC, metric, q_tile, c_tile)          # positions into ids
    cand = torch.where(pos >= 0, ids[pos.clamp_min(0)], torch.full_like(pos, -1))
    keep = _heuristic(torch, X, ids, cand, dist, M, metric, h_tile)            # forward selection: at most M (hnsw.hpp:598)
    # edges u -> v (selected) and the offers v <- u
    src = torch.arange(n, device=dev).unsqueeze(1).expand_as(cand)[keep]       # positions
    dst_pos = pos[keep]
    d_uv = dist[keep]
    # every node's pool = its selected + the nodes that selected it; dedupe (u, v) pairs
    a = torch.cat([src, dst_pos])
    b = torch.cat([dst_pos, src])
    d = torch.cat([d_uv, d_uv])
    key = a * n + b
    order = torch.argsort(key, stable=True)
    key, a, b, d = key[order], a[order], b[order], d[order]
    first = torch.ones_like(key, dtype=torch.bool)
    first[1:] = key[1:] != key[:-1]
    a, b, d = a[first], b[first], d[first]
    # per node: pool sorted by distance, truncated to a working width (the heuristic only ever needs the closest few)
    width = min(max(4 * cap, 64), 512)
    order = torch.argsort(d, stable=True)
    a, b, d = a[order], b[order], d[order]
    order = torch.argsort(a, stable=True)
    a, b, d = a[order], b[order], d[order]
    counts = torch.bincount(a, minlength=n)
    starts = torch.cumsum(counts, 0) - counts
    rank = torch.arange(a.numel(), device=dev) - starts[a]
    ok = rank < width
    pool = torch.full((n, width), -1, dtype=torch.long, device=dev)
    pool_d = torch.full((n, width), float("inf"), dtype=torch.float32, device=dev)
    pool[a[ok], rank[ok]] = ids[b[ok]]
    pool_d[a[ok], rank[ok]] = d[ok]
    over = counts > cap
    keep2 = pool >= 0
    if bool(over.any()):
        idx = torch.nonzero(over).squeeze(1)
        keep2[idx] = _heuristic(torch, X, ids[idx], pool[idx], pool_d[idx], cap, metric, h_tile)
    # compact the kept neighbours (already ascending by distance)
    rank2 = torch.cumsum(keep2.long(), 1) - 1
    rows = torch.arange(n, device=dev).unsqueeze(1).expand_as(pool)
    sel = keep2 & (rank2 < cap)
    lists[rows[sel], rank2[sel]] = pool[sel]
    return lists, sel.sum(1)


# ------------------------------------------------------------------------------------------------ public entry point
def build_hnsw_index(X, folder, M=32, efC=100, metric="ip", seed=0, max_level_upper_bound=-1, device=None, pred_kwargs=None,
                     q_tile=4096, c_tile=65536, h_tile=None, allow_tf32=False):
    """Builds the index for the rows of ``X`` (float32 [N, d]) and writes it to ``folder`` in the reference's format.
    Returns a dict with the build statistics.  ``device``: torch device (default: cuda:0; "cpu" is accepted for tiny inputs,
    e.g. format tests on a box without a GPU)."""
    import torch

    if metric not in _HNSW_T:
        raise ValueError(f"metric must be 'ip' or 'l2', got {metric!r}")
    X = np.ascontiguousarray(X, dtype=np.float32)
    N, d = X.shape
    if N < 1:
        raise ValueError("empty input")
    dev = torch.device(device if device is not None else "cuda:0")
    if dev.type == "cuda":
        torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    maxM, maxM0 = int(M), 2 * int(M)
    if h_tile is None:
        h_tile = max(16, min(2048, (256 << 20) // (4 * max(4 * maxM0, 64) * max(4 * maxM0, 64, d))))

    # 1. levels (hnsw.hpp:785-793) and the entry point
    rng = np.random.default_rng(seed)
    u = 1.0 - rng.random(N)  # (0, 1]
    levels = np.floor(-np.log(u) * (1.0 / math.log(float(maxM)))).astype(np.int64)
    if max_level_upper_bound >= 0:
        levels = np.minimum(levels, int(max_level_upper_bound))
    max_level = int(levels.max())
    init_node = int(np.argmax(levels == max_level))

    Xd = torch.from_numpy(X).to(dev)
    lvl = torch.from_numpy(levels).to(dev)
    level_lists = []
    for l in range(0, max_level + 1):
        ids = torch.nonzero(lvl >= l).squeeze(1)
        cap = maxM0 if l == 0 else maxM
        lists, deg = _build_level(torch, Xd, ids, maxM, cap, int(efC), metric, q_tile, c_tile, h_tile)
        level_lists.append((ids.cpu().numpy(), lists.cpu().numpy(), deg.cpu().numpy()))

    # 2. records.  GraphL0: per node [deg u32][maxM0 ids u32][len u32][d f32] (hnsw.hpp:47-91, :104-178)
    rec = 4 * (1 + maxM0) + 4 + 4 * d
    l0 = np.zeros((N, rec), dtype=np.uint8)
    ids0, lists0, deg0 = level_lists[0]
    head = np.zeros((N, 1 + maxM0), dtype="<u4")
    head[ids0, 0] = deg0
    nb = np.where(lists0 >= 0, lists0, 0).astype("<u4")
    head[ids0, 1:] = nb
    l0[:, : 4 * (1 + maxM0)] = head.view(np.uint8)
    l0[:, 4 * (1 + maxM0): 4 * (1 + maxM0) + 4] = np.full((N, 1), d, dtype="<u4").view(np.uint8)
    l0[:, 4 * (1 + maxM0) + 4:] = X.view(np.uint8).reshape(N, 4 * d)
    mem_start = (np.arange(N + 1, dtype="<u8") * rec)
    # GraphL1: per node max_level slots of [deg u32][maxM ids u32] (hnsw.hpp:188-219); every node gets all slots
    level_mem = 1 + maxM
    node_mem = max_level * level_mem
    l1 = np.zeros(N * node_mem, dtype="<u4")
    for l in range(1, max_level + 1):
        ids_l, lists_l, deg_l = level_lists[l]
        base = ids_l.astype(np.int64) * node_mem + (l - 1) * level_mem
        l1[base] = deg_l
        cols = np.where(lists_l >= 0, lists_l, 0).astype("<u4")
        l1[(base[:, None] + 1 + np.arange(maxM)[None, :]).ravel()] = cols.ravel()

    # 3. files
    c_model = os.path.join(folder, "c_model")
    os.makedirs(c_model, exist_ok=True)
    blocks = [_scalar(N), _scalar(maxM), _scalar(maxM0), _scalar(int(efC)), _scalar(max_level), _scalar(init_node),
              _scalar(N), _scalar(d), _scalar(maxM0), _scalar(rec)] + _vector(mem_start) + _vector(l0.reshape(-1)) + \
             [_scalar(N), _scalar(max_level), _scalar(maxM), _scalar(node_mem), _scalar(level_mem)] + _vector(l1)
    write_mmap_store(os.path.join(c_model, "index.mmap_store"), blocks)
    with open(os.path.join(c_model, "config.json"), "w", encoding="utf-8") as f:
        json.dump({"hnsw_t": _HNSW_T[metric], "version": "v2.0",
                   "train_params": {"num_node": N, "maxM": maxM, "maxM0": maxM0, "efC": int(efC), "max_level": max_level,
                                    "init_node": init_node}}, f, indent=4)
    pk = {"efS": 100, "topk": 10, "threads": 1}
    pk.update(pred_kwargs or {})
    with open(os.path.join(folder, "param.json"), "w", encoding="utf-8") as f:
        json.dump({"model": "HNSW", "data_type": "drm", "metric_type": metric, "num_item": N, "feat_dim": d,
                   "train_kwargs": {"M": maxM, "efC": int(efC), "builder": "pecos_b200.hnsw_build (batch, exact kNN + heuristic)"},
                   "pred_kwargs": pk}, f, indent=1)
    return {"num_node": N, "feat_dim": d, "max_level": max_level, "init_node": init_node,
            "mean_degree_l0": float(deg0.mean()), "nodes_per_level": [int(t[0].size) for t in level_lists]}
