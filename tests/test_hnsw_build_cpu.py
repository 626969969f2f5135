"""CPU tests of the GPU HNSW builder's HOST logic and file format (pecos_b200/hnsw_build.py run with device="cpu" on tiny inputs;
the GPU run is tests/test_hnsw_build_gpu.py).  The written index must be loadable by the REFERENCE library (oracle/_ref) and by
the restatement, searches on it must agree bit for bit, the graph must respect the reference's structural invariants, and its
recall must be at the level of an index trained by the reference itself."""
import json
import os

import numpy as np
import pytest


def _unit(rng, n, d):
    X = rng.standard_normal((n, d)).astype(np.float32)
    return X / np.linalg.norm(X, axis=1, keepdims=True)


def _recall(idx, exact):
    return float(np.mean([len(set(idx[i]) & set(exact[i])) / exact.shape[1] for i in range(idx.shape[0])]))


@pytest.mark.parametrize("N,d,M,efC,metric", [(1500, 24, 12, 50, "ip"), (1200, 17, 6, 30, "l2"), (40, 5, 4, 10, "l2")])
def test_built_index_loads_everywhere_and_recalls(tmp_path, built, have_ref, N, d, M, efC, metric):
    from oracle import restatement
    from pecos_b200.hnsw_build import build_hnsw_index

    rng = np.random.default_rng(N)
    X, Q = _unit(rng, N, d), _unit(rng, 100, d)
    folder = str(tmp_path / "idx")
    stats = build_hnsw_index(X, folder, M=M, efC=efC, metric=metric, seed=3, device="cpu")
    assert stats["num_node"] == N and stats["nodes_per_level"][0] == N
    cfg = json.load(open(os.path.join(folder, "c_model", "config.json")))
    assert cfg["version"] == "v2.0" and cfg["train_params"]["maxM0"] == 2 * M
    o = restatement.OracleHNSW(folder, isa=0)
    # structural invariants of the reference's records (hnsw.hpp:47-220): degrees within capacity, no self loops, no
    # duplicates, neighbours sorted by ascending distance, vectors stored verbatim
    assert np.array_equal(o.vectors(), X)
    rec = o.l0.reshape(N, o.l0_node_mem)
    head = np.ascontiguousarray(rec[:, : 4 * (1 + 2 * M)]).view(np.uint32).reshape(N, 1 + 2 * M)
    assert head[:, 0].max() <= 2 * M and head[:, 0].min() >= (1 if N > 1 else 0)
    for u in range(0, N, max(1, N // 50)):
        nb = head[u, 1: 1 + head[u, 0]]
        assert u not in nb and len(set(nb.tolist())) == nb.size
        dist = 1 - X[nb] @ X[u] if metric == "ip" else ((X[nb] - X[u]) ** 2).sum(1)
        assert np.all(np.diff(dist) >= -1e-5)
    oi, od = o.predict(Q, 80, 10)
    exact = np.argsort((1 - Q @ X.T) if metric == "ip" else ((Q[:, None, :] - X[None, :, :]) ** 2).sum(-1), axis=1)[:, :10]
    assert _recall(oi, exact) >= 0.97
    if have_ref:
        from oracle import ref

        r = ref.RefHNSW.load(os.path.join(folder, "c_model"), metric)  # the REFERENCE loads the file we wrote
        ri, rd = r.predict(Q, 80, 10, threads=1)
        assert np.array_equal(ri, oi) and np.array_equal(rd.view(np.uint32), od.view(np.uint32))
        trained = ref.RefHNSW.train(X, M=M, efC=efC, metric=metric, threads=1)
        ti, _ = trained.predict(Q, 80, 10, threads=1)
        assert _recall(ri, exact) >= _recall(ti, exact) - 0.02  # recall-level parity with the reference's own build


def test_mmap_store_writer_round_trip(tmp_path):
    from oracle.restatement import read_mmap_store
    from pecos_b200.hnsw_build import write_mmap_store

    blocks = [np.arange(5, dtype="<u4"), np.array([7], dtype="<u8"), np.arange(33, dtype=np.uint8), np.zeros(0, dtype=np.float32)]
    p = str(tmp_path / "x.mmap_store")
    write_mmap_store(p, blocks)
    got = read_mmap_store(p)
    assert len(got) == 4 and all(np.array_equal(g, b.view(np.uint8).reshape(-1)) for g, b in zip(got, blocks))
    raw = np.fromfile(p, dtype=np.uint8)
    assert bytes(raw[-16:-10]) == b"\x93PECOS" and raw.size % 8 == 0
