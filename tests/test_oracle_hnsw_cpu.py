"""CPU tests that PIN THE HNSW ORACLE: the plain-C restatement (oracle/hnsw_oracle.c) must reproduce, bit for bit, the
reference-recorded golden search results (ids, order, distance bits) of

* tests/golden/hnsw_toy  -- the reference's own 90-point, d = 2 fixture index (+ two indices trained by the reference), and
* tests/golden/hnsw_mid  -- reference-built indices with d in {768, 128, 70, 20}, both metrics, a six-fold duplicated index
  (exact ties), efS up to 600 (tests/golden/make_golden_hnsw.py).

The goldens were produced by the avx512f clone of the reference's distance kernels (provenance.json), restated as isa=0.
"""
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("suite", ["hnsw_toy", "hnsw_mid"])
def test_hnsw_restatement_reproduces_reference_goldens(built, suite):
    from oracle import restatement

    gold = os.path.join(HERE, "golden", suite)
    assert json.load(open(os.path.join(gold, "provenance.json")))["distance_isa_clone"] == "avx512f"
    E = np.load(os.path.join(gold, "expected.npz"))
    index = json.load(open(os.path.join(gold, "expected_index.json")))
    models, n = {}, 0
    for it in index:
        folder = os.path.join(gold, it["model"])
        m = models.get(it["model"]) or models.setdefault(it["model"], restatement.OracleHNSW(folder, isa=0))
        Q = np.load(os.path.join(folder, "Q.npy")) if suite == "hnsw_mid" else np.load(os.path.join(gold, "X.tst.npy"))
        idx, dist = m.predict(Q, it["efS"], it["topk"])
        assert np.array_equal(idx, E[it["key"] + "|idx"]), it["key"]
        assert np.array_equal(dist.view(np.uint32), E[it["key"] + "|dist"].view(np.uint32)), it["key"]
        n += 1
    assert n >= 20


def test_hnsw_mid_goldens_cover_the_simd_paths():
    """The fixture set must keep exercising: full 16-lane blocks (d >= 16), the 4-wide remainder and the scalar tail
    (d = 70 = 64 + 4 + 2), a large d (768), both metrics, and exact ties."""
    prov = json.load(open(os.path.join(HERE, "golden", "hnsw_mid", "provenance.json")))
    dims = {c["d"] for c in prov["cases"]}
    assert {768, 128, 70} <= dims
    assert {c["metric"] for c in prov["cases"]} == {"ip", "l2"}
    assert any(c["dup"] > 1 for c in prov["cases"])
    assert max(e for e, _ in prov["grid"]) >= 600
