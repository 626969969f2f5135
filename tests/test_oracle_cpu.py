"""CPU tests that PIN THE ORACLE: our plain-C restatement vs (a) the committed golden vectors produced by the reference
(tests/golden/make_golden.py) and (b) oracle/_ref -- the reference library itself -- when it is present.

These are the tests behind the "parity pinned" statement in oracle/xlinear_oracle.c and DESIGN.md.
"""
import json
import os

import numpy as np
import pytest
import scipy.sparse as smat

from pecos_b200 import synth

from .util import assert_csr_parity, random_tree

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "xlinear_toy")


def _expected(E, key):
    return E[key + "|indptr"], E[key + "|indices"], E[key + "|data"]


@pytest.fixture(scope="module")
def golden(built):
    E = np.load(os.path.join(GOLD, "expected.npz"))
    index = json.load(open(os.path.join(GOLD, "expected_index.json")))
    Xt = smat.load_npz(os.path.join(GOLD, "Xt.npz")).tocsr().astype(np.float32)
    Xt.sort_indices()
    return E, index, Xt


def test_restatement_reproduces_reference_golden_vectors(golden):
    """Every (model, post-processor, beam, topk, csr|drm) entry recorded from the reference: ids, ranks AND score bits."""
    from oracle import restatement

    E, index, Xt = golden
    models = {}
    n = 0
    for item in index:
        name = item["model"]
        if name not in models:
            models[name] = restatement.OracleXLinear(os.path.join(GOLD, name, "ranker"))
        Xq = Xt if item["kind"] == "csr" else np.ascontiguousarray(Xt.toarray())
        got = models[name].predict(Xq, item["beam_size"] or 0, item["post_processor"], item["only_topk"] or 0)
        indptr, indices, data = _expected(E, item["key"])
        want = smat.csr_matrix((data, indices, indptr), shape=tuple(item["shape"]))
        exact = assert_csr_parity(got, want, rtol=0.0, what=item["key"])
        assert exact == 1.0
        n += 1
    assert n == len(index) and n >= 80


def test_default_prediction_matches_the_reference_repo_golden_file(golden):
    """tests/golden/xlinear_toy/Yt_pred_reference_golden.npy is the reference's own test/tst-data/xmc/xlinear/Yt_pred.npz
    (test_xlinear.py:366-370 compares with abs=1e-6)."""
    from oracle import restatement

    E, index, Xt = golden
    gold = np.load(os.path.join(GOLD, "Yt_pred_reference_golden.npy"))
    got = restatement.OracleXLinear(os.path.join(GOLD, "model", "ranker")).predict(Xt).toarray()
    assert np.abs(got - gold).max() <= 1e-6


@pytest.mark.parametrize("permute,prune,bias", [(False, 0.0, 1.0), (True, 0.0, 1.0), (True, 0.25, 1.0), (False, 0.0, -1.0)])
def test_restatement_equals_reference_library_on_random_trees(tmp_path, built, have_ref, permute, prune, bias):
    if not have_ref:
        pytest.skip("oracle/_ref not built (reference sources absent)")
    from oracle import ref, restatement

    folder = str(tmp_path / "m")
    layers = random_tree(5, [3, 18, 160], 220, 18, bias=bias, permute=permute, prune=prune)
    synth.save_xlinear_model(folder, layers, bias=bias, only_topk=6)
    X = synth.make_queries(6, 40, 220, 25)
    r = ref.RefXLinear(os.path.join(folder, "ranker"))
    o = restatement.OracleXLinear(os.path.join(folder, "ranker"))
    for pp in [None, "noop", "sigmoid", "log-sigmoid", "l2-hinge", "log-l3-hinge"]:
        for beam, topk in [(0, 0), (2, 3), (8, 8), (30, 200)]:
            a, b = r.predict(X, beam, pp, topk), o.predict(X, beam, pp, topk)
            assert assert_csr_parity(b, a, rtol=0.0, what=f"csr {pp} {beam} {topk}") == 1.0
        a, b = r.predict(X.toarray(), 4, pp, 5), o.predict(X.toarray(), 4, pp, 5)
        assert assert_csr_parity(b, a, rtol=0.0, what=f"drm {pp}") == 1.0


def test_reference_library_layer_types_agree(tmp_path, built, have_ref):
    """BINARY_SEARCH_CHUNKED / HASH_CHUNKED / CSC return the same ids (reference test_xlinear.py:179-187); this is why
    the GPU engine serves all three requests from one layout."""
    if not have_ref:
        pytest.skip("oracle/_ref not built")
    from oracle import ref

    folder = str(tmp_path / "m")
    synth.save_xlinear_model(folder, random_tree(9, [4, 30, 250], 300, 20), bias=1.0, only_topk=5)
    X = synth.make_queries(10, 30, 300, 30)
    base = ref.RefXLinear(os.path.join(folder, "ranker"), "BINARY_SEARCH_CHUNKED").predict(X, 5, None, 5)
    for t in ("HASH_CHUNKED", "CSC"):
        other = ref.RefXLinear(os.path.join(folder, "ranker"), t).predict(X, 5, None, 5)
        assert np.array_equal(base.indices, other.indices)
        assert np.allclose(base.data, other.data, atol=1e-6)
