#!/usr/bin/env python
"""Records the REFERENCE library's result for a 2,000-query slice of BASELINE.json configs[2] (synthetic 3M-label tree, beam 20,
top-10; generator pecos_b200/synth.py with fixed seeds) into tests/golden/synthetic3m_slice/expected.npz.

Runs HERE (CPU container, needs oracle/_ref; ~10 min: the reference builds its chunked layout single-threaded at load).  The GPU
test regenerates the same model from the same seeds on the GPU box and compares the CUDA path with this file -- a parity pin
at the full size of the 3M-label configuration that does not need oracle/_ref at test time."""
import os
import sys
import tempfile
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)


def main():
    from oracle import ref
    from pecos_b200 import synth

    folder = os.environ.get("PB200_S_FOLDER") or os.path.join(tempfile.gettempdir(), "pecos_b200_bench", "synthetic-3m")
    t0 = time.time()
    synth.build_workload("synthetic-3m", folder, scale_queries=8)
    cfg = synth.WORKLOADS["synthetic-3m"]
    X = synth.make_queries(cfg["query_seed"], 2000, cfg["D"], cfg["nnz_per_row"], synth.zipf_cdf(cfg["D"]))
    print("model + queries ready", round(time.time() - t0, 1), "s")
    m = ref.RefXLinear(os.path.join(folder, "ranker"))
    print("reference loaded", round(time.time() - t0, 1), "s")
    P = m.predict(X, cfg["beam_size"], None, cfg["only_topk"], -1)
    out = os.path.join(HERE, "synthetic3m_slice")
    os.makedirs(out, exist_ok=True)
    np.savez_compressed(os.path.join(out, "expected.npz"), indptr=P.indptr.astype(np.int64), indices=P.indices.astype(np.uint32),
                        data=P.data.astype(np.float32), query_rows=np.int64(2000), query_seed=np.int64(cfg["query_seed"]))
    print("written", out, P.shape, P.nnz, round(time.time() - t0, 1), "s")


if __name__ == "__main__":
    main()
