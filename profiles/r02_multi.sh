#!/bin/bash
# Round-2 multi-GPU session (through `gpurun --gpus N`): usage  bash profiles/r02_multi.sh <tag> <N>
#   1. bench.py under torchrun at N ranks: headline eurlex-4k (weak, replicas) + secondary synthetic-3m STRONG-scaled (100k queries
#      split by nnz) + synthetic-3m-sharded (leaf layer split over the ranks, ONE NCCL all-gather, bit-identical to unsharded)
#   2. tests/dist_index_shard_check.py (NCCL index-sharding check on a smaller tree)
#   3. in-library fan-out: ONE process, PB200_DEVICES=all, the C-ABI call splits its rows over the N GPUs
tag=${1:-r02_n}; N=${2:-2}
o=gpurun_out; mkdir -p $o
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $o/${tag}_bench_n$N.json 2> $o/${tag}_bench_n$N.err || tail -5 $o/${tag}_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tests/dist_index_shard_check.py > $o/${tag}_index_shard_check_n$N.json 2> $o/${tag}_index_shard_check_n$N.err || tail -5 $o/${tag}_index_shard_check_n$N.err
PB200_DEVICES=all python - > $o/${tag}_fanout_n$N.json 2> $o/${tag}_fanout_n$N.err <<'PY'
import json, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
from pecos_b200 import core, synth
from pecos_b200.xlinear import XLinearModel
lib = core.get_clib(); c = lib.clib_float32
out = {"check": "in-library query fan-out (one process, PB200_DEVICES=all)", "devices": int(c.pb200_device_count())}
for wl, steps in (("eurlex-4k", 10), ("synthetic-3m", 3)):
    folder = os.path.join(os.environ.get("PB200_BENCH_CACHE", "/tmp/pecos_b200_bench"), wl)
    synth.build_workload(wl, folder, scale_queries=8)
    cfg = synth.WORKLOADS[wl]
    X = synth.make_queries(cfg["query_seed"], cfg["Q"], cfg["D"], cfg["nnz_per_row"], synth.zipf_cdf(cfg["D"]) if cfg["zipf"] else None)
    m = XLinearModel.load(folder, is_predict_only=True)
    reps = int(c.pb200_xlinear_replicas(m.model.model_chain))
    got = m.predict(X, beam_size=cfg["beam_size"], only_topk=cfg["only_topk"])
    t = []
    for _ in range(steps):
        t0 = time.perf_counter(); got = m.predict(X, beam_size=cfg["beam_size"], only_topk=cfg["only_topk"]); t.append(time.perf_counter() - t0)
    os.environ.pop("PB200_DEVICES")
    single = XLinearModel.load(folder, is_predict_only=True)
    want = single.predict(X[:4096], beam_size=cfg["beam_size"], only_topk=cfg["only_topk"])
    t1 = []
    for _ in range(steps):
        t0 = time.perf_counter(); single.predict(X, beam_size=cfg["beam_size"], only_topk=cfg["only_topk"]); t1.append(time.perf_counter() - t0)
    os.environ["PB200_DEVICES"] = "all"
    g = got[:4096]
    ok = bool(np.array_equal(g.indptr, want.indptr) and np.array_equal(g.indices, want.indices) and np.array_equal(g.data.view(np.uint32), want.data.view(np.uint32)))
    out[wl] = {"replicas": reps, "queries": int(X.shape[0]), "bit_identical_to_one_gpu": ok,
               "fanout_qps_python_api_pageable": X.shape[0] / min(t), "one_gpu_qps_python_api_pageable": X.shape[0] / min(t1)}
    del m, single
print(json.dumps(out))
PY
python - <<PY
import json
for f in ("$o/${tag}_bench_n$N.json", "$o/${tag}_index_shard_check_n$N.json", "$o/${tag}_fanout_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if "secondary" in d:
        s = d["secondary"]
        print(f.split("/")[-1], "E value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "| S:", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in s.items()})
    else:
        print(f.split("/")[-1], json.dumps(d)[:600])
PY
