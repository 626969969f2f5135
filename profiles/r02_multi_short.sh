#!/bin/bash
# Short multi-GPU check of the final round-2 code (through `gpurun --gpus N`): usage  bash profiles/r02_multi_short.sh <tag> <N>
#   1. bench.py under torchrun at N ranks (headline + secondary synthetic-3m strong-scaled + index-sharded)
#   2. bench.py --workload hnsw-sparse-100k under torchrun at N ranks (replicas)
#   3. the fan-out tests (PB200_DEVICES) incl. real devices 0..N-1 for one HNSW call
tag=${1:-r02_u}; N=${2:-2}
o=gpurun_out; mkdir -p $o
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > $o/${tag}_bench_n$N.json 2> $o/${tag}_bench_n$N.err || tail -5 $o/${tag}_bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --workload hnsw-sparse-100k --gpus $N --steps 5 --warmup 3 > $o/${tag}_bench_hnsw_sparse100k_n$N.json 2> $o/${tag}_bench_hnsw_sparse100k_n$N.err || tail -5 $o/${tag}_bench_hnsw_sparse100k_n$N.err
python -m pytest tests/test_fanout_gpu.py -x -q -m gpu > $o/${tag}_gpu_tests_fanout.log 2>&1; tail -3 $o/${tag}_gpu_tests_fanout.log
python - <<PY
import json
for f in ("$o/${tag}_bench_n$N.json", "$o/${tag}_bench_hnsw_sparse100k_n$N.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    s = d.get("secondary", {})
    print(f.split("/")[-1], "value", round(d["value"]), "n_gpus", d["n_gpus"], "e2e", round(d["e2e"]["value"]), "parity", d.get("parity", {}).get("ids_bit_equal"),
          "| secondary:", {k: (round(v["value"]) if isinstance(v, dict) and "value" in v else v) for k, v in s.items()})
PY
