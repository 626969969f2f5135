#!/bin/bash
# First GPU call of round 2 (run through gpurun from the repo root): validate what was written blind at the end of round 1
# and measure the experimental kernel, in one box lease.  Outputs go to gpurun_out/<tag>_*.
#   1. the validated suite (must stay green)
#   2. opt-in tests: single-layer entry points (python chain) and the chunk-major score kernel
#   3. S bench, default kernels vs PB200_XL_KERNEL_MODE=5 (chunk-major on layers 0-4), per-layer kernel times
#   4. one full ncu capture of xl_cm_scores_kernel (source view) if it survived step 2
tag=${1:-r02_a}
o=gpurun_out
python -m pytest tests -x -q -m gpu > $o/${tag}_gpu_tests_validated.log 2>&1; tail -2 $o/${tag}_gpu_tests_validated.log
PB200_UNVALIDATED=1 timeout 600 python -m pytest tests/test_single_layer_gpu.py -q -m gpu > $o/${tag}_gpu_tests_single_layer.log 2>&1; tail -15 $o/${tag}_gpu_tests_single_layer.log
PB200_UNVALIDATED=1 timeout 600 python -m pytest tests/test_chunk_major_gpu.py -q -m gpu > $o/${tag}_gpu_tests_chunk_major.log 2>&1; tail -15 $o/${tag}_gpu_tests_chunk_major.log
python bench.py --workload synthetic-3m --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_synthetic3m_default.json 2> $o/${tag}_bench_synthetic3m_default.err
PB200_XL_KERNEL_MODE=5 timeout 600 python bench.py --workload synthetic-3m --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_synthetic3m_mode5.json 2> $o/${tag}_bench_synthetic3m_mode5.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$o/${tag}_bench_*.json")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"))
    print("   ", [(k["kernel"][3:], round(k["ms"],3)) for k in d["roofline"]["kernels"]])
PY
if grep -q " passed" $o/${tag}_gpu_tests_chunk_major.log && ! grep -q "failed\|error" $o/${tag}_gpu_tests_chunk_major.log; then
  PB200_XL_KERNEL_MODE=5 ncu --set full --clock-control none --import-source on -k regex:xl_cm_scores -s 3 -c 1 -o $o/${tag}_ncu_cm_synthetic3m python bench.py --workload synthetic-3m --steps 1 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
fi
