#!/bin/bash
# Round-2 GPU session script: GPU tests, E / S benches (default vs PB200_XL_KERNEL_MODE=6 = query-major kernels only),
# optional ncu captures.  Usage (through gpurun, from the repo root): bash profiles/r02_run.sh <tag> [tests] [e] [s] [ncu_e] [ncu_s]
tag=${1:-r02_x}; shift
o=gpurun_out
mkdir -p $o
summ() { python - "$@" <<'PY'
import json,sys
for f in sys.argv[1:]:
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("/")[-1], "value", round(d["value"]), "ms", round(d.get("ms_per_step",0),4), "e2e", round((d.get("e2e") or {}).get("value",0)), "parity", (d.get("parity") or {}).get("ids_bit_equal"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    r=d.get("roofline") or {}
    print("   ", [(k["kernel"][3:].replace("_kernel",""), round(k["ms"],3)) for k in r.get("kernels",[])])
PY
}
for what in "$@"; do
case $what in
tests) python -m pytest tests -x -q -m gpu > $o/${tag}_gpu_tests.log 2>&1; tail -4 $o/${tag}_gpu_tests.log;;
e) python bench.py --steps 20 --warmup 5 --no-secondary > $o/${tag}_bench_eurlex4k.json 2> $o/${tag}_bench_eurlex4k.err || tail -5 $o/${tag}_bench_eurlex4k.err
   PB200_XL_KERNEL_MODE=6 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > $o/${tag}_bench_eurlex4k_mode6.json 2> $o/${tag}_bench_eurlex4k_mode6.err
   summ $o/${tag}_bench_eurlex4k.json $o/${tag}_bench_eurlex4k_mode6.json;;
s) python bench.py --workload synthetic-3m --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_synthetic3m.json 2> $o/${tag}_bench_synthetic3m.err || tail -5 $o/${tag}_bench_synthetic3m.err
   summ $o/${tag}_bench_synthetic3m.json;;
s7) PB200_XL_KERNEL_MODE=6 python bench.py --workload synthetic-3m --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_synthetic3m_mode6.json 2> $o/${tag}_bench_synthetic3m_mode6.err
   summ $o/${tag}_bench_synthetic3m_mode6.json;;
h100k) python bench.py --workload hnsw-100k --steps 5 --warmup 3 > $o/${tag}_bench_hnsw100k.json 2> $o/${tag}_bench_hnsw100k.err || tail -5 $o/${tag}_bench_hnsw100k.err
   summ $o/${tag}_bench_hnsw100k.json;;
h1m) python bench.py --workload hnsw-1m --steps 3 --warmup 3 > $o/${tag}_bench_hnsw1m.json 2> $o/${tag}_bench_hnsw1m.err || tail -5 $o/${tag}_bench_hnsw1m.err
   summ $o/${tag}_bench_hnsw1m.json;;
full) python __graft_entry__.py --smoke 2>&1 | tail -1; python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_default.json 2> $o/${tag}_bench_default.err || tail -5 $o/${tag}_bench_default.err
   summ $o/${tag}_bench_default.json
   python - <<PY
import json
d=json.loads(open("$o/${tag}_bench_default.json").read().strip().splitlines()[-1])
s=d.get("secondary",{})
print("secondary:", {k:(round(v.get("value",0)) if isinstance(v,dict) else v) for k,v in s.items()})
PY
   ;;
pipe) for p in 0 2 3 8; do PB200_XL_PIPELINE=$p python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > $o/${tag}_bench_eurlex4k_pipe$p.json 2> /dev/null; summ $o/${tag}_bench_eurlex4k_pipe$p.json | head -1; done
   PB200_XL_KERNEL_MODE=6 PB200_XL_PIPELINE=8 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > $o/${tag}_bench_eurlex4k_mode6_pipe8.json 2> /dev/null; summ $o/${tag}_bench_eurlex4k_mode6_pipe8.json | head -1;;
ncu_sleaf) ncu --set full --clock-control none --import-source on -k regex:xl_chunk_scores_kernel -s 11 -c 1 -o $o/${tag}_ncu_chunk_synthetic3m_leaf python bench.py --workload synthetic-3m --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> $o/${tag}_ncu_sleaf.err;;
ncu_h) ncu --set full --clock-control none -k regex:hnsw_search -s 2 -c 1 -o $o/${tag}_ncu_hnsw1m python bench.py --workload hnsw-1m --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $o/${tag}_ncu_h.err;;
s8) PB200_XL_KERNEL_MODE=8 python bench.py --workload synthetic-3m --steps 5 --warmup 3 --no-cpu-baseline > $o/${tag}_bench_synthetic3m_mode8.json 2> $o/${tag}_bench_synthetic3m_mode8.err
   summ $o/${tag}_bench_synthetic3m_mode8.json;;
cmtests) python -m pytest tests/test_chunk_major_gpu.py tests/test_xlinear_gpu.py -x -q -m gpu > $o/${tag}_gpu_tests_cm.log 2>&1; tail -4 $o/${tag}_gpu_tests_cm.log;;
ncu_cmg) ncu --set full --clock-control none --import-source on -k regex:xl_cmg_scores -s 5 -c 1 -o $o/${tag}_ncu_cmg_synthetic3m python bench.py --workload synthetic-3m --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> $o/${tag}_ncu_cmg.err;;
ref) python bench.py --impl reference --steps 5 --warmup 2 > $o/${tag}_bench_reference_arm.json 2> $o/${tag}_bench_reference_arm.err; cut -c1-600 $o/${tag}_bench_reference_arm.json;;
ncu_e) ncu --set full --clock-control none --import-source on -k regex:xl_cm_scores_kernel -s ${NCU_SKIP_E:-9} -c 1 -o $o/${tag}_ncu_cm_eurlex4k python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2> $o/${tag}_ncu_e.err;;
ncu_s) ncu --set full --clock-control none --import-source on -k regex:xl_cm_scores_kernel -s ${NCU_SKIP_S:-13} -c 1 -o $o/${tag}_ncu_cm_synthetic3m python bench.py --workload synthetic-3m --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $o/${tag}_ncu_s.err;;
launches) ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_eurlex4k.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-secondary > /dev/null 2>&1;;
sptests) python -m pytest tests/test_hnsw_sparse_gpu.py tests/test_hnsw_gpu.py -x -q -m gpu > $o/${tag}_gpu_tests_hnsw.log 2>&1; tail -6 $o/${tag}_gpu_tests_hnsw.log;;
hsp100k) python bench.py --workload hnsw-sparse-100k --steps 5 --warmup 3 > $o/${tag}_bench_hnsw_sparse100k.json 2> $o/${tag}_bench_hnsw_sparse100k.err || tail -5 $o/${tag}_bench_hnsw_sparse100k.err
   summ $o/${tag}_bench_hnsw_sparse100k.json;;
hrcv1) python bench.py --workload hnsw-rcv1 --steps 5 --warmup 3 > $o/${tag}_bench_hnsw_rcv1.json 2> $o/${tag}_bench_hnsw_rcv1.err || tail -5 $o/${tag}_bench_hnsw_rcv1.err
   summ $o/${tag}_bench_hnsw_rcv1.json;;
ncu_hsp) ncu --set full --clock-control none --import-source on -k regex:hnsw_search -s 2 -c 1 -o $o/${tag}_ncu_hnsw_sparse python bench.py --workload ${NCU_HSP_WORKLOAD:-hnsw-sparse-100k} --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2> $o/${tag}_ncu_hsp.err;;
esac
done
