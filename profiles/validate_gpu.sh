#!/bin/bash
# End-of-round validation on one B200 (run through gpurun from the repo root): GPU parity tests, the bench lines, the ncu
# launch list and one full ncu capture of the dominant kernel.  Outputs go to gpurun_out/<tag>_*.
tag=${1:-r01_u}
o=gpurun_out
python -m pytest tests -x -q -m gpu > $o/${tag}_gpu_tests_all.log 2>&1; tail -3 $o/${tag}_gpu_tests_all.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 20 --warmup 3 > $o/${tag}_bench_eurlex4k.json 2> $o/${tag}_bench_eurlex4k.err
python bench.py --impl reference --steps 5 --warmup 1 > $o/${tag}_bench_reference_arm.json 2> /dev/null
python bench.py --workload synthetic-3m --steps 5 --warmup 3 > $o/${tag}_bench_synthetic3m.json 2> $o/${tag}_bench_synthetic3m.err
python bench.py --workload hnsw-100k --steps 5 --warmup 3 > $o/${tag}_bench_hnsw100k.json 2> $o/${tag}_bench_hnsw100k.err
for n in 2 3; do PB200_XL_PIPELINE=$n python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null > $o/${tag}_bench_eurlex4k_subtiles$n.json; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches_eurlex4k.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:xl_chunk_scores -s 8 -c 1 -o $o/${tag}_ncu_cs_eurlex4k python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$o/${tag}_bench_*.json")):
    try: d=json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e: print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(d["value"]), d.get("ms_per_step"), (d.get("e2e") or {}).get("value"), (d.get("roofline") or {}).get("frac"), (d.get("cpu_baseline") or {}).get("value"))
    if "roofline" in d and "kernels" in d["roofline"]: print("   ", [(k["kernel"][3:], round(k["ms"],3)) for k in d["roofline"]["kernels"]])
PY
