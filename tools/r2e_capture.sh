#!/bin/bash
# end-of-round evidence (round 2, final): full GPU suite, bench (both arms), launch list, ncu --set full of the tcgen05 concat backward kernel
tag=${1:-r2e}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/tests_$tag.log; cat gpurun_out/tests_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > gpurun_out/ncu_l.log 2>&1
MGA_CONCAT_BWD=tc python tools/concat_bwd_check.py > /dev/null 2>&1 && \
MGA_CONCAT_BWD=tc timeout 300 ncu --set full --clock-control none --import-source on -k regex:concat_fwd_kernel -s 0 -c 1 -f -o gpurun_out/concat_bwd_$tag python tools/concat_bwd_check.py 512 > gpurun_out/ncu_cb.log 2>&1
python - <<P
import json
for n in ("bench_$tag","bench_ref_$tag"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("step_frac"), (d.get("e2e") or {}).get("value"))
        for k,v in (d.get("workloads") or {}).items(): print("  ", k, v.get("ms_per_step"), v.get("step_frac"))
    except Exception as e:
        print(n, "ERR", e)
P
