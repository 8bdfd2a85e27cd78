#!/bin/bash
# end-of-round evidence (round 2): bench (both arms), launch list of the default bench command, ncu --set full of the dominant kernel of the
# block (cl_bwd at P3) and of the tcgen05 concat kernel.  Every ncu pass runs only after the same command has exited 0 without ncu.
tag=${1:-r2}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > gpurun_out/ncu_l.log 2>&1
python tools/run_level.py cfg2 0 both > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_bwd -s 2 -c 1 -f -o gpurun_out/cl_bwd_p3_$tag python tools/run_level.py cfg2 0 both > gpurun_out/ncu_clb.log 2>&1
python tools/concat_prof.py 32 > /dev/null 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:concat_fwd_res -s 1 -c 1 -f -o gpurun_out/concat_fwd_$tag python tools/concat_prof.py 32 > gpurun_out/ncu_cc.log 2>&1
python - <<P
import json
for n in ("bench_$tag","bench_ref_$tag"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("step_frac"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
P
