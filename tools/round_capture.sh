#!/bin/bash
# end-of-iteration evidence: bench (both arms), launch list, full ncu captures of the two dominant kernels
tag=${1:-r1b}
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc=$?"
python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_cfg3_$tag.json 2> gpurun_out/bench_cfg3_$tag.err; echo "cfg3 rc=$?"
python bench.py --workload cfg5 --steps 50 --warmup 5 --no-cpu > gpurun_out/bench_cfg5_$tag.json 2> gpurun_out/bench_cfg5_$tag.err; echo "cfg5 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-graph --one-stream > gpurun_out/ncu_l.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_fwd -s 2 -c 1 -f -o gpurun_out/cl_fwd_p3_$tag python tools/run_level.py cfg2 0 fwd > gpurun_out/ncu_clf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_bwd -s 2 -c 1 -f -o gpurun_out/cl_bwd_p3_$tag python tools/run_level.py cfg2 0 both > gpurun_out/ncu_clb.log 2>&1
python - <<P
import json
for n in ("bench_$tag","bench_cfg3_$tag","bench_cfg5_$tag","bench_ref_$tag"):
    try:
        d=json.load(open(f"gpurun_out/{n}.json"))
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("step_frac"), d.get("variants"), (d.get("e2e") or {}).get("value"))
    except Exception as e:
        print(n, "ERR", e)
P
