#!/bin/bash
# end-of-round evidence (round 2, last session): full GPU suite, bench (both arms, the driver's flags), launch list, ncu --set full of the
# MaskSPADE feature-side kernels at the P3 shape
tag=${1:-r2g}
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 > gpurun_out/tests_$tag.log; cat gpurun_out/tests_$tag.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python bench.py --impl reference --gpus 1 --steps 5 --warmup 3 > gpurun_out/bench_ref_$tag.json 2>/dev/null; echo "ref rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-variant --no-e2e --no-workloads > gpurun_out/ncu_l.log 2>&1
for a in "64 80 80 64 f32" "128 40 40 64 f32" "256 20 20 64 f32" "128 80 80 64 bf16"; do python tools/spade_prof.py $a; done > gpurun_out/spade_prof_$tag.log 2>&1; cat gpurun_out/spade_prof_$tag.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spade_ -s 6 -c 2 -f -o gpurun_out/spade_$tag python tools/spade_prof.py > gpurun_out/ncu_sp.log 2>&1
python - <<P
import json
for n in ("bench_$tag","bench_ref_$tag"):
    try:
        d=json.loads(open(f"gpurun_out/{n}.json").read().strip().splitlines()[-1])
        print(n, d.get("ms_per_step"), d.get("value"), (d.get("roofline") or {}).get("step_frac"), (d.get("e2e") or {}).get("value"))
        for k,v in (d.get("workloads") or {}).items(): print("  ", k, v.get("ms_per_step"), v.get("step_frac"))
        sp = d.get("spade_block")
        if sp: print("   spade", sp.get("ms_per_step"), (sp.get("roofline") or {}).get("frac"), sp.get("kernels"), sp.get("error"))
    except Exception as e:
        print(n, "ERR", e)
P
