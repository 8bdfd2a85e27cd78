#!/bin/bash
# fresh per-line instruction attribution of the cluster kernels at P3 (BASELINE configs[1]) + a short bench
tag=${1:-r2d}
python bench.py --no-cpu --no-variant --no-e2e --no-workloads > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python tools/run_level.py cfg2 0 both > /dev/null 2>&1 && {
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_fwd -s 2 -c 1 -f -o gpurun_out/cl_fwd_p3_$tag python tools/run_level.py cfg2 0 both > gpurun_out/ncu_clf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_bwd -s 2 -c 1 -f -o gpurun_out/cl_bwd_p3_$tag python tools/run_level.py cfg2 0 both > gpurun_out/ncu_clb.log 2>&1
}
for k in fwd bwd; do
  python tools/ncu_lines.py gpurun_out/cl_${k}_p3_$tag.ncu-rep 70 --inst > gpurun_out/lines_${k}_$tag.txt 2>&1
  python tools/ncu_sass.py gpurun_out/cl_${k}_p3_$tag.ncu-rep > gpurun_out/sass_${k}_$tag.txt 2>&1
done
python - <<P
import json
d=json.loads(open("gpurun_out/bench_$tag.json").read().strip().splitlines()[-1])
print(d.get("ms_per_step"), d.get("value"), d["roofline"]["step_frac"], [(k["kernel"],k["level"],k["ms"]) for k in d["kernels"]])
P
