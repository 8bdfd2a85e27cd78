#!/bin/bash
# one GPU iteration: parity subset, bench, per-phase timeline, ncu full captures of the two cluster kernels (P3)
tag=${1:-it}
timeout 600 python -m pytest tests/test_gpu_cbam.py -x -q -m gpu -k "cluster or full_size" > gpurun_out/t_$tag.log 2>&1; echo "pytest rc=$?" >> gpurun_out/t_$tag.log
tail -5 gpurun_out/t_$tag.log
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-variant > gpurun_out/b_$tag.json 2> gpurun_out/b_$tag.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open("gpurun_out/b_$tag.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "frac", d["value"]/6453.4)
print([(k["kernel"],k["level"],round(k["ms"]*1e3,1)) for k in d["kernels"]])
P
python tools/timeline.py cfg2 fwd > gpurun_out/tl_fwd_$tag.log 2>&1; python tools/timeline.py cfg2 bwd > gpurun_out/tl_bwd_$tag.log 2>&1
cat gpurun_out/tl_fwd_$tag.log gpurun_out/tl_bwd_$tag.log
if [ "$2" != "noncu" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_fwd -s 2 -c 1 -f -o gpurun_out/cl_fwd_p3_$tag python tools/run_level.py cfg2 0 fwd > gpurun_out/ncu_clf.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cl_bwd -s 2 -c 1 -f -o gpurun_out/cl_bwd_p3_$tag python tools/run_level.py cfg2 0 both > gpurun_out/ncu_clb.log 2>&1
fi
exit 0
