#!/bin/bash
# quick iteration check: parity of the block tests, P3 timelines, short bench
timeout 900 python -m pytest tests/test_gpu_cbam.py tests/test_gpu_cluster_geometry.py tests/test_gpu_reference.py -x -q -m gpu 2>&1 | tail -3
for w in fwd bwd; do MGA_LIBNAME=libmga_cbam_tuning.so python tools/timeline.py cfg2 $w 2>&1 | grep -A1 "level P" | grep -v "^--"; done
python bench.py --steps 100 --warmup 5 --no-cpu --no-variant --no-e2e --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('ms', d['ms_per_step'], 'frac', d['roofline']['step_frac'], [(k['kernel'][:6],k['level'],round(k['ms']*1e3)) for k in d['kernels'] if k['kernel']!='bwd_wgrad'])"
