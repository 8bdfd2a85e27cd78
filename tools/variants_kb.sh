#!/bin/bash
python bench.py --workload cfg5 --steps 30 --warmup 5 --no-cpu --no-variant --no-e2e --no-workloads 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('cfg5 ms', d['ms_per_step'], 'frac', d['roofline']['step_frac'], [(k['kernel'],k['level'],round(k['ms']*1e3)) for k in d['kernels'] if k['level'] in ('P3','P4')])"
MGA_FORCE_SPLIT=1 python -m pytest tests/test_gpu_cbam.py tests/test_gpu_reference.py -x -q -m gpu 2>&1 | tail -1
python -m pytest tests/test_gpu_next.py tests/test_gpu_cbam.py -x -q -m gpu -k "concat or gates" 2>&1 | tail -1
python tools/concat_bwd_prof.py 128 2>&1 | grep -E "==|bwd_dx|sam_reduce|cam_pool"
