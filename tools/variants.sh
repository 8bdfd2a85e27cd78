#!/bin/bash
# bench the step under several library builds / env settings: "name|LIB|ENV..." lines on stdin or args
run() {
  name=$1; lib=$2; shift 2
  out=$(env MGA_LIBNAME=$lib "$@" timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu --no-variant --no-e2e $BENCH_ARGS 2>gpurun_out/var_$name.err)
  echo "$out" > gpurun_out/var_$name.json
  python - "$name" <<P
import json,sys
try:
    d=json.load(open("gpurun_out/var_"+sys.argv[1]+".json"))
    print(f"{sys.argv[1]:28s} ms {d['ms_per_step']:.4f}  GB/s {d['value']:.0f}  frac {d['value']/6453.4:.3f} ", [(k['kernel'][:6],k['level'],round(k['ms']*1e3)) for k in d['kernels'] if k['kernel']!='bwd_wgrad'])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
P
}
L=libmga_cbam.so
run base $L
run any_b330 $L MGA_CL_ANYCS=1 MGA_CL_KB_B=330
run any_b380 $L MGA_CL_ANYCS=1 MGA_CL_KB_B=380
run any_b280 $L MGA_CL_ANYCS=1 MGA_CL_KB_B=280
run any_f190 $L MGA_CL_ANYCS=1 MGA_CL_KB_F=190
run any_f170 $L MGA_CL_ANYCS=1 MGA_CL_KB_F=170
run any_f300 $L MGA_CL_ANYCS=1 MGA_CL_KB_F=300
MGA_CL_DEBUG=1 MGA_CL_ANYCS=1 MGA_CL_KB_B=330 MGA_CL_KB_F=190 python tools/run_level.py cfg2 0 both 1 2>&1 | grep mga | sort -u
