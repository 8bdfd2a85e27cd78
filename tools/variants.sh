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
run all_cluster $L
run p5_split $L MGA_CL_MINKB_F=500 MGA_CL_MINKB_B=500
run p45_split $L MGA_CL_MINKB_F=1000 MGA_CL_MINKB_B=1000
run p5fwd_split $L MGA_CL_MINKB_F=500
run p5bwd_split $L MGA_CL_MINKB_B=500
run p3_split $L MGA_CL_MAXKB_F=1000 MGA_CL_MAXKB_B=1000
run p3fwd_split $L MGA_CL_MAXKB_F=1000
run all_split $L MGA_CL=0
