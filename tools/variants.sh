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
# ---- edit below: one `run <name> <library> [ENV=value ...]` line per variant (BENCH_ARGS adds bench.py flags)
L=libmga_cbam.so
run default $L
run per_phase $L MGA_CL=0
BENCH_ARGS=--one-stream run default_one_stream $L
BENCH_ARGS=--one-stream run per_phase_one_stream $L MGA_CL=0
