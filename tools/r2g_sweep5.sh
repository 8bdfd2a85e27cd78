source <(sed -n '1,/^# ---- edit below/p' tools/variants.sh)
export BENCH_ARGS="--no-workloads --workload cfg5 --steps 20"
T=libmga_cbam_tuning.so
run c5_default $T MGA_CL_DEBUG=1
run c5_kbf640 $T MGA_CL_KB_F=320 MGA_CL_DEBUG=1
run c5_kbb1280 $T MGA_CL_KB_B=640 MGA_CL_DEBUG=1
run c5_both $T MGA_CL_KB_F=320 MGA_CL_KB_B=640 MGA_CL_DEBUG=1
