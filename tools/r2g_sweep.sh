source <(sed -n '1,/^# ---- edit below/p' tools/variants.sh)
export BENCH_ARGS="--no-workloads"
T=libmga_cbam_tuning.so
run default $T
run kbf112 $T MGA_CL_KB_F=112
run kbf448 $T MGA_CL_KB_F=448
run kbb224 $T MGA_CL_KB_B=224
run kbb896 $T MGA_CL_KB_B=896
run kbf112_kbb224 $T MGA_CL_KB_F=112 MGA_CL_KB_B=224
run pf_f $T MGA_CL_PREFETCH_F=1
run pf_b $T MGA_CL_PREFETCH_B=1
