source <(sed -n '1,/^# ---- edit below/p' tools/variants.sh)
export BENCH_ARGS="--no-workloads"
run split_cfg2 libmga_cbam_tuning.so MGA_CL=0
BENCH_ARGS="--no-workloads --one-stream" run split_cfg2_one libmga_cbam_tuning.so MGA_CL=0
