# usage: bash tools/cmp_variants.sh "<lib names>" "<level indices>"
for v in $1; do
  export MGA_LIBNAME=$v
  for lv in $2; do
    python tools/run_level.py cfg2 $lv both 2 --split > gpurun_out/rl.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/l_tmp.csv python tools/run_level.py cfg2 $lv both 2 --split > gpurun_out/ncu2.log 2>&1
    echo "== $v level $lv"; python tools/ncu_durations.py gpurun_out/l_tmp.csv 12 | grep -E "pool|reduce|bwd_dx|rescale|total"
  done
  timeout 200 python bench.py --steps 100 --warmup 10 --no-cpu --no-variant 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['ms_per_step'], d['roofline']['step_frac'])"
done
