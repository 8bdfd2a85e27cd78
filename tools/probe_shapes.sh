#!/bin/bash
# every (shape, dtype, launch path) in its own process with a short timeout -- finds a wedged kernel without hanging the run
# usage: bash tools/probe_shapes.sh   (edit CASES)
CASES=("2 18 8 16 bfloat16" "2 50 12 12 float32" "1 3 8 8 float32" "2 130 20 20 float16" "3 7 16 16 float32" "2 34 24 8 bfloat16")
for c in "${CASES[@]}"; do
  for path in cluster split; do
    envs=""
    case $path in split) envs="MGA_FORCE_SPLIT=1";; esac
    for mode in "multiply add" "add multiply"; do
      out=$(env $envs PYTHONPATH=. timeout 25 python -m tests._cluster_case $c $mode 2>&1 | tail -1)
      echo "case=[$c] path=$path mode=[$mode] -> ${out:0:100}"
    done
  done
done
