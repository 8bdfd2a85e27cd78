#!/bin/bash
# which (shape, launch path) hangs?  every case in its own process with a short timeout
for shape in "2 16 4 4" "1 8 8 8" "2 18 8 12"; do
  for path in cluster split flow fused; do
    envs=""
    case $path in split) envs="MGA_FORCE_SPLIT=1";; flow) envs="MGA_USE_FLOW=1";; fused) envs="MGA_USE_FUSED=1";; esac
    for mode in "multiply add" "add multiply"; do
      out=$(env $envs PYTHONPATH=. timeout 25 python -m tests._cluster_case $shape float32 $mode 2>&1 | tail -1)
      echo "shape=[$shape] path=$path mode=[$mode] rc=$? -> ${out:0:80}"
    done
  done
done
