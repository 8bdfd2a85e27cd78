#!/bin/bash
# sweep of tools/ringbw.bin: B C H W CS NT ST CH passes write idle_ns
R=./tools/ringbw.bin
# P3 fp32 geometry (cluster 8 x 10 rows): pure HBM read, then fwd-like 3 passes + write, with and without idle phases
for ST in 2 3 4 6; do $R 64 64 80 80 8 512 $ST 4 1 0 0; done
$R 64 64 80 80 8 512 4 2 1 0 0
$R 64 64 80 80 8 512 8 2 1 0 0
$R 64 64 80 80 8 512 4 4 1 1 0
$R 64 64 80 80 8 512 4 4 3 1 0
$R 64 64 80 80 8 512 3 4 3 1 0
$R 64 64 80 80 8 512 6 2 3 1 0
$R 64 64 80 80 8 512 4 4 3 1 3000
$R 64 64 80 80 8 512 4 4 3 1 6000
$R 64 64 80 80 8 256 4 4 3 1 3000
$R 256 64 80 80 8 512 4 4 3 1 3000
# bwd-like traffic: 5 reads + 1 write of the slice (x3, g2 -> modelled as 5 passes over x)
$R 64 64 80 80 8 512 4 4 5 1 0
$R 64 64 80 80 8 512 4 4 5 1 3000
# cluster-16-like slices (5 rows)
$R 64 64 80 80 16 512 4 4 3 1 3000
$R 64 64 80 80 16 256 4 4 3 1 3000
# P4 / P5 geometry: short rows
$R 64 128 40 40 4 512 4 8 3 1 0
$R 64 128 40 40 4 512 4 8 3 1 3000
$R 64 256 20 20 2 512 4 16 3 1 0
$R 64 256 20 20 2 512 4 16 3 1 3000
