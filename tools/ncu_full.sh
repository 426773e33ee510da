#!/bin/bash
# ncu capture at BENCH size (4096 leaves x 4096 playouts) for the workloads given as arguments.
set -u
mkdir -p gpurun_out
for w in "$@"; do
  python tools/profile_one.py $w 4096 4096 2 > gpurun_out/plainfull_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/proffull_$w python tools/profile_one.py $w 4096 4096 2 > gpurun_out/ncufull_$w.log 2>&1
  tail -1 gpurun_out/plainfull_$w.log
done
