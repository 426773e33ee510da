#!/bin/bash
# Full ncu capture of the rollout kernel for each workload (run under gpurun; one GPU).
set -u
mkdir -p gpurun_out
for w in breakthrough8x8_eps_mast connect4_uniform_midgame breakthrough8x8_uniform hex11_uniform_amaf othello_decisive_win knightthrough8x8_eps_mast; do
  python tools/profile_one.py $w 4096 512 3 > gpurun_out/plain_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 3 -c 1 -f -o gpurun_out/prof_$w python tools/profile_one.py $w 4096 512 3 > gpurun_out/ncu_$w.log 2>&1
  tail -1 gpurun_out/plain_$w.log
done
