#!/bin/bash
for w in connect4_uniform_midgame breakthrough8x8_uniform; do
for cfg in "64 4096 2" "256 4096 2" "256 256 2" "512 512 2" "128 4096 4" "256 4096 4" "256 1024 8" "512 4096 2" "1024 1024 2"; do
  set -- $cfg
  echo -n "$w min=$1 max=$2 div=$3: "
  MR_GRAB_MIN=$1 MR_GRAB_MAX=$2 MR_GRAB_DIV=$3 python tools/profile_one.py $w 4096 4096 6 | tail -1
done; done
