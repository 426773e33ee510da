#!/bin/bash
for v in false true; do
  cp mcts_b200/_build/libmcts_b200_$v.so mcts_b200/_build/libmcts_b200.so
  for w in connect4_uniform_midgame breakthrough8x8_uniform breakthrough8x8_eps_mast knightthrough8x8_eps_mast; do
    echo -n "rolled=$v "; python tools/profile_one.py $w 4096 4096 5 | tail -1
  done
done
