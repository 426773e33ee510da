#!/bin/bash
cp mcts_b200/_build/libmcts_b200.so /tmp/lib_orig.so
for v in orig v858 v867; do
  if [ $v = orig ]; then cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so; else cp mcts_b200/_build/libmcts_b200_$v.so mcts_b200/_build/libmcts_b200.so; fi
  for w in connect4_uniform_midgame knightthrough8x8_eps_mast hex11_uniform_amaf; do echo -n "$v "; python tools/profile_one.py $w 4096 4096 5 | tail -1; done
done
cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so
