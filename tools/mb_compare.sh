#!/bin/bash
cp mcts_b200/_build/libmcts_b200.so /tmp/lib_orig.so
for v in orig mb6 mb8; do
  if [ $v = orig ]; then cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so; else cp mcts_b200/_build/libmcts_b200_$v.so mcts_b200/_build/libmcts_b200.so; fi
  for w in breakthrough8x8_eps_mast breakthrough8x8_uniform; do echo -n "$v "; python tools/profile_one.py $w 4096 4096 5 | tail -1; done
done
cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so
