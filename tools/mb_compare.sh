#!/bin/bash
# A/B timing of library variants: tools/mb_compare.sh "<variants>" "<workloads>"  (variants = suffixes of _build/libmcts_b200_<v>.so)
cp mcts_b200/_build/libmcts_b200.so /tmp/lib_orig.so
for v in $1; do
  cp mcts_b200/_build/libmcts_b200_$v.so mcts_b200/_build/libmcts_b200.so
  for w in $2; do echo -n "$v "; python tools/profile_one.py $w 4096 4096 5 | tail -1; done
done
cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so
