#!/bin/bash
# A/B of library variants by executed instructions and duration (ncu, one kernel launch each):
# tools/ab_inst.sh "<variants>" "<workloads>"   (variant "cur" = the library as built)
cp mcts_b200/_build/libmcts_b200.so /tmp/lib_orig.so
for v in $1; do
  if [ "$v" != "cur" ]; then cp mcts_b200/_build/libmcts_b200_$v.so mcts_b200/_build/libmcts_b200.so; else cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so; fi
  for w in $2; do
    ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread --clock-control none -k regex:rollout_kernel -s 2 -c 1 --csv python tools/profile_one.py $w 4096 4096 3 2>/dev/null | grep -E "inst_executed|time_duration|issue_active|registers" | awk -F'","' -v v=$v -v w=$w '{gsub(/"/,"",$NF); printf "%s %s %s %s\n", v, w, $(NF-2), $NF}'
  done
done
cp /tmp/lib_orig.so mcts_b200/_build/libmcts_b200.so
