#!/bin/bash
# ncu capture at BENCH size (4096 leaves x 4096 playouts) for the workloads given as arguments.  The counter summary is
# written next to the report (gpurun_out/ncusum_<w>.txt); with NCU_DROP_REP=1 the ~10 MB report itself is deleted on the
# box (gpurun copies back at most 64 MiB).
set -u
mkdir -p gpurun_out
for w in "$@"; do
  python tools/profile_one.py $w 4096 4096 2 > gpurun_out/plainfull_$w.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:rollout_kernel -s 2 -c 1 -f -o gpurun_out/proffull_$w python tools/profile_one.py $w 4096 4096 2 > gpurun_out/ncufull_$w.log 2>&1
  python tools/ncu_summary.py gpurun_out/proffull_$w.ncu-rep > gpurun_out/ncusum_$w.txt 2>&1
  if [ "${NCU_DROP_REP:-0}" = "1" ]; then rm -f gpurun_out/proffull_$w.ncu-rep; fi
  tail -1 gpurun_out/plainfull_$w.log
done
