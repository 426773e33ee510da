#!/bin/bash
# Round evidence: bench line, reference line, ncu launch list of the bench command, full ncu capture of every bench
# workload's kernel at bench size.  Run under gpurun on one GPU.
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err
tail -c 400 gpurun_out/bench.json
python tools/measure_int_peaks.py > gpurun_out/int_peaks.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
bash tools/ncu_full.sh breakthrough8x8_eps_mast connect4_uniform_midgame
NCU_DROP_REP=1 bash tools/ncu_full.sh breakthrough8x8_uniform hex11_uniform_amaf othello_decisive_win knightthrough8x8_eps_mast hex11_eps_mast
