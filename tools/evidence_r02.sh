#!/bin/bash
# Round-2 evidence: smoke, reference line, bench line, ncu launch list of the bench command, full ncu capture of the
# bench workloads' kernels at bench size (summaries only; the headline's report is kept).  Run under gpurun on one GPU.
set -u
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2_smoke.log 2>&1; tail -5 gpurun_out/r2_smoke.log
python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err
tail -c 300 gpurun_out/r2_bench.json; echo
python tools/measure_int_peaks.py gpurun_out/r2_int_peaks.json > /dev/null 2>&1
python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 3 --warmup 3 --headline-only --no-cpu-baseline > gpurun_out/r2_ncu_launches.log 2>&1
bash tools/ncu_full.sh breakthrough8x8_eps_mast
NCU_DROP_REP=1 bash tools/ncu_full.sh hex11_eps_mast knightthrough8x8_eps_mast_opening8 connect4_uniform_midgame breakthrough8x8_uniform hex11_uniform_amaf othello_decisive_win knightthrough8x8_eps_mast
python tools/ncu_lines.py gpurun_out/proffull_breakthrough8x8_eps_mast.ncu-rep > gpurun_out/r2_ncu_lines_headline.txt 2>&1
ls -la gpurun_out | tail -30
