#!/bin/bash
# Run on the GPU box (gpurun): warm caches, then (1) per-launch durations of one bench step, (2) one full ncu
# capture of the dominant kernel.  Outputs under gpurun_out/.
set -x
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python bench.py --steps 1 --warmup 1 > gpurun_out/bench_warm.json 2> gpurun_out/bench_warm.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 1 --warmup 0 > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 24 -c 4 -f -o gpurun_out/prof_conv \
    python bench.py --steps 1 --warmup 0 > /dev/null 2> gpurun_out/ncu_full.err
ls -la gpurun_out
