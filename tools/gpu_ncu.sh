#!/bin/bash
# ncu captures of the final engine (gpurun): (1) per-launch durations of one bench step, (2) --set full on four
# launches of the convolution kernel, (3) --set full on the other kernels of one small volume.  The large small-kernel
# report is exported to CSV on the box and deleted (gpurun brings back at most 64 MiB).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err; echo "ncu launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 24 -c 4 -f -o $O/prof_conv \
    python bench.py --steps 1 --warmup 0 > /dev/null 2> $O/ncu_full.err; echo "ncu full rc=$?"
ncu -i $O/prof_conv.ncu-rep --page raw --csv > $O/prof_conv_raw.csv 2>/dev/null
timeout 400 ncu --set full --clock-control none -k regex:'^(?!.*conv_tc).*' -s 40 -c 45 -f -o $O/prof_small \
    python tools/profile_small_kernels.py > $O/small.log 2> $O/ncu_small.err; echo "ncu small rc=$?"
ncu -i $O/prof_small.ncu-rep --page raw --csv > $O/prof_small_raw.csv 2>/dev/null
rm -f $O/prof_small.ncu-rep
du -sh $O; ls -la $O
