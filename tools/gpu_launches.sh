#!/bin/bash
# Last GPU call of the round: one clean bench line, then the ncu launch list of one bench step (only this repo's
# kernels are profiled: the synthetic-weight training that bench.py runs first on a fresh box uses torch kernels).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_final.json 2> $O/bench_final.err; echo "bench rc=$?"; cat $O/bench_final.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches.csv \
    -k regex:'^(conv_tc_kernel|stem_kernel|upsample2x_kernel|bodymask_kernel|resize_kernel|reshape_kernel|ccl_.*|roots_.*|root_area_kernel|best_root_kernel|select_root_kernel|region_stats_kernel|merge_loop_kernel|map_labels_kernel|keep_complement_kernel|seed_outside_kernel|clear_outside_kernel|paint_.*|scan_blocks_kernel|bbox_init_kernel|binarize_kernel|max_u8_kernel|fuse_kernel|prep_conv_weights_kernel)' \
    python bench.py --steps 1 --warmup 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err; echo "ncu launches rc=$?"
grep -c conv_tc $O/launches.csv; wc -l $O/launches.csv
