#!/bin/bash
# Round 2, call 3: the convolution epilogue with coalesced direct stores + 4 / 7 weight stages against the TMA-store build
# (CHECK lines, per-layer timing, role stalls), BN = 64 tiles for the short-K Cout = 128 layers, full GPU suite, bench
# lines (C2, C4), ncu exports as CSV only (gpurun_out must stay below 64 MiB).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 tools/conv_probe 37 2 0 0 0 > $O/r3_probe_direct_c2.log 2>&1; echo "probe direct rc=$?"
grep -E "CHECK|TOTAL|FAIL|error" $O/r3_probe_direct_c2.log | cut -c1-170
timeout 100 tools/conv_probe_tma 37 2 1 0 0 > $O/r3_probe_tma_c2.log 2>&1; echo "tma: $(grep TOTAL $O/r3_probe_tma_c2.log)"
paste <(grep TIME $O/r3_probe_direct_c2.log | awk '{print $2, $(NF-3), $(NF-1)}') <(grep TIME $O/r3_probe_tma_c2.log | awk '{print $(NF-3), $(NF-1)}')
timeout 100 tools/conv_probe 37 1 1 0 0 > $O/r3_probe_direct_c1.log 2>&1; echo "direct c1: $(grep TOTAL $O/r3_probe_direct_c1.log)"
# layers 1,2,17 = down1.block0, down1.block3, up2.block3 (Cout 128, K <= 18 k-blocks) with BN = 64 tiles; then also 16 (up2.block0)
timeout 100 tools/conv_probe 37 2 0 0 0 3 0x20006 > $O/r3_probe_bn64.log 2>&1; grep -E "BN = 64|TOTAL|FAIL" $O/r3_probe_bn64.log | cut -c1-200
timeout 100 tools/conv_probe 37 2 1 0 0 3 0x30006 > $O/r3_probe_bn64b.log 2>&1; grep -E "BN = 64|TOTAL" $O/r3_probe_bn64b.log | cut -c1-200
timeout 120 tools/conv_probe_prof 37 2 1 0 0 > $O/r3_prof_direct_c2.log 2>&1; grep PROF $O/r3_prof_direct_c2.log | cut -c1-250
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > $O/pytest_gpu_r2c3.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|C2 |C3 |C4 " $O/pytest_gpu_r2c3.log | cut -c1-220 | tail -12
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r3_bench_C2.json 2> $O/r3_bench_C2.err; echo "bench C2 rc=$?"
timeout 400 python bench.py --config C4 --steps 3 --warmup 2 > $O/r3_bench_C4.json 2> $O/r3_bench_C4.err; echo "bench C4 rc=$?"
timeout 400 python bench.py --config C3 --steps 3 --warmup 2 > $O/r3_bench_C3.json 2> $O/r3_bench_C3.err
timeout 500 python bench.py --config C5 --steps 2 --warmup 2 > $O/r3_bench_C5.json 2> $O/r3_bench_C5.err
LM_BN64_MASK=0x20006 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r3_bench_C2_bn64.json 2> $O/r3_bench_C2_bn64.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r3_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
# launch list of one bench step (plain launches so that every kernel shows up)
LM_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r3_launches.csv \
    python bench.py --steps 1 --warmup 0 > $O/r3_bench_under_ncu.json 2> $O/r3_bench_under_ncu.err; echo "ncu launches rc=$?"
# one full 37-slice wave of the convolution kernel, all metrics -> raw CSV (the report itself is too large to bring back)
timeout 600 ncu --set full --clock-control none -k regex:conv_tc_kernel -c 21 -f -o /tmp/r3_conv_wave tools/conv_probe 37 2 1 0 0 0 > $O/r3_ncu_conv_wave.log 2>&1; echo "ncu conv wave rc=$?"
ncu -i /tmp/r3_conv_wave.ncu-rep --page raw --csv > $O/r3_conv_wave_raw.csv 2>/dev/null; ls -la $O/r3_conv_wave_raw.csv
# the small kernels (stem, upsample, bodymask, resize, ccl, ...) -> raw CSV
LM_GRAPHS=0 timeout 600 ncu --set full --clock-control none -k regex:'^(?!.*conv_tc).*' -s 60 -c 130 -f -o /tmp/r3_small python tools/profile_small_kernels.py > $O/r3_small.log 2> $O/r3_ncu_small.err; echo "ncu small rc=$?"
ncu -i /tmp/r3_small.ncu-rep --page raw --csv > $O/r3_small_raw.csv 2>/dev/null; ls -la $O/r3_small_raw.csv
# source-level view of the two streaming kernels that sit far below the HBM roofline
LM_GRAPHS=0 timeout 300 ncu --set full --clock-control none --import-source on -k regex:'stem_kernel_v2|upsample2x_cells' -s 4 -c 2 -f -o /tmp/r3_stream python tools/profile_small_kernels.py > /dev/null 2>&1
ncu -i /tmp/r3_stream.ncu-rep --page source --csv > $O/r3_stream_source.csv 2>/dev/null; ncu -i /tmp/r3_stream.ncu-rep --page raw --csv > $O/r3_stream_raw.csv 2>/dev/null
du -sh $O; ls -la $O | tail -40
