#!/bin/bash
# Round 2, call 2: full GPU suite with the fixes, role-stall profiles (single vs pair kernel), ncu captures (full conv wave
# for the per-layer traffic table, the small kernels, two pair launches), launch list of one bench step, bench lines
# for C2 (graphs on / off), C3, C4, C5.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -q -s --durations=12 > $O/pytest_gpu_r2c2.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error|C2 |C3 |C4 |float|rescaled|scaled" $O/pytest_gpu_r2c2.log | cut -c1-220 | tail -30
timeout 120 tools/conv_probe_prof 37 1 1 0 0 > $O/prof_single_c1.log 2>&1; echo "prof single rc=$?"
timeout 120 tools/conv_probe_prof 37 1 1 0 1 > $O/prof_pair_c1.log 2>&1; echo "prof pair rc=$?"
timeout 120 tools/conv_probe_prof 37 2 1 0 0 > $O/prof_single_c2.log 2>&1
grep PROF $O/prof_single_c1.log | cut -c1-260; grep PROF $O/prof_pair_c1.log | cut -c1-260
timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_C2.json 2> $O/bench_C2.err; echo "bench C2 rc=$?"
LM_GRAPHS=0 timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_C2_nographs.json 2> $O/bench_C2_nographs.err
timeout 400 python bench.py --config C3 --steps 3 --warmup 2 > $O/bench_C3.json 2> $O/bench_C3.err; echo "bench C3 rc=$?"
timeout 400 python bench.py --config C4 --steps 3 --warmup 2 > $O/bench_C4.json 2> $O/bench_C4.err; echo "bench C4 rc=$?"
timeout 500 python bench.py --config C5 --steps 2 --warmup 2 > $O/bench_C5.json 2> $O/bench_C5.err; echo "bench C5 rc=$?"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_ref_C2.json 2> $O/bench_ref_C2.err; echo "bench ref rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_C*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
# launch list of one bench step (plain launches so that every kernel shows up)
LM_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/launches_r2.csv \
    python bench.py --steps 1 --warmup 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err; echo "ncu launches rc=$?"
# one full 37-slice wave of the convolution kernel, all metrics (per-layer DRAM traffic / tensor pipe / smem)
timeout 600 ncu --set full --clock-control none -k regex:conv_tc_kernel -c 21 -f -o $O/r2_conv_wave tools/conv_probe 37 1 1 0 0 0 > $O/ncu_conv_wave.log 2>&1; echo "ncu conv wave rc=$?"
# two launches of the pair kernel with source correlation (why is it slower?)
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair_kernel --launch-skip 10 --launch-count 1 -f -o $O/r2_pair_bn128 tools/conv_probe 37 1 1 0 1 0 > $O/ncu_pair128.log 2>&1; echo "ncu pair128 rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_pair_kernel --launch-skip 19 --launch-count 1 -f -o $O/r2_pair_bn64 tools/conv_probe 37 1 1 0 1 0 > $O/ncu_pair64.log 2>&1; echo "ncu pair64 rc=$?"
# the small kernels (stem v1/v2, upsample v1/v2, bodymask, ccl ...)
LM_GRAPHS=0 timeout 600 ncu --set full --clock-control none -k regex:'^(?!.*conv_tc).*' -s 60 -c 120 -f -o $O/r2_small python tools/profile_small_kernels.py > $O/small_r2.log 2> $O/ncu_small_r2.err; echo "ncu small rc=$?"
LM_GRAPHS=0 LM_STEM_V2=0 LM_UPSAMPLE_V2=0 timeout 600 ncu --set full --clock-control none -k regex:'stem_kernel|upsample2x' -s 10 -c 10 -f -o $O/r2_small_v1 python tools/profile_small_kernels.py > $O/small_r2_v1.log 2> $O/ncu_small_r2_v1.err; echo "ncu small v1 rc=$?"
ls -la $O | tail -30
