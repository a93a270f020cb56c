#!/bin/bash
# Round 2, call 8: epilogue v2 (both operand planes packed once with cvt.rn.f16x2, out_scale folded into the shared-memory
# BN constants) + looped taps on the one-channel-block layers; CHECKs, timing against always-unrolled / always-looped,
# role profile, FULL GPU suite, bench lines of all configs, ncu launch list and one full conv wave (CSV exports only).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 tools/conv_probe 37 2 0 0 0 > $O/r8_single_check.log 2>&1; echo "single check rc=$? ok=$(grep -c ' ok' $O/r8_single_check.log) $(grep TOTAL $O/r8_single_check.log)"; grep -E "FAIL|bad=[1-9]|timeout" $O/r8_single_check.log | head -5
timeout 200 tools/conv_probe 37 2 0 0 1 > $O/r8_pair_check.log 2>&1; echo "pair check rc=$? ok=$(grep -c ' ok' $O/r8_pair_check.log) $(grep TOTAL $O/r8_pair_check.log)"; grep -E "FAIL|bad=[1-9]|timeout" $O/r8_pair_check.log | head -5
for v in conv_probe conv_probe_tl0 conv_probe_tl2; do timeout 100 tools/$v 37 2 1 0 0 > $O/r8_$v.log 2>&1; echo "$v: $(grep TOTAL $O/r8_$v.log)"; done
paste <(grep TIME $O/r8_conv_probe.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r8_conv_probe_tl0.log | awk '{print $(NF-3)}') <(grep TIME $O/r8_conv_probe_tl2.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_prof 37 2 1 0 0 > $O/r8_prof.log 2>&1; grep "PROF " $O/r8_prof.log | cut -c1-260
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > $O/pytest_gpu_r2c8.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error" $O/pytest_gpu_r2c8.log | cut -c1-220 | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r8_bench_C2.json 2> $O/r8_bench_C2.err; echo "bench C2 rc=$?"
timeout 400 python bench.py --config C3 --steps 3 --warmup 3 > $O/r8_bench_C3.json 2> $O/r8_bench_C3.err
timeout 400 python bench.py --config C4 --steps 3 --warmup 3 > $O/r8_bench_C4.json 2> $O/r8_bench_C4.err
timeout 500 python bench.py --config C5 --steps 2 --warmup 3 > $O/r8_bench_C5.json 2> $O/r8_bench_C5.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r8_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
LM_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r8_launches.csv \
    python bench.py --steps 1 --warmup 0 > $O/r8_bench_under_ncu.json 2> $O/r8_bench_under_ncu.err; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:conv_tc_kernel -c 21 -f -o /tmp/r8_conv_wave tools/conv_probe 37 2 1 0 0 0 > $O/r8_ncu_conv_wave.log 2>&1; echo "ncu conv wave rc=$?"
ncu -i /tmp/r8_conv_wave.ncu-rep --page raw --csv > $O/r8_conv_wave_raw.csv 2>/dev/null; ls -la $O/r8_conv_wave_raw.csv
du -sh $O
