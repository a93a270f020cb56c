#!/bin/bash
# Round 2, call 11 (final tree): smoke, the FULL GPU suite, the bench line of the default configuration (C2), the static-corner
# upsample kernel against the two older ones (bench lines + the bit-identity tests inside the suite), 74-slice waves, the ncu
# launch list of one bench step and an ncu --set full pass over the forward's small kernels (CSV exports only).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 300 python __graft_entry__.py smoke > $O/r11_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r11_smoke.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > $O/r11_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error" $O/r11_pytest_gpu.log | cut -c1-220 | tail -6
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r11_bench_C2.json 2> $O/r11_bench_C2.err; echo "bench C2 rc=$?"
LM_UPSAMPLE_V2=1 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r11_bench_C2_up1.json 2> $O/r11_bench_C2_up1.err; echo "bench C2 upsample_v2=1 rc=$?"
LM_WAVE_SLICES=74 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r11_bench_C2_w74.json 2> $O/r11_bench_C2_w74.err; echo "bench C2 waves of 74 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r11_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
LM_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r11_launches.csv \
    python bench.py --steps 1 --warmup 0 > $O/r11_bench_under_ncu.json 2> $O/r11_bench_under_ncu.err; echo "ncu launches rc=$?"
LM_GRAPHS=0 timeout 400 ncu --set full --clock-control none -k regex:'stem_kernel|upsample2x|bodymask|resize_kernel' -c 12 -f -o /tmp/r11_small \
    python bench.py --steps 1 --warmup 0 > $O/r11_ncu_small.log 2>&1; echo "ncu small rc=$?"
ncu -i /tmp/r11_small.ncu-rep --page raw --csv > $O/r11_small_raw.csv 2>/dev/null; ls -la $O/r11_small_raw.csv
timeout 300 python bench.py --config C3 --steps 3 --warmup 3 > $O/r11_bench_C3.json 2> $O/r11_bench_C3.err; echo "bench C3 rc=$?"
timeout 300 python bench.py --config C4 --steps 3 --warmup 3 > $O/r11_bench_C4.json 2> $O/r11_bench_C4.err; echo "bench C4 rc=$?"
timeout 300 python bench.py --config C5 --steps 2 --warmup 3 > $O/r11_bench_C5.json 2> $O/r11_bench_C5.err; echo "bench C5 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r11_bench_C[345].json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
