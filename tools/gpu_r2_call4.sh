#!/bin/bash
# Round 2, call 4: epilogue variants of the convolution kernel (a: TMA stores; b: + prefetch of the per-channel constants;
# c: direct stores + prefetch), stem_kernel_v3 (bit-identity + timing + ncu), full GPU suite, bench lines.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
for v in a b c d; do timeout 100 tools/conv_probe_$v 37 2 1 0 0 > $O/r4_probe_$v.log 2>&1; echo "variant $v: $(grep TOTAL $O/r4_probe_$v.log)"; done
timeout 200 tools/conv_probe 37 2 0 0 0 > $O/r4_probe_default.log 2>&1; grep -E "FAIL|TOTAL" $O/r4_probe_default.log | cut -c1-160
paste <(grep TIME $O/r4_probe_a.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r4_probe_b.log | awk '{print $(NF-3)}') <(grep TIME $O/r4_probe_c.log | awk '{print $(NF-3)}') <(grep TIME $O/r4_probe_d.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_prof 37 2 1 0 0 > $O/r4_prof_default_c2.log 2>&1; grep PROF $O/r4_prof_default_c2.log | cut -c1-250 | head -8
timeout 300 python -m pytest tests/test_gpu_zz_experimental.py tests/test_gpu_forward.py -m gpu -q -s > $O/r4_pytest_fast.log 2>&1; echo "fast pytest rc=$?"; tail -3 $O/r4_pytest_fast.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r4_bench_C2.json 2> $O/r4_bench_C2.err; echo "bench C2 rc=$?"
LM_STEM_V2=1 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r4_bench_C2_stemv2.json 2> $O/r4_bench_C2_stemv2.err
LM_MERGE_CTAS=1 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r4_bench_C2_merge1.json 2> $O/r4_bench_C2_merge1.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r4_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
LM_GRAPHS=0 timeout 300 ncu --set full --clock-control none -k regex:'stem_kernel' -s 2 -c 2 -f -o /tmp/r4_stem python tools/profile_small_kernels.py > /dev/null 2> $O/r4_ncu_stem.err
ncu -i /tmp/r4_stem.ncu-rep --page raw --csv > $O/r4_stem_raw.csv 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -s --durations=8 > $O/pytest_gpu_r2c4.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error" $O/pytest_gpu_r2c4.log | tail -5
du -sh $O
