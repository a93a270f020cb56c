#!/bin/bash
# Round 2, call 9: weight stages multicast across clusters of two CTAs (conv_tc_kernel<BN, 2>): CHECK lines, per-layer timing
# against independent CTAs, role profile, the engine-level bit-identity test, bench with the option on.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 120 tools/conv_probe 37 2 0 0 0 3 0 2 > $O/r9_mc_check.log 2>&1; echo "mcast check rc=$? ok=$(grep -c ' ok' $O/r9_mc_check.log) $(grep TOTAL $O/r9_mc_check.log)"; grep -E "FAIL|bad=[1-9]|timeout|rror" $O/r9_mc_check.log | head -5
timeout 100 tools/conv_probe 37 2 1 0 0 3 0 0 > $O/r9_plain.log 2>&1; echo "plain: $(grep TOTAL $O/r9_plain.log)"
timeout 100 tools/conv_probe 37 2 1 0 0 3 0 2 > $O/r9_mc.log 2>&1; echo "mcast: $(grep TOTAL $O/r9_mc.log)"
paste <(grep TIME $O/r9_plain.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r9_mc.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_prof 37 2 1 0 0 3 0 2 > $O/r9_mc_prof.log 2>&1; grep "PROF " $O/r9_mc_prof.log | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q > $O/r9_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r9_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r9_bench_C2.json 2> $O/r9_bench_C2.err; echo "bench C2 rc=$?"
LM_WEIGHT_MCAST=2 timeout 400 python bench.py --steps 5 --warmup 3 > $O/r9_bench_C2_mc.json 2> $O/r9_bench_C2_mc.err; echo "bench C2 mcast rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r9_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
