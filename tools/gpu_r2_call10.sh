#!/bin/bash
# Round 2, call 10: the pair kernel with the looped-tap rule against one CTA per tile (are pairs worth it on the 64-channel
# layers?), packed conversions in the stem / upsample stores (bit-identity tests + bench).
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 120 tools/conv_probe 37 2 0 0 1 > $O/r10_pair_check.log 2>&1; echo "pair check rc=$? ok=$(grep -c ' ok' $O/r10_pair_check.log) $(grep TOTAL $O/r10_pair_check.log)"; grep -E "FAIL|bad=[1-9]|timeout|rror" $O/r10_pair_check.log | head -5
timeout 100 tools/conv_probe 37 2 1 0 0 > $O/r10_single.log 2>&1; echo "single: $(grep TOTAL $O/r10_single.log)"
timeout 100 tools/conv_probe 37 2 1 0 1 > $O/r10_pair.log 2>&1; echo "pair: $(grep TOTAL $O/r10_pair.log)"
timeout 100 tools/conv_probe 37 1 1 0 1 > $O/r10_pair_c1.log 2>&1; echo "pair chunk 1: $(grep TOTAL $O/r10_pair_c1.log)"
paste <(grep TIME $O/r10_single.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r10_pair.log | awk '{print $(NF-3)}') <(grep TIME $O/r10_pair_c1.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_prof 37 2 1 0 1 > $O/r10_pair_prof.log 2>&1; grep "PROF " $O/r10_pair_prof.log | grep "down0\|up3" | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_zz_experimental.py tests/test_gpu_forward.py tests/test_gpu_goldens.py tests/test_gpu_cli.py -m gpu -q > $O/r10_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r10_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r10_bench_C2.json 2> $O/r10_bench_C2.err; echo "bench C2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r10_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
