#!/bin/bash
# Round 2, call 6: epilogue with per-warp constants in shared memory + setmaxnreg register reallocation (88 / 208):
# correctness (CHECK), timing against the same code without setmaxnreg, 72 / 216 split, direct stores, dual issuers,
# issuer time split (PROF2), then the fast GPU tests (forward, experimental, CLI) and the bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 tools/conv_probe 37 2 0 0 0 > $O/r6_single_check.log 2>&1; echo "single check rc=$? ok=$(grep -c ' ok' $O/r6_single_check.log) $(grep TOTAL $O/r6_single_check.log)"; grep -E "FAIL|bad=[1-9]" $O/r6_single_check.log | head -5
timeout 200 tools/conv_probe 37 2 0 0 1 > $O/r6_pair_check.log 2>&1; echo "pair check rc=$? ok=$(grep -c ' ok' $O/r6_pair_check.log) $(grep TOTAL $O/r6_pair_check.log)"; grep -E "FAIL|bad=[1-9]" $O/r6_pair_check.log | head -5
for v in conv_probe conv_probe_nosmx conv_probe_r216 conv_probe_d; do timeout 100 tools/$v 37 2 1 0 0 > $O/r6_$v.log 2>&1; echo "$v: $(grep TOTAL $O/r6_$v.log)"; done
timeout 100 tools/conv_probe 37 2 1 1 0 > $O/r6_dual.log 2>&1; echo "dual issue: $(grep TOTAL $O/r6_dual.log)"
timeout 100 tools/conv_probe 37 1 1 0 0 > $O/r6_chunk1.log 2>&1; echo "chunk 1: $(grep TOTAL $O/r6_chunk1.log)"
paste <(grep TIME $O/r6_conv_probe.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r6_conv_probe_nosmx.log | awk '{print $(NF-3)}') <(grep TIME $O/r6_conv_probe_r216.log | awk '{print $(NF-3)}') <(grep TIME $O/r6_conv_probe_d.log | awk '{print $(NF-3)}') <(grep TIME $O/r6_dual.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_prof 37 2 1 0 0 > $O/r6_prof.log 2>&1; grep PROF $O/r6_prof.log | cut -c1-260
timeout 600 python -m pytest tests/test_gpu_zz_experimental.py tests/test_gpu_forward.py tests/test_gpu_cli.py tests/test_gpu_goldens.py -m gpu -q > $O/r6_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r6_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r6_bench_C2.json 2> $O/r6_bench_C2.err; echo "bench C2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r6_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
