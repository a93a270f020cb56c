#!/bin/bash
# Round 2, call 7: (1) issue loop with the nine taps as a LOOP instead of unrolled (instruction footprint 9080 -> 6808 /
# 6376 -> 4240 instructions: does instruction fetch starve the issuer on the epilogue-heavy 64-channel layers?);
# (2) BN = 128 with two epilogue groups of four warps (128 accumulator columns per thread, 72 / 216 registers);
# (3) CLI test, bench line.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 tools/conv_probe_g2 37 2 0 0 0 > $O/r7_g2_check.log 2>&1; echo "g2 check rc=$? ok=$(grep -c ' ok' $O/r7_g2_check.log) $(grep TOTAL $O/r7_g2_check.log)"; grep -E "FAIL|bad=[1-9]|timeout" $O/r7_g2_check.log | head -5
timeout 200 tools/conv_probe_tu0 37 2 0 0 0 > $O/r7_tu0_check.log 2>&1; echo "tu0 check rc=$? ok=$(grep -c ' ok' $O/r7_tu0_check.log) $(grep TOTAL $O/r7_tu0_check.log)"; grep -E "FAIL|bad=[1-9]|timeout" $O/r7_tu0_check.log | head -5
for v in conv_probe conv_probe_tu0 conv_probe_g2; do timeout 100 tools/$v 37 2 1 0 0 > $O/r7_$v.log 2>&1; echo "$v: $(grep TOTAL $O/r7_$v.log)"; done
paste <(grep TIME $O/r7_conv_probe.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r7_conv_probe_tu0.log | awk '{print $(NF-3)}') <(grep TIME $O/r7_conv_probe_g2.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_tu0_prof 37 2 1 0 0 > $O/r7_tu0_prof.log 2>&1; grep PROF $O/r7_tu0_prof.log | grep "down0.block3\|down1.block0\|down4.block3\|up3.block" | cut -c1-260
timeout 120 tools/conv_probe_g2_prof 37 2 1 0 0 > $O/r7_g2_prof.log 2>&1; grep PROF $O/r7_g2_prof.log | grep "down1\|down2.block0\|down4.block3\|up2.block3" | cut -c1-260
timeout 300 python -m pytest tests/test_gpu_cli.py -m gpu -q > $O/r7_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r7_pytest.log
timeout 400 python bench.py --steps 5 --warmup 3 > $O/r7_bench_C2.json 2> $O/r7_bench_C2.err; echo "bench C2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r7_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
