#!/bin/bash
# Round 2, call 1: validate the kernels prepared at the end of round 1 (cta_group::2 pairs, stem_v2), the new
# host-free post-processing, the new parity tests; then bench lines with each option.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee $O/r2c1_gpu.txt
timeout 60 tools/pair_probe > $O/pair_probe.log 2>&1; echo "pair_probe rc=$?"; tail -30 $O/pair_probe.log
timeout 180 tools/conv_probe 37 1 0 0 1 > $O/pair_probe_c1.log 2>&1; echo "pair conv probe rc=$?"
grep -E "CHECK|TOTAL|timeout|error" $O/pair_probe_c1.log | cut -c1-200
timeout 60 tools/conv_probe 37 1 1 0 0 > $O/single_probe_c1.log 2>&1; echo "single: $(grep TOTAL $O/single_probe_c1.log)"
timeout 60 tools/conv_probe 37 2 1 0 1 > $O/pair_probe_c2.log 2>&1; echo "pair c2: $(grep TOTAL $O/pair_probe_c2.log)"
timeout 60 tools/conv_probe 37 2 1 0 0 > $O/single_probe_c2.log 2>&1; echo "single c2: $(grep TOTAL $O/single_probe_c2.log)"
LM_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q > $O/pytest_experimental.log 2>&1; echo "experimental pytest rc=$?"
tail -15 $O/pytest_experimental.log
timeout 1200 python -m pytest tests -m gpu -q -s --durations=15 > $O/pytest_gpu_r2c1.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|error|Error|C2 |C3 |C4 |end to end|fused|golden|differing|max\|dscore" $O/pytest_gpu_r2c1.log | cut -c1-220 | tail -60
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err; echo "bench rc=$?"
for opt in LM_STEM_V2 LM_UPSAMPLE_V2 LM_CTA_PAIRS; do
  env $opt=1 timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_$opt.json 2> $O/bench_$opt.err; echo "$opt rc=$?"
done
LM_CCL_RULE=0 timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_LM_CCL_RULE0.json 2> $O/bench_LM_CCL_RULE0.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), d["config"].get("engine_env_options"), d["config"]["stage_ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
