#!/bin/bash
# Validation order for the paths that were prepared without GPU time.  Run under gpurun,
# after tools/build_probes.sh (here) and python -m lungmask_b200.build.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
# 0. the cta_group::2 conventions the pair kernel relies on (seconds; run tools/build_probes.sh locally first)
timeout 60 tools/pair_probe > $O/pair_probe.log 2>&1; echo "pair_probe rc=$?"; cat $O/pair_probe.log
# 1. CTA-pair convolution kernel: bit-exact integer checks first (layout / barrier mistakes), then float checks + timing
timeout 180 tools/conv_probe 37 1 0 0 1 > $O/pair_probe_c1.log 2>&1; echo "pair probe rc=$?"
grep -E "CHECK|TOTAL|timeout|error" $O/pair_probe_c1.log | cut -c1-200
timeout 60 tools/conv_probe 37 1 1 0 0 > $O/single_probe_c1.log 2>&1; echo "single: $(grep TOTAL $O/single_probe_c1.log)"
timeout 60 tools/conv_probe 37 2 1 0 1 > $O/pair_probe_c2.log 2>&1; echo "pair c2: $(grep TOTAL $O/pair_probe_c2.log)"
# 2. every experimental path must reproduce the default path bit for bit
LM_TEST_EXPERIMENTAL=1 timeout 600 python -m pytest tests/test_gpu_zz_experimental.py -m gpu -q > $O/pytest_experimental.log 2>&1; echo "experimental pytest rc=$?"
tail -5 $O/pytest_experimental.log
# 3. bench lines: default, then each option
timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_default.json 2> $O/bench_default.err
for opt in LM_STEM_V2 LM_CCL_REDUCED LM_CTA_PAIRS; do
  env $opt=1 timeout 300 python bench.py --steps 5 --warmup 3 > $O/bench_$opt.json 2> $O/bench_$opt.err; echo "$opt rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), d["config"].get("engine_env_options"), d["config"]["stage_ms_per_step"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
