#!/bin/bash
# Round 2, call 5: (1) the CTA-pair kernel with cheaper remote arrives (LM_PAIR_ARRIVE 0 = .release.cluster as measured in
# call 1, 1 = default semantics, 2 = .relaxed.cluster; p1s = 1 + direct stores -> 4 / 7 weight stages);
# (2) epilogue variant d; (3) the 32-bit index arithmetic / two-CTA body mask / I/O changes: GPU suite + bench lines.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 tools/conv_probe_p1 37 2 0 0 1 > $O/r5_pair_p1_check.log 2>&1; echo "pair p1 check rc=$?"; grep -E "FAIL|bad=[1-9]|TOTAL" $O/r5_pair_p1_check.log | cut -c1-200
grep -c " ok" $O/r5_pair_p1_check.log
for v in p0 p1 p2 p1s; do timeout 100 tools/conv_probe_$v 37 2 1 0 1 > $O/r5_pair_$v.log 2>&1; echo "pair $v chunk 2: $(grep TOTAL $O/r5_pair_$v.log)"; done
timeout 100 tools/conv_probe_p1 37 1 1 0 1 > $O/r5_pair_p1_c1.log 2>&1; echo "pair p1 chunk 1: $(grep TOTAL $O/r5_pair_p1_c1.log)"
timeout 100 tools/conv_probe_p1 37 2 1 0 0 > $O/r5_single.log 2>&1; echo "single chunk 2: $(grep TOTAL $O/r5_single.log)"
paste <(grep TIME $O/r5_single.log | awk '{print $2, $(NF-3)}') <(grep TIME $O/r5_pair_p0.log | awk '{print $(NF-3)}') <(grep TIME $O/r5_pair_p1.log | awk '{print $(NF-3)}') <(grep TIME $O/r5_pair_p2.log | awk '{print $(NF-3)}') <(grep TIME $O/r5_pair_p1s.log | awk '{print $(NF-3)}')
timeout 120 tools/conv_probe_p1_prof 37 2 1 0 1 > $O/r5_pair_p1_prof.log 2>&1; grep PROF $O/r5_pair_p1_prof.log | cut -c1-250
timeout 200 tools/conv_probe_d 37 2 0 0 0 > $O/r5_probe_d.log 2>&1; echo "variant d rc=$?: $(grep -E 'TOTAL' $O/r5_probe_d.log)"; grep -E "FAIL|bad=[1-9]" $O/r5_probe_d.log | head -5
timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_goldens.py tests/test_gpu_forward.py tests/test_gpu_fullsize.py -m gpu -q -x > $O/r5_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r5_pytest.log
for c in C2 C4; do timeout 400 python bench.py --config $c --steps 5 --warmup 3 > $O/r5_bench_$c.json 2> $O/r5_bench_$c.err; echo "bench $c rc=$?"; done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r5_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
du -sh $O
