#!/bin/bash
# Round 2, call 12: stem_kernel_v4 (v3 in pixel pairs, 64 registers, four blocks per SM; commit 0fb7d46, removed after this
# call: bit-identical but 0.256 ms per wave against 0.187 ms, profiles/r02_call12_stem_kernel.md) as the default - smoke, the GPU suite without
# the two CPU-oracle-heavy full-size tests (3.3 of its 4.9 minutes; they ran on the call-11 tree and the stem variants are
# compared bit for bit inside test_gpu_zz_experimental.py), bench C2 with v4 and with v3, ncu --set full of the stem kernels.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 200 python __graft_entry__.py smoke > $O/r12_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r12_smoke.log | cut -c1-300
timeout 600 python -m pytest tests -m gpu -q -s --durations=5 \
    --deselect tests/test_gpu_fullsize.py::test_c2_bench_workload_against_oracle --deselect tests/test_gpu_fullsize.py::test_c3_c4_subvolume_against_oracle \
    > $O/r12_pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|Error" $O/r12_pytest_gpu.log | cut -c1-220 | tail -6
timeout 300 python bench.py --steps 5 --warmup 3 > $O/r12_bench_C2.json 2> $O/r12_bench_C2.err; echo "bench C2 rc=$?"
LM_STEM_V2=2 timeout 300 python bench.py --steps 5 --warmup 3 > $O/r12_bench_C2_stem3.json 2> $O/r12_bench_C2_stem3.err; echo "bench C2 stem_kernel_v3 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r12_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
LM_GRAPHS=0 timeout 200 ncu --set full --clock-control none -k regex:'stem_kernel' -c 2 -f -o /tmp/r12_stem \
    python bench.py --steps 1 --warmup 0 > $O/r12_ncu_stem.log 2>&1; echo "ncu stem rc=$?"
ncu -i /tmp/r12_stem.ncu-rep --page raw --csv > $O/r12_stem_raw.csv 2>/dev/null; ls -la $O/r12_stem_raw.csv
du -sh $O
