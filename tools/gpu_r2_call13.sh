#!/bin/bash
# Round 2, call 13 (the round's last GPU minutes): stem_kernel_v3<IT, true> (next tile's samples fetched one tile ahead) as the
# default - the bit-identity tests of all kernel variants, the forward / golden parity tests, bench C2 with it and with the
# plain v3, ncu --set full of the stem launches.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 150 python -m pytest tests/test_gpu_zz_experimental.py tests/test_gpu_forward.py tests/test_gpu_goldens.py -m gpu -q -x > $O/r13_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r13_pytest.log | cut -c1-200
timeout 120 python bench.py --steps 5 --warmup 3 > $O/r13_bench_C2.json 2> $O/r13_bench_C2.err; echo "bench C2 rc=$?"
LM_STEM_V2=2 timeout 120 python bench.py --steps 5 --warmup 3 > $O/r13_bench_C2_stem2.json 2> $O/r13_bench_C2_stem2.err; echo "bench C2 plain v3 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r13_bench_*.json")):
    try:
        d = json.load(open(f))
        print(f, round(d["value"], 1), "e2e", round(d["e2e"]["value"], 1), "frac", round(d["roofline"]["frac"], 4), "dice", d.get("dice"), "flips", d.get("label_flips"), d["config"]["stage_ms_per_step"], d["clocks"])
    except Exception as e:
        print(f, "unreadable:", e)
PY
LM_GRAPHS=0 timeout 100 ncu --set full --clock-control none -k regex:'stem_kernel' -c 2 -f -o /tmp/r13_stem \
    python bench.py --steps 1 --warmup 0 > $O/r13_ncu_stem.log 2>&1; echo "ncu stem rc=$?"
ncu -i /tmp/r13_stem.ncu-rep --page raw --csv > $O/r13_stem_raw.csv 2>/dev/null; ls -la $O/r13_stem_raw.csv
