#!/bin/bash
# Final measurement round (gpurun): probe checks, GPU tests, bench with one and two MMA issuers, then the ncu
# captures (launch list of one bench step, --set full on the convolution kernel and on the small kernels) with the
# faster issuer setting.  Everything lands in gpurun_out/.
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 120 tools/conv_probe 37 1 0 0 > $O/conv_probe_c1.log 2>&1; echo "probe c1 single rc=$?"
grep -E "FAIL|TOTAL|range" $O/conv_probe_c1.log | cut -c1-200
timeout 60 tools/conv_probe 37 1 1 1 > $O/conv_probe_c1_dual.log 2>&1; echo "c1 dual: $(grep TOTAL $O/conv_probe_c1_dual.log)"
timeout 60 tools/conv_probe 37 2 1 0 > $O/conv_probe_c2.log 2>&1; echo "c2 single: $(grep TOTAL $O/conv_probe_c2.log)"
timeout 60 tools/conv_probe 37 2 1 1 > $O/conv_probe_c2_dual.log 2>&1; echo "c2 dual: $(grep TOTAL $O/conv_probe_c2_dual.log)"
timeout 600 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu.log
LM_DUAL_ISSUE=0 timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_single.json 2> $O/bench_single.err; echo "bench single rc=$?"
LM_DUAL_ISSUE=1 timeout 400 python bench.py --steps 5 --warmup 3 > $O/bench_dual.json 2> $O/bench_dual.err; echo "bench dual rc=$?"
cat $O/bench_single.json $O/bench_dual.json | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l); print(d['config']['mma_issuers_per_cta'], 'issuer(s): value', round(d['value'], 1), 'e2e', round(d['e2e']['value'], 1), 'frac', round(d['roofline']['frac'], 4), 'stages', d['config']['stage_ms_per_step'])
"
BEST=$(python -c "
import json
a = json.load(open('$O/bench_single.json'))['value']; b = json.load(open('$O/bench_dual.json'))['value']
print(1 if b > a * 1.01 else 0)")
echo "profiling with LM_DUAL_ISSUE=$BEST" | tee $O/profile_setting.txt
export LM_DUAL_ISSUE=$BEST
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file $O/launches.csv \
    python bench.py --steps 1 --warmup 0 > $O/bench_under_ncu.json 2> $O/bench_under_ncu.err; echo "ncu launches rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel -s 24 -c 4 -f -o $O/prof_conv \
    python bench.py --steps 1 --warmup 0 > /dev/null 2> $O/ncu_full.err; echo "ncu full rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:'^(?!.*conv_tc).*' -s 40 -c 80 -f -o $O/prof_small \
    python tools/profile_small_kernels.py > $O/small.log 2> $O/ncu_small.err; echo "ncu small rc=$?"
ls -la $O | tail -20
