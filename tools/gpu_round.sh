#!/bin/bash
# One GPU round (gpurun): convolution probe (checks + timings), the GPU test-suite, one bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/conv_probe 37 1 0 0 > gpurun_out/conv_probe_c1.log 2>&1; echo "probe c1 single rc=$?"
grep -E "CHECK|TOTAL|range" gpurun_out/conv_probe_c1.log | cut -c1-220
timeout 120 tools/conv_probe 37 1 0 1 > gpurun_out/conv_probe_c1_dual.log 2>&1; echo "probe c1 dual rc=$?"
grep -E "FAIL|TOTAL|range" gpurun_out/conv_probe_c1_dual.log | cut -c1-220
timeout 60 tools/conv_probe 37 2 1 0 > gpurun_out/conv_probe_c2.log 2>&1; echo "c2 single: $(grep TOTAL gpurun_out/conv_probe_c2.log)"
timeout 60 tools/conv_probe 37 2 1 1 > gpurun_out/conv_probe_c2_dual.log 2>&1; echo "c2 dual: $(grep TOTAL gpurun_out/conv_probe_c2_dual.log)"
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
