#!/bin/bash
# Quick GPU check of the convolution kernel (gpurun): per-layer correctness + timing, then whole-net score error
# for the chunk policies.  Outputs under gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 tools/conv_probe 37 1 > gpurun_out/conv_probe_c1.log 2>&1; echo "probe c1 rc=$?"
timeout 60 tools/conv_probe 37 2 1 > gpurun_out/conv_probe_c2.log 2>&1; echo "probe c2 rc=$?"
timeout 60 tools/conv_probe_prof 37 1 1 > gpurun_out/conv_probe_prof_c1.log 2>&1; echo "prof rc=$?"
grep -E "CHECK|TOTAL|range" gpurun_out/conv_probe_c1.log | cut -c1-200
grep -E "TOTAL" gpurun_out/conv_probe_c2.log
timeout 300 python tools/debug_gpu.py > gpurun_out/debug_gpu.log 2>&1; echo "debug rc=$?"
tail -12 gpurun_out/debug_gpu.log
