#!/bin/bash
# One GPU round (gpurun): ablation timings of the convolution kernel, the GPU test-suite, one bench line.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for e in 0 1 16 2 4 8 32 3 7 39; do
  if [ -x tools/conv_probe_exp$e ]; then
    timeout 60 tools/conv_probe_exp$e 37 1 1 > gpurun_out/exp$e.log 2>&1
    echo "EXP $e: $(grep TOTAL gpurun_out/exp$e.log)"
  fi
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"
cat gpurun_out/bench.json
