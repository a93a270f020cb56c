#!/bin/bash
# ncu --set full with source correlation on two launches of the convolution kernel (conv_probe, timing-only mode):
# up0.block0 (BN=128) and up3.block0 (BN=64).  Reports land in gpurun_out/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 41 --launch-count 1 -f -o gpurun_out/ncu_conv_bn128 tools/conv_probe 37 1 1 > gpurun_out/ncu_bn128.log 2>&1; echo "rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 77 --launch-count 1 -f -o gpurun_out/ncu_conv_bn64 tools/conv_probe 37 1 1 > gpurun_out/ncu_bn64.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
