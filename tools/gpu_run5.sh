#!/bin/bash
mkdir -p gpurun_out
timeout 300 build/ffn_test check640 > gpurun_out/r5_check640.log 2>&1; cat gpurun_out/r5_check640.log | cut -c1-700
for cfg in "3 0" "3 4" "3 8" "3 12" "3 15"; do timeout 120 build/ffn_test one $cfg; done > gpurun_out/r5_ffn_knobs.log 2>&1
cat gpurun_out/r5_ffn_knobs.log
timeout 900 python -m pytest tests/test_lsk_gpu.py -m gpu -q -k "dwconv_generic or lsk_select or patch_embed or linear_gelu or colstat" 2>&1 | tail -25 | cut -c1-300 > gpurun_out/r5_lsk_ops.log; cat gpurun_out/r5_lsk_ops.log
