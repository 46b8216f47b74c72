#!/bin/bash
# round-2 GPU run #4: fused FFN with hoisted descriptors (timing + NaN-aware check), op tests, LSK variants
mkdir -p gpurun_out
timeout 600 build/ffn_test check > gpurun_out/r4_ffn_check.log 2>&1; echo "rc=$?" >> gpurun_out/r4_ffn_check.log
tail -30 gpurun_out/r4_ffn_check.log
timeout 300 build/ffn_test time > gpurun_out/r4_ffn_time.log 2>&1
cat gpurun_out/r4_ffn_time.log
timeout 600 python -m pytest tests/test_ops_gpu.py -m gpu -q -x -k "fused_ffn or layernorm_to or ep_plan" -s > gpurun_out/r4_ops.log 2>&1; tail -30 gpurun_out/r4_ops.log | cut -c1-400
timeout 900 python tools/diag_lsk.py 256 > gpurun_out/r4_diag_lsk.log 2>&1
cat gpurun_out/r4_diag_lsk.log | cut -c1-1500
timeout 300 compute-sanitizer --tool initcheck build/ffn_test check640 > gpurun_out/r4_initcheck.log 2>&1; tail -15 gpurun_out/r4_initcheck.log
