#!/bin/bash
# round-2 GPU run #2: standalone fused-FFN check + timing, then the whole GPU suite
mkdir -p gpurun_out
timeout 600 build/ffn_test check > gpurun_out/r2_ffn_check.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ffn_check.log
tail -25 gpurun_out/r2_ffn_check.log
timeout 300 build/ffn_test time > gpurun_out/r2_ffn_time.log 2>&1; echo "rc=$?" >> gpurun_out/r2_ffn_time.log
cat gpurun_out/r2_ffn_time.log
SM3_FUSED_FFN=0 timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 -s > gpurun_out/r2_tests_unfused.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_unfused.log
tail -5 gpurun_out/r2_tests_unfused.log
if grep -q "FFN TEST PASSED" gpurun_out/r2_ffn_check.log; then
  timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_ops_gpu.py -m gpu -q --maxfail=10 -s > gpurun_out/r2_tests_fused.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2_tests_fused.log
  tail -5 gpurun_out/r2_tests_fused.log
  timeout 600 python bench.py --global-batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r2_bench_fused_gb8.json 2> gpurun_out/r2_bench_fused_gb8.err
  head -c 600 gpurun_out/r2_bench_fused_gb8.json
fi
