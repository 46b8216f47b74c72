#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/diag_lsk.py > gpurun_out/r7_diag_lsk.log 2>&1; cat gpurun_out/r7_diag_lsk.log | cut -c1-900
timeout 200 build/ffn_test check > gpurun_out/r7_ffn_check.log 2>&1; grep -c " ok" gpurun_out/r7_ffn_check.log; grep "FAIL\|non-finite\|PASSED\|FAILED" gpurun_out/r7_ffn_check.log | head
timeout 900 python -m pytest tests/test_backbone_gpu.py tests/test_ops_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -12 | cut -c1-250 > gpurun_out/r7_tests_default.log; cat gpurun_out/r7_tests_default.log
SM3_FUSED_BWD=1 timeout 900 python -m pytest tests/test_backbone_gpu.py -m gpu -q --maxfail=8 2>&1 | tail -12 | cut -c1-250 > gpurun_out/r7_tests_trio.log; cat gpurun_out/r7_tests_trio.log
timeout 600 python bench.py --global-batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r7_bench_gb8.json 2> gpurun_out/r7_bench_gb8.err; head -c 400 gpurun_out/r7_bench_gb8.json; tail -2 gpurun_out/r7_bench_gb8.err
