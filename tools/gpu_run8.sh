#!/bin/bash
mkdir -p gpurun_out
timeout 200 build/ffn_test check > gpurun_out/r8_ffn_check.log 2>&1; grep -c " ok" gpurun_out/r8_ffn_check.log; grep "FAIL\|non-finite\|PASSED\|FAILED" gpurun_out/r8_ffn_check.log | head -5
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r8_tests.log; tail -14 gpurun_out/r8_tests.log | cut -c1-300
timeout 600 python bench.py --config lsk_s --global-batch 4 --steps 5 --warmup 2 --no-cpu-baseline --no-gpu-eager > gpurun_out/r8_bench_lsk_n1.json 2> gpurun_out/r8_bench_lsk_n1.err; head -c 500 gpurun_out/r8_bench_lsk_n1.json; tail -3 gpurun_out/r8_bench_lsk_n1.err
timeout 900 python bench.py --config b_e16 --global-batch 2 --steps 3 --warmup 2 --no-cpu-baseline --no-gpu-eager > gpurun_out/r8_bench_b16_n1.json 2> gpurun_out/r8_bench_b16_n1.err; head -c 500 gpurun_out/r8_bench_b16_n1.json; tail -3 gpurun_out/r8_bench_b16_n1.err
