#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 4 6; do echo "== SM3_WDEBUG=$d"; SM3_WDEBUG=$d timeout 120 build/ffn_test check640 2>&1 | grep -A4 "wgrd\|non-finite" | cut -c1-300; done > gpurun_out/r6_wdebug.log 2>&1; cat gpurun_out/r6_wdebug.log
timeout 120 build/mma_bench > gpurun_out/r6_mma_bench.log 2>&1; cat gpurun_out/r6_mma_bench.log
