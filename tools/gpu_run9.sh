#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_lsk_gpu.py -m gpu -q --maxfail=20 > gpurun_out/r9_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r9_tests.log; tail -25 gpurun_out/r9_tests.log | cut -c1-400
timeout 900 python tools/diag_lsk.py 256 768 > gpurun_out/r9_diag.log 2>&1; tail -60 gpurun_out/r9_diag.log | cut -c1-250
