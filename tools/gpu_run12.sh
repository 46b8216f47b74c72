#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r12_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r12_tests.log; tail -12 gpurun_out/r12_tests.log | cut -c1-300
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k "regex:gemm_bf16x3|ffn_chain" --launch-count 126 --csv --log-file gpurun_out/r12_traffic.csv python bench.py --steps 1 --warmup 0 --global-batch 8 --no-cpu-baseline --no-gpu-eager > gpurun_out/r12_traffic.log 2>&1
python tools/gemm_traffic.py gpurun_out/r12_traffic.csv gpurun_out/r12_gemm_traffic.json 2>&1 | tail -12
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r12_launches.csv python bench.py --steps 1 --warmup 1 --global-batch 8 --no-cpu-baseline --no-gpu-eager > gpurun_out/r12_launches.log 2>&1
python profiles/summarize_launches.py gpurun_out/r12_launches.csv > gpurun_out/r12_launches.txt 2>&1; head -30 gpurun_out/r12_launches.txt | cut -c1-200
timeout 900 python bench.py > gpurun_out/r12_bench.json 2> gpurun_out/r12_bench.err; head -c 1500 gpurun_out/r12_bench.json; tail -3 gpurun_out/r12_bench.err
