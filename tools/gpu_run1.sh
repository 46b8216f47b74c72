#!/bin/bash
# round-2 GPU run #1: parity suite at the benchmarked shapes + bench lines with the new comparators + launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/r1_gpu.txt; nproc >> gpurun_out/r1_gpu.txt; free -g >> gpurun_out/r1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -q --maxfail=12 -x -s > gpurun_out/r1_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r1_tests.log
timeout 600 python bench.py --global-batch 8 --steps 10 --warmup 3 > gpurun_out/r1_bench_gb8.json 2> gpurun_out/r1_bench_gb8.err
timeout 600 python bench.py --no-cpu-baseline --no-gpu-eager > gpurun_out/r1_bench_gb32.json 2> gpurun_out/r1_bench_gb32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_launches.csv \
   python bench.py --global-batch 8 --steps 1 --warmup 1 --no-cpu-baseline --no-gpu-eager > gpurun_out/r1_ncu.log 2>&1
python profiles/summarize_launches.py gpurun_out/r1_launches.csv 2 > gpurun_out/r1_launches.txt 2>&1
tail -3 gpurun_out/r1_tests.log; cat gpurun_out/r1_bench_gb8.json | head -c 3000
