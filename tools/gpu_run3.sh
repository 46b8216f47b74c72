#!/bin/bash
# round-2 GPU run #3: fused-FFN perf knobs + ncu, full GPU suite (fused path on), LSK-S diagnostic, bench
mkdir -p gpurun_out
for cfg in "3 0" "1 0" "3 1" "3 3"; do timeout 120 build/ffn_test one $cfg; done > gpurun_out/r3_ffn_knobs.log 2>&1
cat gpurun_out/r3_ffn_knobs.log
i=0
for skip in 2 14 26; do
  i=$((i+1))
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:ffn_ --launch-skip $skip --launch-count 1 -f -o gpurun_out/r3_ffn_k$i build/ffn_test one > gpurun_out/r3_ncu_k$i.log 2>&1
  if [ -f gpurun_out/r3_ffn_k$i.ncu-rep ]; then python profiles/summarize_ncu.py gpurun_out/r3_ffn_k$i.ncu-rep > gpurun_out/r3_ncu_ffn_k$i.txt 2>&1; rm -f gpurun_out/r3_ffn_k$i.ncu-rep; fi
done
timeout 2400 python -m pytest tests -m gpu -q --maxfail=12 -s > gpurun_out/r3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r3_tests.log
tail -15 gpurun_out/r3_tests.log
timeout 900 python tools/diag_lsk.py 256 1024 > gpurun_out/r3_diag_lsk.log 2>&1
cat gpurun_out/r3_diag_lsk.log | tail -70
timeout 600 python bench.py --global-batch 8 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-eager > gpurun_out/r3_bench_fused_gb8.json 2> gpurun_out/r3_bench_fused_gb8.err
head -c 700 gpurun_out/r3_bench_fused_gb8.json; tail -3 gpurun_out/r3_bench_fused_gb8.err
