#!/bin/bash
# A/B on one B200, same session: fused FFN forward vs the GEMM -> act_pack -> GEMM forward (SM3_FUSED_FFN=0), and the
# micro-batch size of the cfg3 step (global 32 as 4x8, 2x16, 1x32)
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-eager --steps 5 --warmup 3"
for tag in fused unfused; do
  if [ $tag = unfused ]; then export SM3_FUSED_FFN=0; else unset SM3_FUSED_FFN; fi
  timeout 300 python bench.py $B --global-batch 8 > gpurun_out/ab_gb8_$tag.json 2> gpurun_out/ab_gb8_$tag.err; head -c 250 gpurun_out/ab_gb8_$tag.json; echo
done
unset SM3_FUSED_FFN
for mb in 16 32; do
  timeout 300 python bench.py $B --micro-batch $mb > gpurun_out/ab_gb32_mb$mb.json 2> gpurun_out/ab_gb32_mb$mb.err; head -c 250 gpurun_out/ab_gb32_mb$mb.json; echo; tail -1 gpurun_out/ab_gb32_mb$mb.err | cut -c1-200
done
