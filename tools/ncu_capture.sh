#!/bin/bash
# usage: tools/ncu_capture.sh <tag> "<kernel-regex>:<skip>:<count>" ...   (run on the GPU box via gpurun)
# One `ncu --set full` pass per spec over one fwd+bwd step of bench.py; each report is condensed to text with
# profiles/summarize_ncu.py on the box and the (large) .ncu-rep is deleted so gpurun_out/ stays under its 64 MiB cap.
tag=$1; shift
mkdir -p gpurun_out
for spec in "$@"; do
  k=${spec%%:*}; r=${spec#*:}; s=${r%%:*}; c=${r#*:}
  name=${tag}_$(echo $k | tr -c 'A-Za-z0-9_\n' '_')_$s
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip $s --launch-count $c \
      -f -o gpurun_out/$name python bench.py --steps 1 --warmup 0 --global-batch 8 --cuda-graph off --no-cpu-baseline --no-gpu-eager > gpurun_out/$name.log 2>&1
  if [ -f gpurun_out/$name.ncu-rep ]; then
    python profiles/summarize_ncu.py gpurun_out/$name.ncu-rep > gpurun_out/$name.txt 2>&1
    rm -f gpurun_out/$name.ncu-rep
  fi
done
