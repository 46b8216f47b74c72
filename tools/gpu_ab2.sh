#!/bin/bash
# cross-check of the micro-batch split (gradient checksum mb32 vs mb8, deterministic routing) and the noisy-gating default
mkdir -p gpurun_out
B="--no-cpu-baseline --no-gpu-eager --steps 2 --warmup 3"
timeout 120 python bench.py $B --micro-batch 32 > gpurun_out/ab2_mb32.json 2> gpurun_out/ab2_mb32.err; python -c "import json;j=json.load(open('gpurun_out/ab2_mb32.json'));print('mb32',j['value'],j['grad_l1'],j['step_scalar'])"
timeout 120 python bench.py $B --micro-batch 8 > gpurun_out/ab2_mb8.json 2> gpurun_out/ab2_mb8.err; python -c "import json;j=json.load(open('gpurun_out/ab2_mb8.json'));print('mb8',j['value'],j['grad_l1'],j['step_scalar'])"
timeout 120 python bench.py $B --micro-batch 32 --noisy-gating config > gpurun_out/ab2_mb32_noisy.json 2> gpurun_out/ab2_mb32_noisy.err; python -c "import json;j=json.load(open('gpurun_out/ab2_mb32_noisy.json'));print('mb32 noisy',j['value'],j['grad_l1'],j['step_scalar'],j['cuda_graph'])"; tail -2 gpurun_out/ab2_mb32_noisy.err | cut -c1-200
