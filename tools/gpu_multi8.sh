#!/bin/bash
# 8-GPU lease: BASELINE config 4 (ConvNeXt-B, E = 16, all blocks MoE, experts sharded 2 per rank over NVLink peer memory) and
# config 3 (ConvNeXt-T e8t2, global batch 32 -> 4 images per GPU) with the gradient all-reduce isolated
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --config b_e16 --global-batch 16 --steps 4 --warmup 3 > gpurun_out/m8_b_e16.json 2> gpurun_out/m8_b_e16.err; head -c 800 gpurun_out/m8_b_e16.json; tail -3 gpurun_out/m8_b_e16.err | cut -c1-300
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 600 $TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/m8_t_e8.out 2> gpurun_out/m8_t_e8.err; grep '"metric"' gpurun_out/m8_t_e8.out | head -c 400; grep -i "nvls" gpurun_out/m8_t_e8.out gpurun_out/m8_t_e8.err | head -3 | cut -c1-200
timeout 600 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --steps 8 --warmup 3 --no-grad-sync > gpurun_out/m8_t_e8_nosync.json 2> gpurun_out/m8_t_e8_nosync.err; head -c 300 gpurun_out/m8_t_e8_nosync.json
SM3_RESERVE_SMS=8 timeout 600 $TR --nproc-per-node 8 --master-port 29624 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/m8_t_e8_reserve8.json 2> gpurun_out/m8_t_e8_reserve8.err; head -c 300 gpurun_out/m8_t_e8_reserve8.json
