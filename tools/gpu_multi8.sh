#!/bin/bash
# 8-GPU lease: BASELINE config 4 (ConvNeXt-B, E = 16, all blocks MoE, experts sharded 2 per rank over NVLink peer memory) and
# config 3 (ConvNeXt-T e8t2, global batch 32 -> 4 images per GPU): graph-captured step with the flat gradient all-reduce vs
# eager launches under DDP, and the eager step without any gradient sync (isolates the cost of the all-reduce)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 800 $TR --nproc-per-node 8 --master-port 29621 bench.py --gpus 8 --config b_e16 --global-batch 16 --steps 4 --warmup 3 > gpurun_out/m8_b_e16.json 2> gpurun_out/m8_b_e16.err; head -c 800 gpurun_out/m8_b_e16.json; tail -3 gpurun_out/m8_b_e16.err | cut -c1-300
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT timeout 500 $TR --nproc-per-node 8 --master-port 29622 bench.py --gpus 8 --steps 8 --warmup 3 > gpurun_out/m8_t_e8_graph.out 2> gpurun_out/m8_t_e8_graph.err; grep '"metric"' gpurun_out/m8_t_e8_graph.out | head -c 500; grep -i "nvls" gpurun_out/m8_t_e8_graph.out gpurun_out/m8_t_e8_graph.err | head -3 | cut -c1-200; grep -v "NCCL INFO" gpurun_out/m8_t_e8_graph.err | tail -2 | cut -c1-300
timeout 500 $TR --nproc-per-node 8 --master-port 29623 bench.py --gpus 8 --steps 8 --warmup 3 --cuda-graph off > gpurun_out/m8_t_e8_eager_ddp.json 2> gpurun_out/m8_t_e8_eager_ddp.err; head -c 300 gpurun_out/m8_t_e8_eager_ddp.json
