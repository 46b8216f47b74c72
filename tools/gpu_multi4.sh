#!/bin/bash
# 4-GPU lease: expert-parallel parity (2 ranks), cfg5 (LSKNet-S, SyncBN, 4 GPUs), cfg3 strong-scaling point + comm isolation,
# cfg4 rehearsal (ConvNeXt-B E16 expert-parallel on 4 ranks)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/m4_topo.txt 2>&1
timeout 600 python -m pytest tests/test_ep_gpu.py -m gpu -q -x > gpurun_out/m4_ep_test.log 2>&1; tail -5 gpurun_out/m4_ep_test.log | cut -c1-400
timeout 600 $TR --nproc-per-node 4 --master-port 29611 bench.py --gpus 4 --config lsk_s --steps 5 --warmup 3 > gpurun_out/m4_lsk_s.json 2> gpurun_out/m4_lsk_s.err; head -c 700 gpurun_out/m4_lsk_s.json; tail -2 gpurun_out/m4_lsk_s.err | cut -c1-300
timeout 600 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/m4_t_e8.json 2> gpurun_out/m4_t_e8.err; head -c 400 gpurun_out/m4_t_e8.json; tail -2 gpurun_out/m4_t_e8.err | cut -c1-300
timeout 600 $TR --nproc-per-node 4 --master-port 29613 bench.py --gpus 4 --steps 6 --warmup 3 --no-grad-sync > gpurun_out/m4_t_e8_nosync.json 2> gpurun_out/m4_t_e8_nosync.err; head -c 300 gpurun_out/m4_t_e8_nosync.json
SM3_RESERVE_SMS=8 timeout 600 $TR --nproc-per-node 4 --master-port 29614 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/m4_t_e8_reserve8.json 2> gpurun_out/m4_t_e8_reserve8.err; head -c 300 gpurun_out/m4_t_e8_reserve8.json
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 $TR --nproc-per-node 4 --master-port 29615 bench.py --gpus 4 --config b_e16 --steps 3 --warmup 2 > gpurun_out/m4_b_e16.json 2> gpurun_out/m4_b_e16.err; grep -v NCCL gpurun_out/m4_b_e16.json | head -c 700; grep -i "nvls\|algo" gpurun_out/m4_b_e16.err gpurun_out/m4_b_e16.json | head -5 | cut -c1-200; grep -v "NCCL INFO" gpurun_out/m4_b_e16.err | tail -4 | cut -c1-300
