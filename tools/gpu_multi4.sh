#!/bin/bash
# 4-GPU lease: expert-parallel parity (2 ranks), cfg5 (LSKNet-S, SyncBN, 4 GPUs) graph-captured vs eager DDP,
# cfg4 rehearsal (ConvNeXt-B E16 expert-parallel on 4 ranks) eager and graph-captured
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi topo -m > gpurun_out/m4_topo.txt 2>&1
timeout 600 python -m pytest tests/test_ep_gpu.py -m gpu -q -x > gpurun_out/m4_ep_test.log 2>&1; tail -5 gpurun_out/m4_ep_test.log | cut -c1-400
timeout 500 $TR --nproc-per-node 4 --master-port 29611 bench.py --gpus 4 --config lsk_s --steps 6 --warmup 3 > gpurun_out/m4_lsk_s_graph.json 2> gpurun_out/m4_lsk_s_graph.err; head -c 700 gpurun_out/m4_lsk_s_graph.json; tail -2 gpurun_out/m4_lsk_s_graph.err | cut -c1-300
timeout 500 $TR --nproc-per-node 4 --master-port 29612 bench.py --gpus 4 --config lsk_s --steps 6 --warmup 3 --cuda-graph off > gpurun_out/m4_lsk_s_eager.json 2> gpurun_out/m4_lsk_s_eager.err; head -c 400 gpurun_out/m4_lsk_s_eager.json; tail -2 gpurun_out/m4_lsk_s_eager.err | cut -c1-300
timeout 700 $TR --nproc-per-node 4 --master-port 29615 bench.py --gpus 4 --config b_e16 --steps 3 --warmup 3 > gpurun_out/m4_b_e16.json 2> gpurun_out/m4_b_e16.err; head -c 700 gpurun_out/m4_b_e16.json; tail -3 gpurun_out/m4_b_e16.err | cut -c1-300
timeout 500 $TR --nproc-per-node 4 --master-port 29617 bench.py --gpus 4 --steps 6 --warmup 3 > gpurun_out/m4_t_e8_graph.json 2> gpurun_out/m4_t_e8_graph.err; head -c 400 gpurun_out/m4_t_e8_graph.json; tail -2 gpurun_out/m4_t_e8_graph.err | cut -c1-300
