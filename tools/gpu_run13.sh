#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_graph_gpu.py tests/test_lsk_gpu.py -m gpu -q --maxfail=10 > gpurun_out/r13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r13_tests.log; tail -25 gpurun_out/r13_tests.log | cut -c1-300
B="--no-cpu-baseline --no-gpu-eager"
timeout 600 python bench.py $B --cuda-graph on > gpurun_out/r13_t_e8_graph.json 2> gpurun_out/r13_t_e8_graph.err; head -c 400 gpurun_out/r13_t_e8_graph.json; echo; tail -3 gpurun_out/r13_t_e8_graph.err | cut -c1-300
timeout 600 python bench.py $B --cuda-graph off > gpurun_out/r13_t_e8_eager.json 2> gpurun_out/r13_t_e8_eager.err; head -c 400 gpurun_out/r13_t_e8_eager.json; echo
timeout 600 python bench.py $B --global-batch 8 --cuda-graph on > gpurun_out/r13_t_e8_gb8_graph.json 2> gpurun_out/r13_t_e8_gb8_graph.err; head -c 400 gpurun_out/r13_t_e8_gb8_graph.json; echo
timeout 600 python bench.py $B --global-batch 4 --micro-batch 4 --cuda-graph on > gpurun_out/r13_t_e8_gb4_graph.json 2> gpurun_out/r13_t_e8_gb4_graph.err; head -c 400 gpurun_out/r13_t_e8_gb4_graph.json; echo
timeout 600 python bench.py $B --global-batch 4 --micro-batch 4 --cuda-graph off > gpurun_out/r13_t_e8_gb4_eager.json 2> gpurun_out/r13_t_e8_gb4_eager.err; head -c 400 gpurun_out/r13_t_e8_gb4_eager.json; echo
timeout 600 python bench.py $B --config lsk_s --global-batch 4 --cuda-graph on > gpurun_out/r13_lsk_graph.json 2> gpurun_out/r13_lsk_graph.err; head -c 400 gpurun_out/r13_lsk_graph.json; echo; tail -3 gpurun_out/r13_lsk_graph.err | cut -c1-300
timeout 600 python bench.py $B --config lsk_s --global-batch 4 --cuda-graph off > gpurun_out/r13_lsk_eager.json 2> gpurun_out/r13_lsk_eager.err; head -c 400 gpurun_out/r13_lsk_eager.json; echo
timeout 900 python bench.py $B --config b_e16 --global-batch 2 --cuda-graph on --steps 4 --warmup 3 > gpurun_out/r13_b16_graph.json 2> gpurun_out/r13_b16_graph.err; head -c 400 gpurun_out/r13_b16_graph.json; echo; tail -3 gpurun_out/r13_b16_graph.err | cut -c1-300
