#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py "tests/test_backbone_gpu.py::test_fp16_autocast_with_grad_scaler" tests/test_backbone_gpu.py -k "graph or fp16 or da_" -m gpu -q --maxfail=10 > gpurun_out/r14_new_tests.log 2>&1; echo "new tests rc=$?" >> gpurun_out/r14_new_tests.log; tail -30 gpurun_out/r14_new_tests.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/r14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r14_tests.log; tail -8 gpurun_out/r14_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/r14_bench.json 2> gpurun_out/r14_bench.err; head -c 600 gpurun_out/r14_bench.json; echo; tail -3 gpurun_out/r14_bench.err | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r14_smoke.log 2>&1; tail -2 gpurun_out/r14_smoke.log
