#!/bin/bash
# Final validation on one B200 (gpurun -- bash tools/gpu_validate.sh): new tests first, the whole -m gpu suite, default bench, smoke
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_graph_gpu.py "tests/test_backbone_gpu.py::test_fp16_autocast_with_grad_scaler" tests/test_backbone_gpu.py -k "graph or fp16 or da_" -m gpu -q --maxfail=10 > gpurun_out/val_new_tests.log 2>&1; echo "new tests rc=$?" >> gpurun_out/val_new_tests.log; tail -30 gpurun_out/val_new_tests.log | cut -c1-300
timeout 1500 python -m pytest tests -m gpu -q --maxfail=10 > gpurun_out/val_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/val_tests.log; tail -8 gpurun_out/val_tests.log | cut -c1-300
timeout 900 python bench.py > gpurun_out/val_bench.json 2> gpurun_out/val_bench.err; head -c 600 gpurun_out/val_bench.json; echo; tail -3 gpurun_out/val_bench.err | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/val_smoke.log 2>&1; tail -2 gpurun_out/val_smoke.log
