#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python tools/diag_lsk.py 512 768 > gpurun_out/r11_diag.log 2>&1; tail -80 gpurun_out/r11_diag.log | cut -c1-250
