#!/bin/bash
mkdir -p gpurun_out
KG_TC_DEBUG=1 timeout 150 python -m pytest tests/test_net_gpu.py -q -x -k "cta_pair" 2>&1 | tail -25 > gpurun_out/exp13_unit.log
tail -12 gpurun_out/exp13_unit.log
nvidia-smi --query-gpu=utilization.gpu,memory.used --format=csv,noheader
