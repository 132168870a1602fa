#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_net_gpu.py -q -x -k "heads_l2" 2>&1 | tail -15 > gpurun_out/exp2_unit.log
tail -3 gpurun_out/exp2_unit.log
timeout 600 python -m pytest tests/test_net_gpu.py -q 2>&1 | tail -15 > gpurun_out/exp2_pytest.log
tail -3 gpurun_out/exp2_pytest.log
KG_TC_DEBUG=1 timeout 300 python tools/op_times.py > gpurun_out/exp2_optimes.log 2>&1
grep -E "tc_shift|^op  5[89]|^op  6|^op  7|total" gpurun_out/exp2_optimes.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp2_bench.json 2> gpurun_out/exp2_bench.err
tail -c 1500 gpurun_out/exp2_bench.json
