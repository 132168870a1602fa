#!/bin/bash
# experiment: tap groups / accumulator chains on the narrow-N head convs
mkdir -p gpurun_out
python -m pytest tests/test_net_gpu.py -q -x 2>&1 | tail -15 > gpurun_out/exp1_pytest.log
KG_SEG_1PASS_LEVELS=5 python -m pytest tests/test_net_gpu.py -q -k "forward_seg_matches_golden" 2>&1 | grep -E "AssertionError|passed|failed" > gpurun_out/exp1_seg5.log
KG_SEG_1PASS_LEVELS=2 python -m pytest tests/test_net_gpu.py -q -k "forward_seg_matches_golden" 2>&1 | grep -E "AssertionError|passed|failed" > gpurun_out/exp1_seg2.log
KG_TC_DEBUG=1 python tools/op_times.py > gpurun_out/exp1_default.log 2>&1
KG_TC_WG=0 python tools/op_times.py > gpurun_out/exp1_wg0.log 2>&1
KG_TC_KSWANT=4 python tools/op_times.py > gpurun_out/exp1_ks4.log 2>&1
KG_TC_KSWANT=8 python tools/op_times.py > gpurun_out/exp1_ks8.log 2>&1
KG_TC_MT=1 KG_TC_KSWANT=8 python tools/op_times.py > gpurun_out/exp1_mt1ks8.log 2>&1
KG_TC_MT=1 KG_TC_KSWANT=1 python tools/op_times.py > gpurun_out/exp1_mt1ks1.log 2>&1
KG_TC_ACC=1 python tools/op_times.py > gpurun_out/exp1_acc1.log 2>&1
tail -3 gpurun_out/exp1_*.log
