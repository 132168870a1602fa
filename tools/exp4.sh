#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_net_gpu.py -q 2>&1 | tail -8 > gpurun_out/exp4_pytest.log
tail -3 gpurun_out/exp4_pytest.log
timeout 300 python tools/op_times.py > gpurun_out/exp4_otma1.log 2>&1
KG_TC_OTMA=0 timeout 300 python tools/op_times.py > gpurun_out/exp4_otma0.log 2>&1
paste <(grep "^op\|^total" gpurun_out/exp4_otma0.log) <(grep "^op\|^total" gpurun_out/exp4_otma1.log | awk '{print $3}')
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/exp4_bench.json 2> gpurun_out/exp4_bench.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/exp4_bench.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})
PY
