#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_decode_gpu.py -q 2>&1 | tail -8 > gpurun_out/exp5_pytest.log
tail -3 gpurun_out/exp5_pytest.log
timeout 600 python bench.py --workload decode --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/exp5_decode.json 2> gpurun_out/exp5_decode.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/exp5_decode.json'))
print(d['value'], d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['roofline']['stages'].items()})
PY
