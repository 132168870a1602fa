#!/bin/bash
timeout 600 python -m pytest tests/test_decode_gpu.py -q 2>&1 | tail -2
for v in 0 1; do
if [ $v = 1 ]; then export KG_BLUR_EXACT=1; fi
timeout 600 python bench.py --workload decode --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('decode exact=$v', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['roofline']['stages'].items()})"
done
