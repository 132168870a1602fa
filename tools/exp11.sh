#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 600 python bench.py --workload decode --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('decode', d['value'], d['ms_per_step'], d['roofline']['kernel'], d['roofline']['achieved'], d['roofline']['frac'])
print({k:round(v['ms_per_step'],4) for k,v in d['roofline']['stages'].items()})"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pipeline', d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})"
