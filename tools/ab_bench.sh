#!/bin/bash
# quick A/B on the GPU box: net parity tests, per-op times, pipeline bench summary
timeout 600 python -m pytest tests/test_net_gpu.py -q 2>&1 | tail -2
timeout 300 python tools/op_times.py 2>&1 | grep -E "^op +(0|2|6|7|10|17|20|30|33|47|48|50|51|57):|total" | tr '\n' ' '; echo
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('pipeline', d['value'], d['ms_per_step'], d['e2e']['value'])
print({k:v['ms_per_step'] for k,v in d['roofline']['stages'].items()})"
