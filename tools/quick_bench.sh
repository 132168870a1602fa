#!/bin/bash
# usage: tools/quick_bench.sh "ENV=1 ENV2=2" label   -> prints value, ms/step and the per-stage table
env $1 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import sys,json
d=json.loads(sys.stdin.read())
st=d['roofline']['stages']
print('$2', d['value'], 'img/s', d['ms_per_step'], 'ms | e2e', d['e2e']['value'], '| frac', d['roofline']['frac'], '|', ' '.join(f\"{k}={v['ms_per_step']:.2f}\" for k,v in st.items()))
"
