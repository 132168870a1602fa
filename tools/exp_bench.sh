#!/bin/bash
# usage (GPU box): tools/exp_bench.sh <tag> "<ENV=.. ENV=..>" [bench args]   -> gpurun_out/<tag>.json + one summary line
tag=$1; envs=$2; shift 2
mkdir -p gpurun_out
env $envs python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline "$@" > gpurun_out/${tag}.json 2> gpurun_out/${tag}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/${tag}.json'))
    st=d['roofline']['stages']
    print('${tag}', '[${envs}]', d['value'], 'img/s', d['ms_per_step'], 'ms e2e', d['e2e']['value'], {k: round(v['ms_per_step'],2) for k,v in st.items()})
except Exception as e:
    print('${tag} FAILED', e); print(open('gpurun_out/${tag}.err').read()[-1500:])
PY
