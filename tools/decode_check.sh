#!/bin/bash
# usage (GPU box): tools/decode_check.sh <tag> : decode parity tests, decode-only bench line, ncu --set full of the vote / blur+peak launches
tag=${1:-dec}
mkdir -p gpurun_out
python -m pytest tests/test_decode_gpu.py tests/test_parity_full_gpu.py -m gpu -q -k "not forward_dec and not free_running and not inference and not weight_update and not stale" 2>&1 | tail -15 > gpurun_out/${tag}_pytest.log
tail -8 gpurun_out/${tag}_pytest.log
python bench.py --workload decode --steps 20 --warmup 3 > gpurun_out/${tag}_bench_decode.json 2> gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open('gpurun_out/${tag}_bench_decode.json'))
print('decode', d['value'], d['ms_per_step'], json.dumps(d['roofline']['stages']))
PY
if [ "${FULL:-1}" = "1" ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'vote_kernel|blur32|exact_peaks|blur_peak' -s 27 -c 9 -f -o gpurun_out/${tag}_decode \
      python bench.py --workload decode --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
  ncu -i gpurun_out/${tag}_decode.ncu-rep --page raw --csv > gpurun_out/${tag}_decode_raw.csv 2>/dev/null
fi
