#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_profile.sh <tag>   -> gpurun_out/<tag>_*
# pytest -m gpu, the default bench line (with CPU and cuDNN baselines), the decode / cfg4 / free-running lines, the ncu launch list of
# one step and `--set full` captures of (a) the tensor-core launches and (b) the decode kernels of that step (captured with --serial,
# one batch in flight, so that the launches of a step are contiguous: forward_dec, decode, forward_seg).
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
python -m pytest tests -m gpu -q 2>&1 | tail -30 > gpurun_out/${tag}_pytest.log
tail -2 gpurun_out/${tag}_pytest.log
python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 300 gpurun_out/${tag}_bench.json; echo
python bench.py --workload decode --steps 20 --warmup 3 > gpurun_out/${tag}_bench_decode.json 2>> gpurun_out/${tag}_bench.err
python bench.py --config cfg4 --steps 8 --warmup 3 --no-gpu-baseline > gpurun_out/${tag}_bench_cfg4.json 2>> gpurun_out/${tag}_bench.err
python bench.py --free-running --steps 8 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_bench_free_running.json 2>> gpurun_out/${tag}_bench.err
# launch list: every launch of the library's kernels in a 3 + 1 step run; tools/last_step.py keeps the timed step (a step starts with
# the 3x3 stem kernel) -> gpurun_out/<tag>_launches.csv
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'kg::|tc_|vote|blur|exact_peaks|group_kernel|nms_kernel|bilinear|maxpool|preprocess|copy_rects|fill_rects' --csv \
    --log-file gpurun_out/${tag}_launches_all.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
python tools/last_step.py gpurun_out/${tag}_launches_all.csv gpurun_out/${tag}_launches.csv
if [ "${FULL:-1}" = "1" ]; then
  TC=$(python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/${tag}_launches.csv')))
hi=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
k=rows[hi].index('Kernel Name')
print(sum(1 for r in rows[hi+1:] if len(r)>k and ('tc_conv' in r[k] or 'tc_shift' in r[k])))
PY
)
  echo "tensor-core launches/step=$TC"
  timeout 900 ncu --set full --clock-control none -k regex:'tc_conv|tc_shift' -s $((3*TC)) -c $TC -f -o gpurun_out/${tag}_tc \
      python bench.py --serial --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  ncu -i gpurun_out/${tag}_tc.ncu-rep --page raw --csv > gpurun_out/${tag}_tc_raw.csv 2>/dev/null
  timeout 600 ncu --set full --clock-control none -k regex:'vote_kernel|blur32|exact_peaks' -s 27 -c 9 -f -o gpurun_out/${tag}_decode \
      python bench.py --serial --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline > gpurun_out/${tag}_ncu_decode.log 2>&1
  ncu -i gpurun_out/${tag}_decode.ncu-rep --page raw --csv > gpurun_out/${tag}_decode_raw.csv 2>/dev/null
  for f in gpurun_out/${tag}_tc.ncu-rep gpurun_out/${tag}_decode.ncu-rep; do
    sz=$(stat -c %s $f 2>/dev/null || echo 0); if [ "$sz" -gt 30000000 ]; then rm -f $f; fi
  done
  ls -la gpurun_out/ | grep ${tag}
fi
