#!/bin/bash
# usage (on the GPU box, via gpurun): tools/gpu_profile.sh <tag>   -> gpurun_out/<tag>_*
# pytest -m gpu, the default bench line, the ncu launch list of one step and one `--set full` capture of that step.
tag=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader > gpurun_out/${tag}_gpu.txt
python -m pytest tests -m gpu -q 2>&1 | tail -60 > gpurun_out/${tag}_pytest.log
python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
tail -c 600 gpurun_out/${tag}_bench.json
# launches per step are printed by bench (gpu_launches / steps); warm-up = 3 steps
L=$(python -c "import json;d=json.load(open('gpurun_out/${tag}_bench.json'));print(d['gpu_launches']//d['steps'])")
echo "launches/step=$L"
SKIP=${SKIP:-$((3*L))}
ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $L --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_launches.log 2>&1
if [ "${FULL:-1}" = "1" ]; then
  timeout 700 ncu --set full --clock-control none -s $SKIP -c $L -f -o gpurun_out/${tag}_step \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_ncu_full.log 2>&1
  ncu -i gpurun_out/${tag}_step.ncu-rep --page raw --csv > gpurun_out/${tag}_step_raw.csv 2>/dev/null
  sz=$(stat -c %s gpurun_out/${tag}_step.ncu-rep)
  if [ "$sz" -gt 40000000 ]; then rm -f gpurun_out/${tag}_step.ncu-rep; fi
  ls -la gpurun_out/
fi
