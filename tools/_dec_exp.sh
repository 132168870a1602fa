for occ in 4 5 6; do
KG_B32_OCC=$occ python bench.py --workload decode --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/dec_occ$occ.json 2>gpurun_out/dec_occ$occ.err
python -c "
import json; d=json.load(open('gpurun_out/dec_occ$occ.json')); print('occ$occ', d['ms_per_step'], {k: round(v['ms_per_step'],3) for k,v in d['roofline']['stages'].items()})" || tail -3 gpurun_out/dec_occ$occ.err
done
KG_B32_OCC=5 python -m pytest tests/test_decode_gpu.py -m gpu -q -x 2>&1 | tail -2
