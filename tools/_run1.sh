set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_tower_gpu.py tests/test_plugin_gpu.py -x -q -m gpu > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2e_tests.log
timeout 300 python bench.py --model eva --no-cpu-baseline > gpurun_out/r2e_eva_tail.json 2> gpurun_out/r2e_eva_tail.err; echo rc=$?
VLMFW_CLS_TAIL=0 timeout 300 python bench.py --model eva --no-cpu-baseline > gpurun_out/r2e_eva_notail.json 2> gpurun_out/r2e_eva_notail.err; echo rc=$?
timeout 300 python bench.py --model eva --batch 256 --no-cpu-baseline > gpurun_out/r2e_eva_tail_b256.json 2> gpurun_out/r2e_eva_tail_b256.err; echo rc=$?
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2e_eva*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['step_frac_burst'], d['gpu_launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
