set -x
PT="python -m pytest -x -q -m gpu --timeout 120 --timeout-method thread"
timeout 300 $PT tests/test_kernels_gpu.py -k "attention_cls" > gpurun_out/r2f_t1.log 2>&1; echo "t1 rc=$?"; tail -3 gpurun_out/r2f_t1.log
timeout 600 $PT tests/test_tower_gpu.py tests/test_plugin_gpu.py > gpurun_out/r2f_t2.log 2>&1; echo "t2 rc=$?"; tail -3 gpurun_out/r2f_t2.log
timeout 300 $PT tests/test_probe_multitask.py tests/test_probe_gpu.py > gpurun_out/r2f_t3.log 2>&1; echo "t3 rc=$?"; tail -3 gpurun_out/r2f_t3.log
timeout 120 python bench.py --model eva --no-cpu-baseline > gpurun_out/r2f_eva_tail.json 2> gpurun_out/r2f_eva_tail.err; echo rc=$?
VLMFW_CLS_TAIL=0 timeout 120 python bench.py --model eva --no-cpu-baseline > gpurun_out/r2f_eva_notail.json 2> gpurun_out/r2f_eva_notail.err; echo rc=$?
timeout 120 python bench.py --model eva --batch 256 --no-cpu-baseline > gpurun_out/r2f_eva_tail_b256.json 2> gpurun_out/r2f_eva_tail_b256.err; echo rc=$?
timeout 120 python tools/bench_probe_multitask.py > gpurun_out/r2f_mt_linear.json 2> gpurun_out/r2f_mt_linear.err; echo rc=$?
timeout 120 python tools/bench_probe_multitask.py --deeper > gpurun_out/r2f_mt_deeper.json 2> gpurun_out/r2f_mt_deeper.err; echo rc=$?
cat gpurun_out/r2f_mt_*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2f_eva*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['step_frac_burst'], d['gpu_launches_per_step'])
    except Exception as e: print(f, 'ERR', e)
PY
