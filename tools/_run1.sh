set -x
TAG=r2z
timeout 1500 python -m pytest tests -m gpu -x -q --timeout 600 --timeout-method thread > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_siglip_f16.json 2> gpurun_out/${TAG}_bench_siglip_f16.err; echo "bench rc=$?"
timeout 400 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/${TAG}_bench_siglip_bf16.json 2> gpurun_out/${TAG}_bench_siglip_bf16.err; echo "bench bf16 rc=$?"
for m in clip eva; do timeout 400 python bench.py --model $m --no-cpu-baseline > gpurun_out/${TAG}_bench_$m.json 2> gpurun_out/${TAG}_bench_$m.err; echo "bench $m rc=$?"; done
timeout 300 python bench.py --workload probe_train > gpurun_out/${TAG}_pt_n1.json 2> gpurun_out/${TAG}_pt_n1.err; echo "pt rc=$?"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2z_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks',{}).get('sm_mhz'), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('step_frac_burst'))
    except Exception as e: print(f, 'ERR', e)
PY
bash tools/run_ncu_towers.sh r2z
