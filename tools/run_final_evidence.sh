# Evidence run of one build on one B200 (profiles/r2_y_*): full GPU test suite (-s keeps the live-oracle parity lines), smoke, the bench
# lines of the three towers (+ bf16), then the ncu launch list of one eager step and --set full of one encoder layer per tower.
# Reports are summarised ON THE BOX (tools/ncu_summary.py + the raw page as csv): gpurun_out/ brings back at most 64 MiB.
set -x
TAG=r2y
timeout 1500 python -m pytest tests -m gpu -x -q -s --timeout 600 --timeout-method thread > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "gpu tests rc=$?"; tail -3 gpurun_out/${TAG}_gpu_tests.log; grep "parity-live" gpurun_out/${TAG}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/${TAG}_smoke.log
timeout 400 python bench.py > gpurun_out/${TAG}_bench_siglip_f16.json 2> gpurun_out/${TAG}_bench_siglip_f16.err; echo "bench rc=$?"
timeout 400 python bench.py --dtype bf16 --no-cpu-baseline > gpurun_out/${TAG}_bench_siglip_bf16.json 2> gpurun_out/${TAG}_bench_siglip_bf16.err; echo "bench bf16 rc=$?"
for m in clip eva; do timeout 400 python bench.py --model $m --no-cpu-baseline > gpurun_out/${TAG}_bench_$m.json 2> gpurun_out/${TAG}_bench_$m.err; echo "bench $m rc=$?"; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks',{}).get('sm_mhz'), d.get('roofline',{}).get('frac'), d.get('roofline',{}).get('step_frac_burst'))
    except Exception as e: print(f, 'ERR', e)
PY
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
for m in siglip clip eva; do
VLMFW_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$m.csv $B --model $m > gpurun_out/${TAG}_ncu_launches_$m.log 2>&1; echo "launches $m rc=$?"
done
full() { name=$1; model=$2; shift 2
  VLMFW_PROFILE_RANGE=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off "$@" -o /tmp/$name -f $B --model $model > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "full $name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep --csv gpurun_out/${TAG}_ncu_$name.csv > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${name}_raw.csv 2>/dev/null
}
full layer_siglip siglip --launch-skip 52 --launch-count 5
full layer_clip clip --launch-skip 55 --launch-count 5
full layer_eva eva --launch-skip 54 --launch-count 5
