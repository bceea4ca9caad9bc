set -x
TAG=r2r
V=$PWD/tools/_variants
BASE=$V/lib_base_r2j.so
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "attention" --timeout 120 --timeout-method thread > gpurun_out/${TAG}_t_att.log 2>&1; echo "att tests rc=$?"; tail -3 gpurun_out/${TAG}_t_att.log
for i in 1 2; do
echo "== new att"; timeout 120 python tools/bench_kernels.py --only att --no-cublas 2>&1 | grep kernel
echo "== base att"; VLMFW_B200_LIB=$BASE timeout 120 python tools/bench_kernels.py --only att --no-cublas 2>&1 | grep kernel
done
for m in siglip; do echo "#### trace $m"; VLMFW_B200_LIB=$V/lib_j_trace.so timeout 120 python tools/att_trace.py $m 2>&1 | grep -v Warn; done
timeout 400 python -m pytest tests/test_tower_gpu.py tests/test_plugin_gpu.py -m gpu -x -q --timeout 180 --timeout-method thread > gpurun_out/${TAG}_t_tower.log 2>&1; echo "tower tests rc=$?"; tail -3 gpurun_out/${TAG}_t_tower.log
for m in clip eva; do
for v in new base; do lib=""; [ $v = base ] && lib=$BASE
VLMFW_B200_LIB=$lib timeout 200 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${m}_$v.json 2> gpurun_out/${TAG}_bench_${m}_$v.err; echo "bench $m $v rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_bench_${m}_$v.json').read().strip().splitlines()[-1]); print('$m $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done; done
