set -x
TAG=r2q
V=$PWD/tools/_variants
BASE=$V/lib_base_r2j.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" --timeout 120 --timeout-method thread > gpurun_out/${TAG}_t_gemm.log 2>&1; echo "gemm tests rc=$?"; tail -3 gpurun_out/${TAG}_t_gemm.log
for i in 1 2; do
echo "== new"; timeout 200 python tools/bench_kernels.py --only siglip --no-cublas 2>&1 | grep kernel | cut -c1-150
echo "== base"; VLMFW_B200_LIB=$BASE timeout 200 python tools/bench_kernels.py --only siglip --no-cublas 2>&1 | grep kernel | cut -c1-150
done
echo "== new eva/clip"; timeout 200 python tools/bench_kernels.py --only fc --no-cublas --m 32896 2>&1 | grep -E "clip|eva" | cut -c1-150
echo "== base eva/clip"; VLMFW_B200_LIB=$BASE timeout 200 python tools/bench_kernels.py --only fc --no-cublas --m 32896 2>&1 | grep -E "clip|eva" | cut -c1-150
for m in siglip; do
for v in new base new base; do lib=""; [ $v = base ] && lib=$BASE
VLMFW_B200_LIB=$lib timeout 200 python bench.py --model $m --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_${m}_$v.json 2> gpurun_out/${TAG}_bench_${m}_$v.err; echo "bench $m $v rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/${TAG}_bench_${m}_$v.json').read().strip().splitlines()[-1]); print('$m $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done; done
