set -x
V=$PWD/tools/_variants
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "gemm" --timeout 120 --timeout-method thread 2>&1 | tail -2
for i in 1 2; do
for v in main nbuf3; do lib=""; [ $v != main ] && lib=$V/lib_$v.so
echo "== $v"; VLMFW_B200_LIB=$lib timeout 200 python tools/bench_kernels.py --only siglip --no-cublas 2>&1 | grep kernel | cut -c1-120
done; done
for v in main nbuf3; do lib=""; [ $v != main ] && lib=$V/lib_$v.so
echo "== $v eva/clip"; VLMFW_B200_LIB=$lib timeout 200 python tools/bench_kernels.py --only fc --no-cublas --m 32896 2>&1 | grep -E "clip|eva" | cut -c1-120
done
for v in main nbuf3 main nbuf3; do lib=""; [ $v != main ] && lib=$V/lib_$v.so
VLMFW_B200_LIB=$lib timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2w_bench_$v.json 2> gpurun_out/r2w_bench_$v.err; echo "bench $v rc=$?"
python -c "import sys,json; d=json.loads(open('gpurun_out/r2w_bench_$v.json').read().strip().splitlines()[-1]); print('siglip $v', d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks']['sm_mhz'], d['roofline']['kernel_ms'], d['roofline']['frac'])"
done
