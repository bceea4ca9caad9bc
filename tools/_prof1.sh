set -x
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
VLMFW_PROFILE_RANGE=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --launch-skip 54 --launch-count 1 -o gpurun_out/r2y_oproj -f $B --model siglip > gpurun_out/r2y_oproj.log 2>&1; echo "rc=$?"
ls -la gpurun_out/r2y_oproj.ncu-rep
