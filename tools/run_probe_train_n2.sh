set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_probe_gpu.py -x -q -m gpu > gpurun_out/r2_probe_tests.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_probe_tests.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_check.py > gpurun_out/r2_dp_check_n2.json 2> gpurun_out/r2_dp_check_n2.err; echo "dp_check rc=$?"
cat gpurun_out/r2_dp_check_n2.json; tail -5 gpurun_out/r2_dp_check_n2.err
timeout 200 python bench.py --workload probe_train --gpus 1 --comm p2p > gpurun_out/r2_pt_n1_p2p.json 2> gpurun_out/r2_pt_n1_p2p.err; echo "pt n1 rc=$?"
for comm in p2p nccl; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload probe_train --gpus 2 --comm $comm > gpurun_out/r2_pt_n2_$comm.json 2> gpurun_out/r2_pt_n2_$comm.err; echo "pt n2 $comm rc=$?"
tail -3 gpurun_out/r2_pt_n2_$comm.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2_pt_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['config'].get('exchange_measured'))
    except Exception as e: print(f, 'ERR', e)
PY
