# 2-GPU check of one build (one box): sharded inference / DP probe training equal the single-process results, headline bench
# and probe-training bench under torchrun.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
timeout 240 $TR --master-port 29521 tools/dp_check.py > gpurun_out/r2y_dp_check_n2.json 2> gpurun_out/r2y_dp_check_n2.err; echo "dp_check rc=$?"; tail -1 gpurun_out/r2y_dp_check_n2.json | cut -c1-600
timeout 300 $TR --master-port 29522 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2y_bench_siglip_n2.json 2> gpurun_out/r2y_bench_siglip_n2.err; echo "bench n2 rc=$?"
timeout 200 $TR --master-port 29523 bench.py --workload probe_train --gpus 2 --comm p2p > gpurun_out/r2y_pt_n2_p2p.json 2> gpurun_out/r2y_pt_n2_p2p.err; echo "pt n2 rc=$?"
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2y_*n2*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('e2e',{}).get('value'), d.get('ms_per_step'), d.get('clocks',{}))
    except Exception as e: print(f, 'ERR', e)
PY
