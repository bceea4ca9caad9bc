# 8-GPU evidence run (one box): N-rank equivalence check, probe-training exchange at 4 / 8 ranks (fused NVLink form and NCCL),
# CLIP / EVA inference at 8 ranks. Everything under its own timeout; outputs are small JSON lines in gpurun_out/.
set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
DP_CHECK_BATCH=16 timeout 240 $TR --nproc-per-node 8 --master-port 29521 tools/dp_check.py > gpurun_out/r2h_dp_check_n8.json 2> gpurun_out/r2h_dp_check_n8.err; echo "dp_check rc=$?"
cat gpurun_out/r2h_dp_check_n8.json
for n in 8 4; do for comm in p2p nccl; do
  timeout 200 $TR --nproc-per-node $n --master-port 2953$n bench.py --workload probe_train --gpus $n --comm $comm > gpurun_out/r2h_pt_n${n}_$comm.json 2> gpurun_out/r2h_pt_n${n}_$comm.err; echo "pt n$n $comm rc=$?"
done; done
for m in clip eva; do
  timeout 300 $TR --nproc-per-node 8 --master-port 29541 bench.py --model $m --gpus 8 > gpurun_out/r2h_bench_${m}_n8.json 2> gpurun_out/r2h_bench_${m}_n8.err; echo "bench $m n8 rc=$?"
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2h_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        if 'value' in d: print(f, d['value'], d['e2e']['value'], d['ms_per_step'], d['clocks'].get('sm_mhz_min_over_ranks'), d['config'].get('exchange_measured'))
    except Exception as e: print(f, 'ERR', e)
PY
