# ncu evidence of one build, one GPU: per tower (a) the launch list of ONE eager step with device times, (b) --set full of the
# five kernels of one encoder layer in the middle of that step, (c) --set full of the small kernels (pooling, patch gather,
# probe, class-token attention). Reports are summarised ON THE BOX (tools/ncu_summary.py + the raw page as csv) and deleted:
# gpurun_out/ brings back at most 64 MiB.
set -x
TAG=${1:-r2}
B="python bench.py --steps 1 --warmup 3 --no-graph --no-cpu-baseline"
for m in siglip clip eva; do
  timeout 200 $B --model $m > gpurun_out/${TAG}_plain_$m.json 2> gpurun_out/${TAG}_plain_$m.err; echo "plain $m rc=$?"
done
for m in siglip clip eva; do
  VLMFW_PROFILE_RANGE=1 timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_$m.csv $B --model $m > gpurun_out/${TAG}_ncu_launches_$m.log 2>&1; echo "launches $m rc=$?"
done
full() {  # name, model, extra ncu args...
  name=$1; model=$2; shift 2
  VLMFW_PROFILE_RANGE=1 timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off "$@" -o /tmp/$name -f $B --model $model > gpurun_out/${TAG}_ncu_$name.log 2>&1; echo "full $name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep --csv gpurun_out/${TAG}_ncu_$name.csv > /dev/null 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > gpurun_out/${TAG}_ncu_${name}_raw.csv 2>/dev/null
}
full layer_siglip siglip --launch-skip 52 --launch-count 5
full layer_clip clip --launch-skip 55 --launch-count 5
full layer_eva eva --launch-skip 54 --launch-count 5
full small_eva eva -k regex:'pool_|im2col|probe_forward|attention_cls|row_stats'
full small_siglip siglip -k regex:'pool_|im2col'
cp /tmp/layer_siglip.ncu-rep gpurun_out/${TAG}_layer_siglip.ncu-rep   # one report kept whole (source page of the hot kernels)
du -sh gpurun_out
