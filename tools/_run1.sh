set -x
for st in 0 601 1201 1801; do
echo "== stagger $st"; VLMFW_WB_STAGGER=$st timeout 120 python tools/bench_kernels.py --only att --no-cublas 2>&1 | grep kernel
done
python - <<'PY'
# which CTAs share an SM? launch a tiny probe: 296 CTAs x 384 threads with 100 KB smem, record smid
import torch
from torch.utils.cpp_extension import load_inline
PY
