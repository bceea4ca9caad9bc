"""One launch sequence of the hot kernels at the headline shapes (SigLIP-So400m, batch 256 -> M = 65536), for
`ncu --set full -k regex:'gemm_bf16_tn|attention_tc|row_stats_cast|pool_ln_mean|im2col'`. Each kernel runs twice
(first = warm-up: tensor maps, func attributes); profile the second with --launch-skip or read both."""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import ops, specs  # noqa: E402

DT = torch.float16
spec = specs.SIGLIP_SO400M
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
M, D, I = B * spec.tokens, spec.hidden, spec.mlp
dev = "cuda"
g = torch.Generator(device=dev).manual_seed(0)
rn = lambda *s, sc=1.0: torch.randn(*s, device=dev, generator=g) * sc  # noqa: E731

h = rn(M, D)
gamma, beta = 1 + rn(D, sc=0.1), rn(D, sc=0.1)
wqkv, bqkv = rn(3 * D, D, sc=1 / math.sqrt(D)).to(DT), rn(3 * D, sc=0.1)
wo, bo = rn(D, D, sc=1 / math.sqrt(D)).to(DT), rn(D, sc=0.1)
w1, b1 = rn(I, D, sc=1 / math.sqrt(D)).to(DT), rn(I, sc=0.1)
# fc2's operands with the tower's row pitch (I rounded up to whole 128-byte lines)
w2 = ops.padded_rows(D, I, DT, dev)
w2.copy_(rn(D, I, sc=1 / math.sqrt(I)))
b2 = rn(D, sc=0.1)
mid = ops.padded_rows(M, I, DT, dev)
pix = torch.randint(0, 256, (B, 224, 224, 3), device=dev, dtype=torch.uint8, generator=g)

w1f, b1c = ops.ln_fold_weight(rn(I, D, sc=1 / math.sqrt(D)), gamma, beta, b1, dtype=DT)
wqf, bqc = ops.ln_fold_weight(rn(3 * D, D, sc=1 / math.sqrt(D)), gamma, beta, bqkv, dtype=DT)
parts = ops.gemm_stats_parts(D)

for it in range(2):  # the tower's launch sequence for one layer (tower.cu), at the headline shapes
    a = ops.im2col(pix, 224, 14, 640, spec.pix_mean, spec.pix_std, dtype=DT)
    xn, stats = ops.row_stats_cast(h, dtype=DT, parts=parts)
    qkv = ops.gemm_tn_ln(xn, wqf, bqc, stats, D, 1e-6)
    ctx = ops.attention(qkv, B, spec.tokens, spec.heads, spec.head_dim)
    stats = ops.gemm_tn_resid16(ctx, wo, bo, xn)  # the fp16 tower's residual stream: xn updated in place
    ops.gemm_tn_ln(xn, w1f, b1c, stats, D, 1e-6, act=spec.act, out=mid)
    stats = ops.gemm_tn_resid16(mid, w2, b2, xn)
    pooled = ops.pool(xn.float().view(B, spec.tokens, D), specs.POOL_LN_MEAN, gamma, beta, 1e-6)
    if os.environ.get("VLMFW_PROFILE_CUBLAS") == "1":  # the library GEMM on the fc1 shape, for a side-by-side ncu read
        torch.matmul(xn, w1.t())
    torch.cuda.synchronize()
print("ok", float(pooled.float().abs().mean()))
