"""Per-kernel timing on one B200 (CUDA events, L2-exceeding operands): our tcgen05 GEMM vs cuBLAS (torch.matmul) on the
tower's GEMM shapes, plus LayerNorm / attention. Usage: python tools/bench_kernels.py [--m 65536]"""
import argparse
import json
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vlmfw_b200 import ops, specs  # noqa: E402

DT = torch.float16


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for s, e in evs:
        s.record()
        fn()
        e.record()
    torch.cuda.synchronize()
    ts = sorted(s.elapsed_time(e) for s, e in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=65536)
    ap.add_argument("--block-n", type=int, default=0)
    ap.add_argument("--only", default="")
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--resid32", action="store_true", help="fp32 residual epilogue (bf16 towers / VLMFW_RESIDUAL=32)")
    ap.add_argument("--dense-ld", action="store_true", help="row pitch = row length (the tower pads 16-bit rows to 128 B)")
    args = ap.parse_args()
    m = args.m
    dev = "cuda"
    rows = []
    shapes = [("siglip.qkv", 3456, 1152, "bias"), ("siglip.o", 1152, 1152, "resid"), ("siglip.fc1", 4304, 1152, "gelu"),
              ("siglip.fc2", 1152, 4304, "resid"), ("clip.fc1", 4096, 1024, "gelu"), ("clip.fc2", 1024, 4096, "resid"),
              ("eva.fc1", 6144, 1408, "gelu"), ("eva.fc2", 1408, 6144, "resid"), ("eva.qkv", 4224, 1408, "bias")]
    for name, n, k, kind in shapes:
        if args.only and args.only not in name:
            continue
        pad = (lambda r, c: torch.empty(r, c, dtype=DT, device=dev)) if args.dense_ld else \
            (lambda r, c: ops.padded_rows(r, c, DT, dev))
        a, w = pad(m, k), pad(n, k)
        a.copy_(torch.randn(m, k, device=dev) * 0.5)
        w.copy_(torch.randn(n, k, device=dev) / math.sqrt(k))
        bias = torch.randn(n, device=dev)
        flops = 2.0 * m * n * k
        if kind == "resid" and not args.resid32:
            # what the fp16 tower runs: the 16-bit residual stream updated in place (+ row statistics)
            h16 = (torch.randn(m, n, device=dev)).to(DT)
            fn = lambda: ops.gemm_tn_resid16(a, w, bias, h16, block_n=args.block_n)  # noqa: E731
        elif kind == "resid":
            h = torch.randn(m, n, device=dev)
            fn = lambda: ops.gemm_tn(a, w, bias, resid=h, out=h, block_n=args.block_n)  # noqa: E731
        else:
            out = pad(m, n)
            # each tower's own activation: SigLIP gelu_pytorch_tanh, CLIP quick_gelu, EVA gelu (erf)
            gelu = specs.ACT_QUICK_GELU if name.startswith("clip") else specs.ACT_GELU_ERF if name.startswith("eva") else specs.ACT_GELU_TANH
            act = gelu if kind == "gelu" else specs.ACT_NONE
            fn = lambda: ops.gemm_tn(a, w, bias, act=act, out=out, block_n=args.block_n)  # noqa: E731
        med, best = timeit(fn)
        wt = w.t()
        cb_out = torch.empty(m, n, dtype=DT, device=dev)
        cmed, cbest = (float("nan"), float("nan")) if args.no_cublas else timeit(lambda: torch.matmul(a, wt, out=cb_out))
        rows.append({"kernel": name, "m": m, "n": n, "k": k, "epi": kind, "ms": round(med, 4), "tflops": round(flops / med / 1e9, 1),
                     "best_tflops": round(flops / best / 1e9, 1), "cublas_ms": round(cmed, 4),
                     "cublas_tflops": round(flops / cmed / 1e9, 1)})
        print(json.dumps(rows[-1]), flush=True)
        del a, w
    if not args.only or "ln" in args.only:
        d = 1152
        x = torch.randn(m, d, device=dev)
        g, b = torch.ones(d, device=dev), torch.zeros(d, device=dev)
        out = torch.empty(m, d, dtype=DT, device=dev)
        med, best = timeit(lambda: ops.layernorm(x, g, b, 1e-6, out=out))
        byts = m * d * 6
        print(json.dumps({"kernel": "layernorm", "rows": m, "dim": d, "ms": round(med, 4), "gbs": round(byts / med / 1e6, 1)}), flush=True)
    if not args.only or args.only.endswith("att"):
        for name, s, hd, dh in [("siglip.att", 256, 16, 72), ("clip.att", 577, 16, 64), ("eva.att", 257, 16, 88)]:
            if args.only and args.only not in name and args.only != "att":
                continue
            bsz = max(1, m // s)
            qkv = (torch.randn(bsz * s, 3 * hd * dh, device=dev)).to(DT)
            med, best = timeit(lambda: ops.attention(qkv, bsz, s, hd, dh))
            flops = 4.0 * bsz * hd * s * s * dh
            print(json.dumps({"kernel": name, "batch": bsz, "ms": round(med, 4), "tflops": round(flops / med / 1e9, 1)}), flush=True)


if __name__ == "__main__":
    main()
