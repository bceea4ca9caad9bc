"""CPU experiment: which 16-bit roundings of the tower dominate the probe-logit error? Emulates the CUDA path's rounding
points (weights, LN output, qkv, softmax P, attention output, MLP hidden) in the fp32 oracle, one at a time and all
together, for bf16 and fp16. Usage: python tools/error_budget.py [siglip_so400m] [n_images]"""
import math
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import tower_oracle as O  # noqa: E402
from vlmfw_b200 import specs, weights  # noqa: E402


def run(spec, w, pv, q, dt):
    r = (lambda t: t.to(dt).float())
    rq = {k: (r if k in q else (lambda t: t)) for k in ("w", "xn", "qkv", "p", "ctx", "mid", "patch", "h")}
    W = {k: (rq["w"](v) if v.dim() == 2 and k != "pos" else v) for k, v in w.items()}
    b = pv.shape[0]
    p, g, d = spec.patch_size, spec.image_size // spec.patch_size, spec.hidden
    patches = rq["patch"](pv.reshape(b, 3, g, p, g, p).permute(0, 2, 4, 1, 3, 5).reshape(b, g * g, 3 * p * p))
    h = patches @ W["patch.w"].t() + (W["patch.b"] if spec.conv_bias else 0)
    if spec.has_cls:
        h = torch.cat([W["cls"].view(1, 1, d).expand(b, 1, d), h], 1)
    h = rq["h"](h + W["pos"])
    if spec.has_pre_ln:
        h = O._ln(h, W["pre_ln.w"], W["pre_ln.b"], spec.ln_eps)
    nh, dh = spec.heads, spec.head_dim
    for i in range(spec.layers):
        L = f"L{i}."
        x = rq["xn"](O._ln(h, W[L + "ln1.w"], W[L + "ln1.b"], spec.ln_eps))
        qkv = rq["qkv"](x @ W[L + "qkv.w"].t() + W[L + "qkv.b"])
        qq, kk, vv = (t.reshape(b, -1, nh, dh).transpose(1, 2) for t in qkv.split(d, -1))
        s = (qq @ kk.transpose(-1, -2)) / math.sqrt(dh)
        e = torch.exp(s - s.amax(-1, keepdim=True))
        ctx = (rq["p"](e) @ vv) / e.sum(-1, keepdim=True)
        ctx = rq["ctx"](ctx.transpose(1, 2).reshape(b, -1, d))
        h = rq["h"](h + ctx @ W[L + "o.w"].t() + W[L + "o.b"])  # "h": a 16-bit residual stream (not what the tower does)
        x = rq["xn"](O._ln(h, W[L + "ln2.w"], W[L + "ln2.b"], spec.ln_eps))
        m = rq["mid"](O._act(x @ W[L + "fc1.w"].t() + W[L + "fc1.b"], spec.act))
        h = rq["h"](h + m @ W[L + "fc2.w"].t() + W[L + "fc2.b"])
    return O.pool(spec, W, h, spec.pool)


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "siglip_so400m"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    spec = specs.SPECS[name]
    w = {k: v.float() for k, v in weights.hf_to_canonical(spec, weights.seeded_hf_state_dict(spec, 1)).items()}
    pv = O.preprocess_u8(weights.seeded_images_u8(n, spec.image_size, 2, smooth=True), spec.pix_mean, spec.pix_std)
    heads = [weights.seeded_head_params(spec.hidden, c, 0, seed=200 + i) for i, c in enumerate((2, 12, 12, 2, 2))]
    with torch.no_grad():
        ref = run(spec, w, pv, set(), torch.float32)
        ref_l = torch.cat([O.head_forward(ref, hp) for hp in heads], 1)
        for dt in (torch.bfloat16, torch.float16):
            only = os.environ.get("BUDGET_ONLY")
            cases = [only.split("+")] if only else None
            for q in cases or (["w"], ["xn"], ["qkv"], ["p"], ["ctx"], ["mid"], ["patch"], ["w", "xn", "qkv", "p", "ctx", "mid", "patch"],
                      ["w", "xn", "qkv", "ctx", "mid", "patch"]):
                f = run(spec, w, pv, set(q), dt)
                lg = torch.cat([O.head_forward(f, hp) for hp in heads], 1)
                cos = torch.nn.functional.cosine_similarity(f, ref).min().item()
                print(f"{str(dt):16s} {'+'.join(q):28s} logit max err {float((lg - ref_l).abs().max()):.5f} rms {float((lg - ref_l).pow(2).mean().sqrt()):.5f} cos {cos:.6f}", flush=True)


if __name__ == "__main__":
    main()
