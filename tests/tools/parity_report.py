"""Numerics margin of the CUDA path against the committed reference golden vectors (tests/golden/*.npz): worst pooled-
feature cosine and probe-logit max-abs error per tower, next to the north-star tolerances (>= 0.999, <= 1e-2).
Usage (GPU box): python tools/parity_report.py [--dtype f16|bf16]"""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from vlmfw_b200 import probes, specs, weights  # noqa: E402
from vlmfw_b200.vision_backbone import backbone_from_spec  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--dtype", default="f16")
    args = ap.parse_args()
    from oracle import tower_oracle as O

    for name in ["siglip_tiny", "clip_tiny", "eva_tiny", "siglip_so400m", "clip_vit_l_336", "eva_vit_g"]:
        g = dict(np.load(os.path.join(ROOT, "tests", "golden", f"{name}.npz")))
        spec = specs.SPECS[name]
        bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, int(g["weight_seed"])), device="cuda:0",
                                compute_dtype=args.dtype)
        imgs = weights.seeded_images_u8(int(g["n"]), spec.image_size, int(g["image_seed"]), smooth=bool(g["smooth"]))
        probe = probes.LinearProbe(bb, 9).to("cuda:0").eval()
        probe.classifier.load_state_dict(O.seeded_head_params(spec.hidden, 9, 0, seed=100))
        feats = probe.extract_features(imgs).float().cpu()
        ref = torch.from_numpy(g["features"])
        logits = probe(imgs).float().cpu()
        refl = torch.from_numpy(g["linear_logits"])
        print(json.dumps({
            "tower": name, "dtype": args.dtype,
            "min_cosine": round(torch.nn.functional.cosine_similarity(feats, ref).min().item(), 7),
            "feature_rel_l2": round(((feats - ref).norm() / ref.norm()).item(), 6),
            "logit_max_abs_err": round((logits - refl).abs().max().item(), 6),
            "labels_equal": bool((logits.argmax(1) == refl.argmax(1)).all())}), flush=True)
        del bb, probe
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
