"""Generate tests/golden/*.npz by running THE REFERENCE ITSELF (imported from /root/reference, unmodified) on seeded
synthetic inputs. Only runs in the build container (needs /root/reference); the .npz files it writes are committed.

    python tests/golden/make_golden.py            # all fixtures
    python tests/golden/make_golden.py tiny       # only the tiny towers

What is the reference here: its adapter classes ``PaliGemmaBackbone`` / ``LLaVABackbone`` / ``BLIP2Backbone``
(models/paligemma.py:88-112, models/llava.py:86-119, models/blip2.py:74-95) wrapped around the live Hugging Face vision
modules, and its ``LinearProbe`` / ``MultiTaskProbe`` (probing/models/*.py) on top, called with lists of PIL images
exactly like ``probing/test/singletask_tester.py:95-98`` does. Harness-side shims (reference files untouched,
SURVEY.md §0 fact 4): ``huggingface_hub.login`` is stubbed (import-time network call at models/paligemma.py:17-19),
and the adapters are constructed directly because ``from_pretrained`` needs a network.
Weights / images / head parameters come from the seeded generators in vlmfw_b200.weights and oracle.tower_oracle, so
the fixtures only store OUTPUTS (features, logits, labels) plus the seeds.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import huggingface_hub  # noqa: E402

huggingface_hub.login = lambda *a, **k: None  # shim 1: no network
for mod in ("matplotlib", "matplotlib.pyplot"):  # shim 2: not installed, imported by reference trainers only
    sys.modules.setdefault(mod, types.ModuleType(mod))

from PIL import Image  # noqa: E402

from models.blip2 import BLIP2Backbone  # noqa: E402  (reference)
from models.llava import LLaVABackbone  # noqa: E402  (reference)
from models.paligemma import PaliGemmaBackbone  # noqa: E402  (reference)
from probing.models.linear_probe import LinearProbe  # noqa: E402  (reference)
from probing.models.multitask_probe import MultiTaskProbe  # noqa: E402  (reference)

from oracle import hf_ref, tower_oracle  # noqa: E402
from vlmfw_b200 import specs, weights  # noqa: E402

ADAPTERS = {specs.ARCH_SIGLIP: PaliGemmaBackbone, specs.ARCH_CLIP: LLaVABackbone, specs.ARCH_EVA: BLIP2Backbone}
MIVIA_TASKS = {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}  # SURVEY.md §8 a12 (config 3)

# fixture name -> (spec, n_images, weight seed, image seed, smooth)
CASES = {
    "siglip_tiny": (specs.SIGLIP_TINY, 6, 11, 21, False),
    "clip_tiny": (specs.CLIP_TINY, 6, 12, 22, True),
    "eva_tiny": (specs.EVA_TINY, 6, 13, 23, False),
    "siglip_so400m": (specs.SIGLIP_SO400M, 4, 1, 2, True),
    "clip_vit_l_336": (specs.CLIP_VIT_L_336, 2, 1, 2, True),
    "eva_vit_g": (specs.EVA_VIT_G, 2, 1, 2, True),
}


def load_head(seq, params):
    seq.load_state_dict({k: v for k, v in params.items()}, strict=True)


def run_case(name):
    spec, n, wseed, iseed, smooth = CASES[name]
    torch.manual_seed(0)
    sd = weights.seeded_hf_state_dict(spec, wseed)
    module = hf_ref.build_hf_vision_model(spec, sd)
    proc = hf_ref.ProcessorShim(hf_ref.build_hf_processor(spec))
    backbone = ADAPTERS[spec.arch](proc, module, spec.hidden, torch.device("cpu"))
    imgs_u8 = weights.seeded_images_u8(n, spec.image_size, iseed, smooth=smooth)
    pil = [Image.fromarray(a.numpy()) for a in imgs_u8]
    out = {"weight_seed": wseed, "image_seed": iseed, "smooth": int(smooth), "n": n}

    lin = LinearProbe(backbone, n_out_classes=9, freeze_backbone=True, dropout_p=0.3).eval()
    load_head(lin.classifier, tower_oracle.seeded_head_params(spec.hidden, 9, 0, seed=100))
    deep = LinearProbe(backbone, n_out_classes=7, freeze_backbone=True, deeper_head=True, hidden_dim=512).eval()
    load_head(deep.classifier, tower_oracle.seeded_head_params(spec.hidden, 7, 512, seed=101))
    multi = MultiTaskProbe(backbone, tasks=MIVIA_TASKS, freeze_backbone=True).eval()
    for i, (t, c) in enumerate(MIVIA_TASKS.items()):
        load_head(multi.heads[t], tower_oracle.seeded_head_params(spec.hidden, c, 0, seed=200 + i))
    with torch.no_grad():
        out["features"] = lin.extract_features(pil).numpy()
        out["linear_logits"] = lin(pil).numpy()
        out["linear_pred"] = lin.predict(pil).numpy()
        out["deeper_logits"] = deep(pil).numpy()
        out["deeper_pred"] = deep.predict(pil).numpy()
        mo = multi(pil)["logits"]
        mp = multi.predict(pil)
        for t in MIVIA_TASKS:
            out[f"multi_logits_{t}"] = mo[t].numpy()
            out[f"multi_pred_{t}"] = mp[t].numpy()
        if spec.arch == specs.ARCH_CLIP:
            out["features_cls"] = backbone(pil, strategy="cls").numpy()
    path = os.path.join(HERE, f"{name}.npz")
    np.savez_compressed(path, **out)
    print(f"{name}: features {out['features'].shape} std {out['features'].std():.4f} -> {path} "
          f"({os.path.getsize(path)} bytes)")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    for nm in CASES:
        if which == "all" or which in nm:
            run_case(nm)
