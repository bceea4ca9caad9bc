"""ORACLE (test infrastructure): the live Hugging Face modules the reference adapters wrap, built offline.

The reference obtains these through ``from_pretrained`` (models/paligemma.py:46, models/llava.py:32, models/blip2.py:32)
which needs a network; here they are constructed from configs with the real architecture constants and loaded with
seeded weights. ``transformers`` is a third-party dependency present in this image (5.5.0), not part of
/root/reference, so this module also works on the GPU box. It is used (a) to validate ``tower_oracle`` and (b) as the
CPU baseline / ``--impl reference`` arm of bench.py, because it IS the code the reference executes for this path.
"""
from __future__ import annotations

from typing import Mapping

import torch

ARCH_SIGLIP, ARCH_CLIP, ARCH_EVA = 0, 1, 2
_ACT = {0: "gelu_pytorch_tanh", 1: "quick_gelu", 2: "gelu"}


def build_hf_vision_model(spec, state_dict: Mapping[str, torch.Tensor]):
    """HF vision module for ``spec`` (SiglipVisionModel / CLIPVisionModel / Blip2VisionModel), eval mode, fp32."""
    import transformers as tf

    common = dict(hidden_size=spec.hidden, intermediate_size=spec.mlp, num_hidden_layers=spec.layers,
                  num_attention_heads=spec.heads, image_size=spec.image_size, patch_size=spec.patch_size,
                  hidden_act=_ACT[spec.act], layer_norm_eps=spec.ln_eps)
    if spec.arch == ARCH_SIGLIP:
        # PaliGemma's tower has no pooling head (hf:models/paligemma/configuration_paligemma.py:67-76)
        model = tf.SiglipVisionModel(tf.SiglipVisionConfig(vision_use_head=False, **common))
    elif spec.arch == ARCH_CLIP:
        model = tf.CLIPVisionModel(tf.CLIPVisionConfig(**common))
    elif spec.arch == ARCH_EVA:
        model = tf.Blip2VisionModel(tf.Blip2VisionConfig(qkv_bias=True, **common))
    else:
        raise ValueError(f"unknown arch {spec.arch}")
    missing, unexpected = model.load_state_dict(dict(state_dict), strict=False)
    missing = [m for m in missing if "position_ids" not in m]
    if missing or unexpected:
        raise RuntimeError(f"state dict mismatch for {spec.name}: missing={missing[:5]} unexpected={unexpected[:5]}")
    return model.eval()


@torch.no_grad()
def hf_features(spec, model, pixel_values: torch.Tensor, strategy: str = "mean", chunk: int = 8) -> torch.Tensor:
    """The adapter pooling on top of the HF module — the three ``forward`` bodies of the reference adapters."""
    outs = []
    for i in range(0, pixel_values.shape[0], chunk):
        pv = pixel_values[i:i + chunk]
        if spec.arch == ARCH_SIGLIP:  # models/paligemma.py:110-112
            outs.append(model(pv).last_hidden_state.mean(dim=1))
        elif spec.arch == ARCH_CLIP:  # models/llava.py:111-117
            h = model(pv).last_hidden_state
            outs.append(h[:, 0] if strategy == "cls" else h.mean(dim=1))
        else:  # models/blip2.py:93-95
            outs.append(model(pixel_values=pv).pooler_output)
    return torch.cat(outs, dim=0)


def build_hf_processor(spec):
    """Offline-constructed HF image processor with the checkpoint's preprocessing constants."""
    import transformers as tf

    size = spec.image_size
    if spec.arch == ARCH_SIGLIP:
        return tf.SiglipImageProcessor(size={"height": size, "width": size}, image_mean=list(spec.pix_mean),
                                       image_std=list(spec.pix_std), resample=3)
    if spec.arch == ARCH_CLIP:
        return tf.CLIPImageProcessor(size={"shortest_edge": size}, crop_size={"height": size, "width": size},
                                     image_mean=list(spec.pix_mean), image_std=list(spec.pix_std))
    return tf.BlipImageProcessor(size={"height": size, "width": size}, image_mean=list(spec.pix_mean),
                                 image_std=list(spec.pix_std))


class ProcessorShim:
    """What the adapters call: ``processor(images=..., text=..., return_tensors="pt")`` -> mapping with
    ``pixel_values`` and a ``.to(device)`` (reference models/paligemma.py:108). ``text`` is ignored — the adapters
    only read ``inputs['pixel_values']``."""

    def __init__(self, image_processor):
        self.image_processor = image_processor

    def __call__(self, images=None, text=None, return_tensors="pt", **kw):
        return self.image_processor(images=images, return_tensors=return_tensors)
