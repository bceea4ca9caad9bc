"""Tower specifications: the three vision towers the reference adapters wrap, plus scaled-down twins for tests.

Constants follow the Hugging Face configs the reference instantiates through ``from_pretrained``
(SURVEY.md §8 model table):

* ``siglip_so400m``  — PaliGemma vision tower, reference ``models/paligemma.py:70`` (output_dim 1152)
* ``clip_vit_l_336`` — LLaVA-1.5 vision tower, reference ``models/llava.py:68`` (output_dim 1024)
* ``eva_vit_g``      — BLIP-2 vision model,    reference ``models/blip2.py:56`` (output_dim 1408)
"""
from __future__ import annotations

from dataclasses import dataclass, replace
from typing import Dict, Tuple

ARCH_SIGLIP, ARCH_CLIP, ARCH_EVA = 0, 1, 2
ACT_GELU_TANH, ACT_QUICK_GELU, ACT_GELU_ERF, ACT_NONE = 0, 1, 2, 3
POOL_LN_MEAN, POOL_MEAN, POOL_CLS, POOL_LN_CLS_LN = 0, 1, 2, 3
PIX_U8_NHWC, PIX_F32_NCHW = 0, 1

_OPENAI_CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)
_OPENAI_CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass(frozen=True)
class TowerSpec:
    name: str
    arch: int
    image_size: int
    patch_size: int
    hidden: int
    heads: int
    mlp: int
    layers: int
    has_cls: bool
    has_pre_ln: bool
    conv_bias: bool
    act: int
    pool: int
    ln_eps: float
    pix_mean: Tuple[float, float, float]
    pix_std: Tuple[float, float, float]

    @property
    def patches(self) -> int:
        g = self.image_size // self.patch_size
        return g * g

    @property
    def tokens(self) -> int:
        return self.patches + (1 if self.has_cls else 0)

    @property
    def head_dim(self) -> int:
        return self.hidden // self.heads

    @property
    def patch_k(self) -> int:
        return 3 * self.patch_size * self.patch_size

    def flops_per_image(self) -> float:
        """Algorithmic FLOPs of one image: 2·M·N·K for every contraction, elementwise ignored."""
        s, p, d, i, l = self.tokens, self.patches, self.hidden, self.mlp, self.layers
        return 2.0 * p * self.patch_k * d + l * (2.0 * s * d * 3 * d + 2.0 * s * d * d + 4.0 * s * s * d + 4.0 * s * d * i)

    def with_pool(self, pool: int) -> "TowerSpec":
        return replace(self, pool=pool)


SIGLIP_SO400M = TowerSpec(
    name="siglip_so400m", arch=ARCH_SIGLIP, image_size=224, patch_size=14, hidden=1152, heads=16, mlp=4304, layers=27,
    has_cls=False, has_pre_ln=False, conv_bias=True, act=ACT_GELU_TANH, pool=POOL_LN_MEAN, ln_eps=1e-6,
    pix_mean=(0.5, 0.5, 0.5), pix_std=(0.5, 0.5, 0.5))

CLIP_VIT_L_336 = TowerSpec(
    name="clip_vit_l_336", arch=ARCH_CLIP, image_size=336, patch_size=14, hidden=1024, heads=16, mlp=4096, layers=24,
    has_cls=True, has_pre_ln=True, conv_bias=False, act=ACT_QUICK_GELU, pool=POOL_MEAN, ln_eps=1e-5,
    pix_mean=_OPENAI_CLIP_MEAN, pix_std=_OPENAI_CLIP_STD)

EVA_VIT_G = TowerSpec(
    name="eva_vit_g", arch=ARCH_EVA, image_size=224, patch_size=14, hidden=1408, heads=16, mlp=6144, layers=39,
    has_cls=True, has_pre_ln=False, conv_bias=True, act=ACT_GELU_ERF, pool=POOL_LN_CLS_LN, ln_eps=1e-6,
    pix_mean=_OPENAI_CLIP_MEAN, pix_std=_OPENAI_CLIP_STD)

# Scaled-down twins: same head_dim / activation / pooling / naming as the real towers (so every kernel variant and
# every awkward dimension is exercised: head_dim 72 and 88, non-multiple-of-64 K, ragged token counts) but 2 layers,
# 2 heads and a 4x4 or 5x5 patch grid, so the CPU oracle finishes in milliseconds.
SIGLIP_TINY = replace(SIGLIP_SO400M, name="siglip_tiny", image_size=56, hidden=144, heads=2, mlp=304, layers=2)
CLIP_TINY = replace(CLIP_VIT_L_336, name="clip_tiny", image_size=70, hidden=128, heads=2, mlp=256, layers=2)
EVA_TINY = replace(EVA_VIT_G, name="eva_tiny", image_size=56, hidden=176, heads=2, mlp=384, layers=2)

SPECS: Dict[str, TowerSpec] = {s.name: s for s in (SIGLIP_SO400M, CLIP_VIT_L_336, EVA_VIT_G, SIGLIP_TINY, CLIP_TINY, EVA_TINY)}

# registry key of the reference's VLMModelFactory (models/model_factory.py:13-17) -> full-size tower
FACTORY_TOWERS: Dict[str, TowerSpec] = {"paligemma": SIGLIP_SO400M, "llava": CLIP_VIT_L_336, "blip2": EVA_VIT_G}


def spec_from_hf_config(cfg, arch: int) -> TowerSpec:
    """Build a spec from a transformers vision config (SiglipVisionConfig / CLIPVisionConfig / Blip2VisionConfig)."""
    base = {ARCH_SIGLIP: SIGLIP_SO400M, ARCH_CLIP: CLIP_VIT_L_336, ARCH_EVA: EVA_VIT_G}[arch]
    act_names = {"gelu_pytorch_tanh": ACT_GELU_TANH, "quick_gelu": ACT_QUICK_GELU, "gelu": ACT_GELU_ERF}
    act = act_names.get(getattr(cfg, "hidden_act", None))
    if act is None:
        raise ValueError(f"unsupported hidden_act {getattr(cfg, 'hidden_act', None)!r}")
    return replace(
        base, name=f"{base.name}_from_hf", image_size=int(cfg.image_size), patch_size=int(cfg.patch_size),
        hidden=int(cfg.hidden_size), heads=int(cfg.num_attention_heads), mlp=int(cfg.intermediate_size),
        layers=int(cfg.num_hidden_layers), act=act, ln_eps=float(cfg.layer_norm_eps))
