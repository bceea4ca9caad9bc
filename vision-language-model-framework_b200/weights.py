"""Weight naming: Hugging Face vision-tower state-dict names -> the canonical names of ``vlmfw_tower_set_weight``.

The reference hands the adapters a HF module (``self.model.vision_tower`` / ``self.model.vision_model`` —
reference models/paligemma.py:70, models/llava.py:68, models/blip2.py:56) and its checkpoints carry
``backbone.vision_model.<hf name>`` keys (reference probing/train/base_trainer.py:142). This module is the only place
that knows those names:

* SigLIP  (hf modeling_siglip.py:116-187, :229-361): ``vision_model.embeddings.*``, ``vision_model.encoder.layers.i.*``
* CLIP    (hf modeling_clip.py:138-218, :261-384):   same plus ``class_embedding`` / ``pre_layrnorm`` (sic)
* BLIP-2  (hf modeling_blip_2.py:184-256, :283-352): no ``vision_model.`` prefix, fused ``self_attn.qkv``,
  ``self_attn.projection``, position/class embeddings are bare Parameters with leading singleton dims

Everything here is host logic on CPU tensors and is covered by the ``not gpu`` tests.
"""
from __future__ import annotations

import math
from typing import Dict, Iterable, List, Mapping, Tuple

import torch

from .specs import ARCH_CLIP, ARCH_EVA, ARCH_SIGLIP, TowerSpec


def canonical_names(spec: TowerSpec) -> List[str]:
    """Every tensor name the C handle requires for this spec (mirrors vlmfw_tower_create's slot table)."""
    names = ["patch.w", "pos"]
    if spec.conv_bias:
        names.append("patch.b")
    if spec.has_cls:
        names.append("cls")
    if spec.has_pre_ln:
        names += ["pre_ln.w", "pre_ln.b"]
    from .specs import POOL_LN_CLS_LN, POOL_LN_MEAN
    if spec.pool in (POOL_LN_MEAN, POOL_LN_CLS_LN):
        names += ["post_ln.w", "post_ln.b"]
    for i in range(spec.layers):
        p = f"L{i}."
        names += [p + s for s in ("ln1.w", "ln1.b", "qkv.w", "qkv.b", "o.w", "o.b", "ln2.w", "ln2.b", "fc1.w", "fc1.b",
                                  "fc2.w", "fc2.b")]
    return names


def canonical_shapes(spec: TowerSpec) -> Dict[str, Tuple[int, ...]]:
    d, i, s, p = spec.hidden, spec.mlp, spec.tokens, spec.patch_size
    shapes: Dict[str, Tuple[int, ...]] = {
        "patch.w": (d, 3 * p * p), "patch.b": (d,), "cls": (d,), "pos": (s, d),
        "pre_ln.w": (d,), "pre_ln.b": (d,), "post_ln.w": (d,), "post_ln.b": (d,),
    }
    for li in range(spec.layers):
        q = f"L{li}."
        shapes.update({
            q + "ln1.w": (d,), q + "ln1.b": (d,), q + "qkv.w": (3 * d, d), q + "qkv.b": (3 * d,),
            q + "o.w": (d, d), q + "o.b": (d,), q + "ln2.w": (d,), q + "ln2.b": (d,),
            q + "fc1.w": (i, d), q + "fc1.b": (i,), q + "fc2.w": (d, i), q + "fc2.b": (d,),
        })
    return {k: v for k, v in shapes.items() if k in set(canonical_names(spec))}


def _strip_prefix(sd: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Accept ``backbone.vision_model.vision_model.x`` (probe checkpoint), ``vision_model.x`` (HF) or bare ``x``."""
    out: Dict[str, torch.Tensor] = {}
    for k, v in sd.items():
        key = k
        for pre in ("backbone.vision_model.", "backbone."):
            if key.startswith(pre):
                key = key[len(pre):]
                break
        while key.startswith("vision_model."):
            key = key[len("vision_model."):]
        out[key] = v
    return out


def hf_to_canonical(spec: TowerSpec, state_dict: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Translate a HF vision-tower state dict into ``{canonical name: fp32 contiguous tensor}``.

    q/k/v projections are concatenated into one ``[3D, D]`` matrix (rows q | k | v) so the tower runs one fused QKV
    GEMM; BLIP-2 ships that layout already (modeling_blip_2.py:301-312). Raises ``KeyError`` naming the first
    missing tensor.
    """
    sd = _strip_prefix(state_dict)
    out: Dict[str, torch.Tensor] = {}

    def get(name: str) -> torch.Tensor:
        if name not in sd:
            raise KeyError(f"vision tower state dict lacks '{name}' (needed for {spec.name})")
        return sd[name].detach().to(torch.float32)

    d = spec.hidden
    out["patch.w"] = get("embeddings.patch_embedding.weight").reshape(d, -1)
    if spec.conv_bias:
        out["patch.b"] = get("embeddings.patch_embedding.bias")
    if spec.arch == ARCH_EVA:
        out["pos"] = get("embeddings.position_embedding").reshape(-1, d)
        out["cls"] = get("embeddings.class_embedding").reshape(d)
    else:
        out["pos"] = get("embeddings.position_embedding.weight")
        if spec.has_cls:
            out["cls"] = get("embeddings.class_embedding").reshape(d)
    if spec.has_pre_ln:
        out["pre_ln.w"] = get("pre_layrnorm.weight")
        out["pre_ln.b"] = get("pre_layrnorm.bias")
    if "post_ln.w" in canonical_names(spec):
        out["post_ln.w"] = get("post_layernorm.weight")
        out["post_ln.b"] = get("post_layernorm.bias")
    for i in range(spec.layers):
        src, dst = f"encoder.layers.{i}.", f"L{i}."
        out[dst + "ln1.w"] = get(src + "layer_norm1.weight")
        out[dst + "ln1.b"] = get(src + "layer_norm1.bias")
        out[dst + "ln2.w"] = get(src + "layer_norm2.weight")
        out[dst + "ln2.b"] = get(src + "layer_norm2.bias")
        if spec.arch == ARCH_EVA:
            out[dst + "qkv.w"] = get(src + "self_attn.qkv.weight")
            if (src + "self_attn.qkv.bias") in sd:
                out[dst + "qkv.b"] = get(src + "self_attn.qkv.bias")
            else:  # qkv_bias=False checkpoints
                out[dst + "qkv.b"] = torch.zeros(3 * d)
            out[dst + "o.w"] = get(src + "self_attn.projection.weight")
            out[dst + "o.b"] = get(src + "self_attn.projection.bias")
        else:
            out[dst + "qkv.w"] = torch.cat([get(src + f"self_attn.{n}_proj.weight") for n in "qkv"], dim=0)
            out[dst + "qkv.b"] = torch.cat([get(src + f"self_attn.{n}_proj.bias") for n in "qkv"], dim=0)
            out[dst + "o.w"] = get(src + "self_attn.out_proj.weight")
            out[dst + "o.b"] = get(src + "self_attn.out_proj.bias")
        out[dst + "fc1.w"] = get(src + "mlp.fc1.weight")
        out[dst + "fc1.b"] = get(src + "mlp.fc1.bias")
        out[dst + "fc2.w"] = get(src + "mlp.fc2.weight")
        out[dst + "fc2.b"] = get(src + "mlp.fc2.bias")
    shapes = canonical_shapes(spec)
    for k, v in out.items():
        want = shapes[k]
        n = 1
        for x in want:
            n *= x
        if v.numel() != n:
            raise ValueError(f"'{k}' has {tuple(v.shape)} ({v.numel()} elements), expected {want}")
        out[k] = v.reshape(want).contiguous()
    return out


def hf_param_names(spec: TowerSpec) -> List[Tuple[str, Tuple[int, ...]]]:
    """(name, shape) of every parameter of the HF module for this spec, in HF's naming (with its own prefix)."""
    d, i, s, p = spec.hidden, spec.mlp, spec.tokens, spec.patch_size
    pre = "" if spec.arch == ARCH_EVA else "vision_model."
    out: List[Tuple[str, Tuple[int, ...]]] = []
    if spec.arch == ARCH_EVA:
        out += [("embeddings.class_embedding", (1, 1, d)), ("embeddings.position_embedding", (1, s, d))]
    elif spec.arch == ARCH_CLIP:
        out += [(pre + "embeddings.class_embedding", (d,))]
    out += [(pre + "embeddings.patch_embedding.weight", (d, 3, p, p))]
    if spec.conv_bias:
        out += [(pre + "embeddings.patch_embedding.bias", (d,))]
    if spec.arch != ARCH_EVA:
        out += [(pre + "embeddings.position_embedding.weight", (s, d))]
    if spec.has_pre_ln:
        out += [(pre + "pre_layrnorm.weight", (d,)), (pre + "pre_layrnorm.bias", (d,))]
    for li in range(spec.layers):
        q = f"{pre}encoder.layers.{li}."
        if spec.arch == ARCH_EVA:
            out += [(q + "self_attn.qkv.weight", (3 * d, d)), (q + "self_attn.qkv.bias", (3 * d,)),
                    (q + "self_attn.projection.weight", (d, d)), (q + "self_attn.projection.bias", (d,))]
        else:
            for n in ("k", "v", "q", "out"):
                out += [(q + f"self_attn.{n}_proj.weight", (d, d)), (q + f"self_attn.{n}_proj.bias", (d,))]
        out += [(q + "layer_norm1.weight", (d,)), (q + "layer_norm1.bias", (d,)),
                (q + "mlp.fc1.weight", (i, d)), (q + "mlp.fc1.bias", (i,)),
                (q + "mlp.fc2.weight", (d, i)), (q + "mlp.fc2.bias", (d,)),
                (q + "layer_norm2.weight", (d,)), (q + "layer_norm2.bias", (d,))]
    out += [(pre + "post_layernorm.weight", (d,)), (pre + "post_layernorm.bias", (d,))]
    return out


def seeded_hf_state_dict(spec: TowerSpec, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded, NON-degenerate random-init weights in HF naming (fp32, CPU).

    HF's default init leaves every bias 0 and every LayerNorm at (1, 0), and Blip2VisionConfig.initializer_range is
    1e-10, so default-initialised towers would not exercise the bias / affine paths (SURVEY.md §8c caveats). Here
    matrices ~ N(0, s) with s chosen per fan-in so activations stay O(1) through 39 layers, biases ~ N(0, 0.02),
    LN gamma ~ 1 + N(0, 0.1), LN beta ~ N(0, 0.1). Each tensor draws from its own generator keyed by (seed, name)
    so the values do not depend on iteration order.
    """
    import zlib

    out: Dict[str, torch.Tensor] = {}
    for name, shape in hf_param_names(spec):
        g = torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
        leaf = name.rsplit(".", 1)[-1]
        is_ln = "layer_norm" in name or "layernorm" in name or "layrnorm" in name
        if is_ln and leaf == "weight":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif is_ln and leaf == "bias":
            t = 0.1 * torch.randn(shape, generator=g)
        elif leaf == "bias":
            t = 0.02 * torch.randn(shape, generator=g)
        elif "position_embedding" in name or "class_embedding" in name:
            t = 0.5 * torch.randn(shape, generator=g)
        elif "patch_embedding.weight" in name:
            t = torch.randn(shape, generator=g) * (1.0 / (3 * spec.patch_size * spec.patch_size) ** 0.5)
        else:  # dense matrices [out, in]
            t = torch.randn(shape, generator=g) * (0.7 / shape[-1] ** 0.5)
        out[name] = t.to(torch.float32)
    return out


def seeded_images_u8(n: int, size: int, seed: int = 0, smooth: bool = False) -> torch.Tensor:
    """Synthetic RGB images [n, size, size, 3] uint8: white noise, or smooth low-frequency fields."""
    import numpy as np

    imgs = []
    for i in range(n):
        rng = np.random.default_rng(seed * 7919 + i)
        if smooth:
            coarse = torch.from_numpy(rng.normal(size=(1, 3, 8, 8)).astype("float32"))
            up = torch.nn.functional.interpolate(coarse, size=(size, size), mode="bicubic", align_corners=False)[0]
            up = (up - up.amin()) / (up.amax() - up.amin() + 1e-6)
            imgs.append((up.permute(1, 2, 0) * 255.0).round().clamp(0, 255).to(torch.uint8))
        else:
            imgs.append(torch.from_numpy(rng.integers(0, 256, (size, size, 3), dtype=np.uint8)))
    return torch.stack(imgs)


def seeded_head_params(dim: int, classes: int, hidden: int = 0, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded, non-degenerate probe-head parameters in the reference's nn.Sequential keys (``0.*`` BatchNorm1d, ``2.*``
    Linear, ``5.*`` second Linear of the deeper head). A fresh BatchNorm1d is the identity up to 1/sqrt(1+eps), which
    would not exercise the affine / running-statistics path."""
    g = torch.Generator().manual_seed(seed)
    mid = hidden if hidden > 0 else classes
    p = {
        "0.weight": 1.0 + 0.1 * torch.randn(dim, generator=g),
        "0.bias": 0.1 * torch.randn(dim, generator=g),
        "0.running_mean": 0.2 * torch.randn(dim, generator=g),
        "0.running_var": 0.5 + torch.rand(dim, generator=g),
        "0.num_batches_tracked": torch.tensor(7),
        "2.weight": torch.randn(mid, dim, generator=g) / math.sqrt(dim),
        "2.bias": 0.1 * torch.randn(mid, generator=g),
    }
    if hidden > 0:
        p["5.weight"] = torch.randn(classes, hidden, generator=g) / math.sqrt(hidden)
        p["5.bias"] = 0.1 * torch.randn(classes, generator=g)
    return p
