"""The plugin surface of the reference, restated: ``VisionBackbone`` and the B200 adapters behind it.

Mirrors reference models/vision_backbone.py:6-67 (constructor ``(processor, vision_model, output_dim, device)``,
abstract ``forward(images) -> [B, D]``, ``unfreeze_last_k_layers``, ``get_lora_target_names``, ``_find_linear``) and
the three adapters models/paligemma.py:88-221, models/llava.py:86-229, models/blip2.py:74-198. What changes is only
what runs inside ``forward``: instead of ``self.vision_model(pixel_values)`` (Hugging Face eager PyTorch) the packed
bf16 weights run through ``libvlmfw_b200.so``. ``self.vision_model`` stays a registered sub-module so
``parameters()``, ``state_dict()`` (``backbone.vision_model.*`` checkpoint keys) and the freeze bookkeeping of
``BaseProbe`` behave exactly as before.
"""
from __future__ import annotations

import os
import re
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, Optional, Sequence, Union

import torch
import torch.nn as nn

from .engine import TowerEngine
from .ingest import ImageIngest
from .specs import ARCH_CLIP, ARCH_EVA, ARCH_SIGLIP, POOL_CLS, POOL_LN_CLS_LN, POOL_LN_MEAN, POOL_MEAN, TowerSpec, \
    spec_from_hf_config


class VisionBackbone(nn.Module, ABC):
    """Uniform adapter: images -> global embeddings [B, D] (reference models/vision_backbone.py:6-67)."""

    def __init__(self, processor, vision_model, output_dim, device):
        super().__init__()
        self.processor = processor
        self.vision_model = vision_model
        self.output_dim = output_dim
        self.device = device
        self.vision_model.to(self.device)

    @abstractmethod
    def forward(self, images):
        ...

    @abstractmethod
    def unfreeze_last_k_layers(self, k, parts, include_embeddings):
        ...

    @abstractmethod
    def get_lora_target_names(self, strategy: Dict) -> List[str]:
        ...

    def _find_linear(self, name_pred: Callable[[str], bool]) -> List[str]:
        return [n for n, m in self.named_modules() if isinstance(m, nn.Linear) and name_pred(n)]


ImagesLike = Union["PIL.Image.Image", Sequence, torch.Tensor]  # noqa: F821


class StagedBatch:
    """A batch whose host->device copy was issued on the backbone's copy stream (``B200VisionBackbone.stage_async``):
    ``tensor`` is the device buffer, ``ready`` the event the consuming stream waits on. Passing it where images are
    expected (``backbone(staged)``, ``probe.predict(staged)``) makes the forward wait for the copy instead of issuing it."""

    __slots__ = ("tensor", "ready", "slot")

    def __init__(self, tensor: torch.Tensor, ready: "torch.cuda.Event", slot: int):
        self.tensor, self.ready, self.slot = tensor, ready, slot

    def __len__(self) -> int:
        return int(self.tensor.shape[0])


class B200VisionBackbone(VisionBackbone):
    """Common machinery of the three B200 adapters: spec inference, weight packing, input staging."""

    ARCH: int = -1
    PROMPT: Optional[str] = None  # the dummy ``text=`` the reference passes to its processor

    def __init__(self, processor, vision_model, output_dim, device, spec: Optional[TowerSpec] = None,
                 compute_dtype: Optional[str] = None):
        super().__init__(processor, vision_model, output_dim, device)
        # "f16" (default) or "bf16": 16-bit format of the tensor-core operands, see VLMFW_CDT_* in vlmfw_b200.h
        self.compute_dtype = compute_dtype or os.environ.get("VLMFW_COMPUTE_DTYPE", "f16")
        if spec is None:
            cfg = getattr(vision_model, "config", None)
            if cfg is None:
                raise ValueError("vision_model has no .config; pass spec= explicitly")
            cfg = getattr(cfg, "vision_config", cfg)
            spec = spec_from_hf_config(cfg, self.ARCH)
        if spec.hidden != output_dim:
            raise ValueError(f"output_dim {output_dim} does not match the tower's hidden size {spec.hidden}")
        self.spec = spec
        self._engines: Dict[tuple, TowerEngine] = {}
        self._ingest: Dict[str, ImageIngest] = {}
        self.gpu_ingest = os.environ.get("VLMFW_INGEST", "gpu") != "hf"
        self._weights_version = 0
        # double-buffered input staging (stage_async): copy stream, two device buffers, "buffer consumed" events
        self._copy_stream: Dict[str, "torch.cuda.Stream"] = {}
        self._stage_bufs: Dict[tuple, torch.Tensor] = {}
        self._stage_free: Dict[tuple, "torch.cuda.Event"] = {}
        self._stage_next = 0
        self.register_load_state_dict_post_hook(lambda module, incompatible: module.invalidate_packed_weights())

    # -- weights -------------------------------------------------------------------------------------------------
    def invalidate_packed_weights(self) -> None:
        """Call after mutating ``vision_model`` parameters in place (load_state_dict does it automatically)."""
        self._weights_version += 1

    def _tower_device(self) -> torch.device:
        # The reference keeps a stale ``self.device`` when trainers build the model on CPU and later ``.to(cuda)`` the
        # probe (probing/train/singletask_trainer.py:45-47, base_trainer.py:35); follow the parameters instead.
        try:
            return next(self.vision_model.parameters()).device
        except StopIteration:
            return torch.device(self.device)

    def _engine(self, pool: int) -> TowerEngine:
        dev = self._tower_device()
        if dev.type != "cuda":
            raise RuntimeError(
                f"{type(self).__name__}: the vision tower lives on '{dev}'. vlmfw_b200 runs this path on a B200 only "
                "and has no CPU fallback — move the model to a CUDA device first.")
        key = (str(dev), pool)
        eng = self._engines.get(key)
        if eng is None or getattr(eng, "_version", -1) != self._weights_version:
            spec = self.spec.with_pool(pool)
            eng = TowerEngine(spec, dev, compute_dtype=self.compute_dtype)
            eng.load_hf_state_dict(self.vision_model.state_dict())
            eng._version = self._weights_version
            self._engines[key] = eng
        return eng

    # -- inputs --------------------------------------------------------------------------------------------------
    def _stage_pixels(self, images: ImagesLike, dev: torch.device) -> torch.Tensor:
        """PIL / list of PIL / uint8 NHWC tensor / float NCHW ``pixel_values`` -> device tensor the engine accepts.

        Images ship as raw uint8 (4x fewer H2D bytes than the fp32 ``pixel_values`` the reference copies,
        models/paligemma.py:108); the rescale + normalize of the HF processor is fused into the patch gather, and
        images that are not already S x S are resized / cropped on the GPU by ``ingest.ImageIngest`` — bit-exact with
        the processor's own uint8 resize. ``gpu_ingest = False`` (or VLMFW_INGEST=hf) keeps the unchanged HF processor
        call for such images instead (CPU bicubic resize, fp32 NCHW ingest).
        """
        s = self.spec.image_size
        if isinstance(images, StagedBatch):
            torch.cuda.current_stream(dev).wait_event(images.ready)
            return images.tensor
        if isinstance(images, torch.Tensor):
            t = images
            if t.dtype == torch.uint8:
                if t.dim() == 3:
                    t = t.unsqueeze(0)
            elif t.is_floating_point():
                if t.dim() == 3:
                    t = t.unsqueeze(0)
                t = t.to(torch.float32)
            else:
                raise TypeError(f"unsupported image tensor dtype {t.dtype}")
            if t.device.type == "cpu" and not t.is_pinned():
                t = t.contiguous().pin_memory()
            return t.to(dev, non_blocking=True)
        import numpy as np

        lst = list(images) if isinstance(images, (list, tuple)) else [images]
        if len(lst) == 0:
            return torch.empty(0, s, s, 3, dtype=torch.uint8, device=dev)
        if all(isinstance(im, torch.Tensor) and im.dtype == torch.uint8 and tuple(im.shape) == (s, s, 3) for im in lst):
            return self._stage_pixels(torch.stack(lst), dev)
        native = all(getattr(im, "size", None) == (s, s) and getattr(im, "mode", None) == "RGB" for im in lst)
        if native:
            host = torch.empty(len(lst), s, s, 3, dtype=torch.uint8).pin_memory()
            arr = host.numpy()
            for i, im in enumerate(lst):
                arr[i] = np.asarray(im)
            return host.to(dev, non_blocking=True)
        if self.gpu_ingest:
            key = str(dev)
            if key not in self._ingest:
                self._ingest[key] = ImageIngest.for_spec(self.spec, dev)
            return self._ingest[key](lst)
        kwargs = {"text": self.PROMPT} if self.PROMPT is not None else {}
        inputs = self.processor(images=lst, return_tensors="pt", **kwargs)
        pv = inputs["pixel_values"].to(torch.float32).contiguous()
        return pv.pin_memory().to(dev, non_blocking=True)

    def stage_async(self, images: ImagesLike) -> StagedBatch:
        """Start the host->device copy of a batch on the backbone's copy stream and return at once.

        The synchronous path (``_stage_pixels``) copies on the compute stream, so copy and tower forward serialise (2 % of
        a SigLIP batch-256 step). Here the copy of batch i+1 runs on its own stream into the other one of two device buffers
        while batch i computes; a buffer is rewritten only after the forward that read it has been enqueued AND finished
        (event recorded by ``_features``). Host side: a pinned uint8 / fp32 tensor is copied as is; anything else goes
        through ``_stage_pixels`` on the copy stream."""
        dev = self._tower_device()
        if dev.type != "cuda":
            raise RuntimeError(f"{type(self).__name__}.stage_async: the tower lives on '{dev}'; there is no CPU fallback")
        key = str(dev)
        cs = self._copy_stream.get(key)
        if cs is None:
            cs = self._copy_stream[key] = torch.cuda.Stream(device=dev)
        slot = self._stage_next & 1
        self._stage_next += 1
        free = self._stage_free.get((key, slot))
        if free is not None:
            cs.wait_event(free)
        with torch.cuda.stream(cs):
            if isinstance(images, torch.Tensor) and images.device.type == "cpu" and images.is_pinned() \
                    and images.dtype in (torch.uint8, torch.float32) and images.dim() == 4:
                bkey = (key, slot, tuple(images.shape), images.dtype)
                buf = self._stage_bufs.get(bkey)
                if buf is None:
                    for k in [k for k in self._stage_bufs if k[:2] == (key, slot)]:
                        del self._stage_bufs[k]
                    buf = self._stage_bufs[bkey] = torch.empty(images.shape, dtype=images.dtype, device=dev)
                buf.copy_(images, non_blocking=True)
                t = buf
            else:
                t = self._stage_pixels(images, dev)
                t.record_stream(torch.cuda.current_stream(dev))
            ready = torch.cuda.Event()
            ready.record(cs)
        return StagedBatch(t, ready, slot)

    def _features(self, images: ImagesLike, pool: int) -> torch.Tensor:
        eng = self._engine(pool)
        with torch.no_grad():  # the reference wraps the tower call in no_grad unconditionally (paligemma.py:110)
            out = eng.forward(self._stage_pixels(images, eng.device))
        if isinstance(images, StagedBatch):  # the staging buffer may be rewritten once this forward has run
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(eng.device))
            self._stage_free[(str(eng.device), images.slot)] = ev
        return out

    # -- reference utilities kept as thin name-matching helpers ----------------------------------------------------
    _ATTN_KEYS: Sequence[str] = (".self_attn.q_proj", ".self_attn.k_proj", ".self_attn.v_proj", ".self_attn.out_proj")
    _MLP_KEYS: Sequence[str] = (".mlp.fc1", ".mlp.fc2")
    _EXTRA_KEYS: Sequence[str] = ("embeddings", "patch_embed", "pos_embed", "position_embedding", "pre_", "post_",
                                  "proj", "projection", "final_layer_norm")

    def _layer_indices(self, named) -> List[int]:
        idx = set()
        for n, _ in named:
            m = re.search(r"encoder\.layers\.(\d+)", n)
            if m:
                idx.add(int(m.group(1)))
        return sorted(idx)

    def unfreeze_last_k_layers(self, k: int = 2, parts: str = "all", include_embeddings: bool = False):
        """Set ``requires_grad`` on the last k encoder layers (reference models/paligemma.py:114-183). Note that, as in
        the reference, features still carry no gradient: the tower forward runs under ``no_grad``."""
        ordered = self._layer_indices(self.vision_model.named_modules())
        if not ordered:
            raise RuntimeError("Non trovo 'encoder.layers.<idx>' in self.vision_model.")
        selected = set(ordered[-int(k):]) if int(k) > 0 else set()

        def want(name: str) -> bool:
            m = re.search(r"encoder\.layers\.(\d+)", name)
            if not (m and int(m.group(1)) in selected):
                return False
            attn = any(key in name for key in self._ATTN_KEYS)
            mlp = any(key in name for key in self._MLP_KEYS)
            norm = ".layer_norm" in name or ".ln" in name or ".norm" in name
            if parts == "all":
                return True
            if parts == "attn":
                return attn or norm
            if parts == "mlp":
                return mlp or norm
            return False

        for n, p in self.vision_model.named_parameters():
            if want(n) or (include_embeddings and any(key in n for key in self._EXTRA_KEYS)):
                p.requires_grad = True
        print(f"[unfreeze_last_k_layers] Scongelati {len(selected)} layer (indici: {sorted(selected)})")

    def get_lora_target_names(self, strategy: Dict) -> List[str]:
        """Names of the nn.Linear modules in the last K encoder layers (reference models/paligemma.py:186-221)."""
        last_k = int(strategy.get("last_k", 2))
        attn_only = bool(strategy.get("attn_only", True))
        idx = self._layer_indices(self.named_modules())
        if not idx:
            return []
        selected = set(range(max(0, idx[-1] - last_k + 1), idx[-1] + 1))
        keys = tuple(k.lstrip(".") for k in self._ATTN_KEYS) + (() if attn_only else tuple(k.lstrip(".") for k in self._MLP_KEYS))

        def is_target(n: str) -> bool:
            m = re.search(r"encoder\.layers\.(\d+)", n)
            return bool(m) and int(m.group(1)) in selected and any(k in n for k in keys)

        return self._find_linear(is_target)


class PaliGemmaBackbone(B200VisionBackbone):
    """SigLIP tower of PaLI-Gemma: post-LN on all tokens, then mean over tokens (reference models/paligemma.py:100-112)."""

    ARCH = ARCH_SIGLIP
    PROMPT = "<image> Describe this image"

    def forward(self, images):
        return self._features(images, POOL_LN_MEAN)


class LLaVABackbone(B200VisionBackbone):
    """CLIP ViT-L/14-336 tower of LLaVA-1.5: mean (default) or CLS over the raw last hidden state
    (reference models/llava.py:101-119)."""

    ARCH = ARCH_CLIP
    PROMPT = "Describe this image"

    def forward(self, images, strategy: str = "mean"):
        if strategy == "cls":
            return self._features(images, POOL_CLS)
        if strategy == "mean":
            return self._features(images, POOL_MEAN)
        raise ValueError(f"Strategia pooling '{strategy}' non supportata")


class BLIP2Backbone(B200VisionBackbone):
    """EVA ViT-g tower of BLIP-2: ``pooler_output`` = LN(LN(h)[:, 0]) (reference models/blip2.py:83-95)."""

    ARCH = ARCH_EVA
    PROMPT = None
    _ATTN_KEYS = (".self_attn.qkv", ".self_attn.projection")

    def forward(self, images):
        return self._features(images, POOL_LN_CLS_LN)


BACKBONES = {ARCH_SIGLIP: PaliGemmaBackbone, ARCH_CLIP: LLaVABackbone, ARCH_EVA: BLIP2Backbone}


def backbone_from_spec(spec: TowerSpec, state_dict=None, device="cuda", processor=None, seed: int = 0,
                       compute_dtype: Optional[str] = None):
    """Offline constructor (tests, bench): HF vision module from config + given (or seeded) weights, wrapped in the
    matching B200 adapter. No network, no ``from_pretrained``."""
    import transformers as tf

    from .weights import seeded_hf_state_dict

    act = {0: "gelu_pytorch_tanh", 1: "quick_gelu", 2: "gelu"}[spec.act]
    common = dict(hidden_size=spec.hidden, intermediate_size=spec.mlp, num_hidden_layers=spec.layers,
                  num_attention_heads=spec.heads, image_size=spec.image_size, patch_size=spec.patch_size,
                  hidden_act=act, layer_norm_eps=spec.ln_eps)
    with torch.device("meta"):
        if spec.arch == ARCH_SIGLIP:
            module = tf.SiglipVisionModel(tf.SiglipVisionConfig(vision_use_head=False, **common))
        elif spec.arch == ARCH_CLIP:
            module = tf.CLIPVisionModel(tf.CLIPVisionConfig(**common))
        else:
            module = tf.Blip2VisionModel(tf.Blip2VisionConfig(qkv_bias=True, **common))
    module = module.to_empty(device="cpu")
    sd = dict(state_dict) if state_dict is not None else seeded_hf_state_dict(spec, seed)
    for name, buf in module.named_buffers():
        if name.endswith("position_ids"):  # non-persistent buffer, left uninitialised by to_empty()
            buf.copy_(torch.arange(buf.shape[-1]).expand(buf.shape))
    module.load_state_dict(sd, strict=True)
    module.eval()
    return BACKBONES[spec.arch](processor, module, spec.hidden, torch.device(device), spec=spec,
                                compute_dtype=compute_dtype)
