"""TowerEngine — Python owner of one ``vlmfw_tower`` handle.

PyTorch is plumbing here (device memory for the workspace / outputs, the current CUDA stream); every kernel on the
path lives in ``libvlmfw_b200.so``. The engine replaces the ``self.vision_model(pixel_values)`` + pooling lines of the
reference adapters (models/paligemma.py:108-112, models/llava.py:109-117, models/blip2.py:91-95).
"""
from __future__ import annotations

import ctypes as C
import warnings
from typing import Dict, Mapping, Optional

import torch

from . import _lib
from .specs import PIX_F32_NCHW, PIX_U8_NHWC, TowerSpec
from .weights import canonical_names, hf_to_canonical


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def require_cuda(device) -> torch.device:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError(
            f"vlmfw_b200 runs the vision-tower path on sm_100a only; got device '{device}'. There is no CPU fallback.")
    if not torch.cuda.is_available():
        raise RuntimeError("vlmfw_b200: CUDA is not available and there is no CPU fallback for the vision-tower path")
    return device


class TowerEngine:
    """One vision tower with packed 16-bit weights resident in HBM, replicated per GPU (data parallel).

    ``compute_dtype``: "f16" (default) or "bf16". fp16 operands are the only mode inside the north-star tolerance
    (probe-logit max abs error <= 1e-2): ``tolerance_ok`` is False for a bf16 tower, whose weight rounding alone costs
    ~1.5e-2 (DESIGN.md numerics), and constructing one warns. The fp16 mode does not depend on the weights' range (a matrix
    beyond +-65504 is stored divided by a power of two, ``weight_scale``); what an fp16 forward can still do is clamp
    an ACTIVATION at +-65504 — ``saturation_count()`` reports it.
    """

    def __init__(self, spec: TowerSpec, device="cuda", compute_dtype: str = "f16"):
        self.spec = spec
        self.device = require_cuda(device)
        self._lib = _lib.load()
        if compute_dtype not in ("f16", "bf16"):
            raise ValueError("compute_dtype must be 'f16' or 'bf16'")
        self.compute_dtype = compute_dtype
        if compute_dtype == "bf16":
            warnings.warn("vlmfw_b200: bf16 tensor-core operands do not meet the 1e-2 probe-logit tolerance (weight rounding "
                          "alone ~1.5e-2; cosine >= 0.999 and labels hold). Use compute_dtype='f16' for parity runs.",
                          RuntimeWarning, stacklevel=2)
        self._h = None
        self._create()
        self._ws: Optional[torch.Tensor] = None
        self._ws_batch = 0
        self._graphs: Dict[tuple, tuple] = {}

    def _create(self) -> None:
        spec = self.spec
        self.close()
        cfg = _lib.TowerCfg(
            arch=spec.arch, image_size=spec.image_size, patch_size=spec.patch_size, hidden=spec.hidden,
            heads=spec.heads, mlp=spec.mlp, layers=spec.layers, has_cls=int(spec.has_cls),
            has_pre_ln=int(spec.has_pre_ln), conv_bias=int(spec.conv_bias), act=spec.act, pool=spec.pool,
            ln_eps=spec.ln_eps, pix_mean=(C.c_float * 3)(*spec.pix_mean), pix_std=(C.c_float * 3)(*spec.pix_std),
            compute_dtype=_lib.CDT_F16 if self.compute_dtype == "f16" else _lib.CDT_BF16)
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vlmfw_tower_create(C.byref(cfg), C.byref(handle)))
        self._h = handle

    # ------------------------------------------------------------------ lifetime
    def close(self) -> None:
        if getattr(self, "_h", None):
            self._lib.vlmfw_tower_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ weights
    @property
    def tolerance_ok(self) -> bool:
        """True when this tower's arithmetic is the one the parity tests hold to the north-star tolerance (fp16 operands)."""
        return self.compute_dtype == "f16"

    @property
    def cls_tail(self) -> bool:
        """True when the last encoder layer runs for the class-token rows only (towers pooled from the class token: EVA's
        ``pooler_output``, CLIP in "cls" mode). The pooled features do not change; ``hidden_state`` then holds final
        values in row 0 of every image only."""
        return bool(self._lib.vlmfw_tower_cls_tail(self._h))

    def set_cls_tail(self, enable: bool) -> None:
        """Switch the class-token-only last layer on / off (off: every row of the last layer is computed). Captured graphs
        are dropped: the kernel sequence changes."""
        _lib.check(self._lib.vlmfw_tower_set_cls_tail(self._h, int(bool(enable))))
        self._graphs.clear()

    def saturation_count(self, reset: bool = False) -> int:
        """16-bit activation stores that clamped at +-65504 since creation / the last reset (0: no forward clipped).
        Synchronises the current stream."""
        n = C.c_int64(0)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vlmfw_tower_saturation(self._h, C.byref(n), int(reset), _stream_ptr(self.device)))
        return int(n.value)

    def weight_scale(self, name: str) -> float:
        """Power of two the 16-bit operand of canonical matrix ``name`` is stored divided by (1.0 unless it left the fp16 range)."""
        v = C.c_float(0)
        _lib.check(self._lib.vlmfw_tower_weight_scale(self._h, name.encode(), C.byref(v)))
        return float(v.value)

    def load_canonical(self, tensors: Mapping[str, torch.Tensor]) -> None:
        """Upload ``{canonical name: tensor}`` (any device / float dtype); matrices are repacked to 16 bits on the GPU and
        the derived operands (LayerNorm folds) are built. A non-finite weight raises ``VlmfwError`` (VLMFW_E_RANGE)."""
        self._load_canonical(tensors)

    def _load_canonical(self, tensors: Mapping[str, torch.Tensor]) -> None:
        with torch.cuda.device(self.device):
            st = _stream_ptr(self.device)
            for name in canonical_names(self.spec):
                if name not in tensors:
                    raise KeyError(f"missing tower weight '{name}'")
                t = tensors[name].detach().to(device=self.device, dtype=torch.float32).contiguous()
                _lib.check(self._lib.vlmfw_tower_set_weight(self._h, name.encode(), t.data_ptr(), t.numel(), st))
                # the copy/pack is stream-ordered; keep `t` alive until it has run
                torch.cuda.current_stream(self.device).synchronize()
            _lib.check(self._lib.vlmfw_tower_prepare(self._h, st))
        self._graphs.clear()

    def load_hf_state_dict(self, state_dict: Mapping[str, torch.Tensor]) -> None:
        self.load_canonical(hf_to_canonical(self.spec, state_dict))

    # ------------------------------------------------------------------ forward
    def workspace_bytes(self, batch: int) -> int:
        return int(self._lib.vlmfw_tower_workspace_bytes(self._h, batch))

    def _workspace(self, batch: int) -> torch.Tensor:
        if self._ws is None or self._ws_batch < batch:
            self._ws = None
            nbytes = self.workspace_bytes(batch)
            # torch's caching allocator returns 512-byte aligned blocks; over-allocate and slice to 1024
            raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=self.device)
            off = (-raw.data_ptr()) % 1024
            self._ws = raw[off:off + nbytes]
            self._ws_raw = raw
            self._ws_batch = batch
            self._graphs.clear()
        return self._ws

    def _check_pixels(self, pixels: torch.Tensor) -> int:
        s = self.spec.image_size
        if pixels.device != self.device and not (pixels.device.type == "cuda" and self.device.index is None):
            raise ValueError(f"pixels live on {pixels.device}, tower on {self.device}")
        if pixels.dtype == torch.uint8:
            if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (s, s, 3):
                raise ValueError(f"uint8 pixels must be [B,{s},{s},3] (NHWC), got {tuple(pixels.shape)}")
            return PIX_U8_NHWC
        if pixels.dtype == torch.float32:
            if pixels.dim() != 4 or tuple(pixels.shape[1:]) != (3, s, s):
                raise ValueError(f"float pixels must be [B,3,{s},{s}] (NCHW pixel_values), got {tuple(pixels.shape)}")
            return PIX_F32_NCHW
        raise ValueError(f"pixels must be uint8 NHWC or float32 NCHW, got {pixels.dtype}")

    def forward(self, pixels: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """pixels (device, uint8 NHWC raw RGB or float32 NCHW ``pixel_values``) -> pooled features [B, D] fp32."""
        fmt = self._check_pixels(pixels)
        pixels = pixels.contiguous()
        b = pixels.shape[0]
        if b == 0:
            return torch.empty(0, self.spec.hidden, dtype=torch.float32, device=self.device)
        ws = self._workspace(b)
        if out is None:
            out = torch.empty(b, self.spec.hidden, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vlmfw_tower_forward(self._h, pixels.data_ptr(), fmt, b, ws.data_ptr(), ws.numel(),
                                                     out.data_ptr(), _stream_ptr(self.device)))
        return out

    __call__ = forward

    def hidden_state(self, batch: int) -> torch.Tensor:
        """Final residual stream [B, S, D] fp32 of the most recent forward (a view into the workspace)."""
        ws = self._workspace(batch)
        n = batch * self.spec.tokens * self.spec.hidden
        p16 = self._lib.vlmfw_tower_hidden16_ptr(self._h, ws.data_ptr(), batch)
        if p16:  # fp16 towers: the residual stream is the 16-bit buffer itself
            off = int(p16) - ws.data_ptr()
            dt = torch.float16 if self.compute_dtype == "f16" else torch.bfloat16
            return ws[off: off + n * 2].view(dt).view(batch, self.spec.tokens, self.spec.hidden)
        return ws[: n * 4].view(torch.float32).view(batch, self.spec.tokens, self.spec.hidden)

    @property
    def last_launches(self) -> int:
        return int(self._lib.vlmfw_tower_last_launches(self._h))

    # ------------------------------------------------------------------ CUDA graph replay
    def graphed(self, pixels: torch.Tensor, out: torch.Tensor):
        """Capture forward(pixels -> out) for these exact buffers into a CUDA graph; returns a ``replay()`` callable.

        The kernel sequence of one forward is fixed for a given (spec, B), so the ~170–280 launches replay from one
        ``cudaGraphLaunch``.
        """
        key = (pixels.data_ptr(), out.data_ptr(), pixels.shape[0], pixels.dtype)
        if key not in self._graphs:
            self.forward(pixels, out)  # warm-up: tensor maps, func attributes, derived weights
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self.forward(pixels, out)
            self._graphs[key] = (g, self.last_launches)
        g, _ = self._graphs[key]
        return g.replay

    # ------------------------------------------------------------------ host-buffer C-ABI path
    def forward_host(self, pixels_u8: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """HOST uint8 NHWC -> HOST fp32 [B, D] through ``vlmfw_tower_forward_host`` (H2D + forward + D2H, synchronous)."""
        if pixels_u8.device.type != "cpu" or pixels_u8.dtype != torch.uint8:
            raise ValueError("forward_host takes a CPU uint8 tensor [B,H,W,3]")
        pixels_u8 = pixels_u8.contiguous()
        b = pixels_u8.shape[0]
        if out is None:
            out = torch.empty(b, self.spec.hidden, dtype=torch.float32).pin_memory()
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vlmfw_tower_forward_host(self._h, pixels_u8.data_ptr(), b, out.data_ptr()))
        return out
