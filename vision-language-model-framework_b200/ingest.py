"""GPU image ingest — what stands in for ``self.processor(images=..., return_tensors="pt")`` in the reference adapters
(models/paligemma.py:108, models/llava.py:109, models/blip2.py:91).

The HF processors resize on the CPU, one image at a time, and hand the tower fp32 ``pixel_values`` (4x the bytes of the
raw image). Here the raw RGB bytes of a ragged batch are packed into one pinned buffer, copied once, and resized /
cropped on the GPU by ``vlmfw_preproc_run`` to uint8 ``[B, S, S, 3]`` — bit-exact with the processor's own resized
image — which the tower ingests directly (rescale + normalize fused into its patch gather).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _lib
from .engine import require_cuda
from .specs import ARCH_CLIP, TowerSpec

RESIZE_SQUARE, RESIZE_SHORTEST_CROP = 0, 1


def resize_mode_for(spec: TowerSpec) -> int:
    """CLIPImageProcessor: shortest edge + centre crop; Siglip / Blip processors: plain square resize."""
    return RESIZE_SHORTEST_CROP if spec.arch == ARCH_CLIP else RESIZE_SQUARE


def resize_weights(in_size: int, out_full: int, offset: int = 0, count: Optional[int] = None
                   ) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """Host-only: (xmin, xsize, int16 weights [count, taps], precision) the kernels use for one axis."""
    lib = _lib.load()
    count = out_full - offset if count is None else count
    taps, prec = C.c_int32(), C.c_int32()
    _lib.check(lib.vlmfw_resize_weights(in_size, out_full, offset, count, None, None, None, 0, C.byref(taps),
                                        C.byref(prec)))
    xmin = np.zeros(count, np.int32)
    xsize = np.zeros(count, np.int32)
    w = np.zeros((count, taps.value), np.int16)
    _lib.check(lib.vlmfw_resize_weights(in_size, out_full, offset, count, xmin.ctypes.data, xsize.ctypes.data,
                                        w.ctypes.data, taps.value, C.byref(taps), C.byref(prec)))
    return xmin, xsize, w, int(prec.value)


def pack_images(images: Sequence) -> Tuple[torch.Tensor, np.ndarray, np.ndarray, np.ndarray]:
    """PIL images / HWC uint8 arrays of any sizes -> (pinned uint8 buffer, offsets int64, heights int32, widths int32).
    Every image starts at a 16-byte aligned offset. Non-RGB PIL images are converted like the processors do."""
    arrays: List[np.ndarray] = []
    for im in images:
        if isinstance(im, torch.Tensor):
            a = im.detach().cpu().numpy()
        elif hasattr(im, "convert"):  # PIL
            if im.mode != "RGB":
                im = im.convert("RGB")
            a = np.asarray(im)
        else:
            a = np.asarray(im)
        if a.dtype != np.uint8 or a.ndim != 3 or a.shape[2] != 3:
            raise ValueError(f"images must be RGB uint8 [H, W, 3], got {a.dtype} {a.shape}")
        arrays.append(a)
    n = len(arrays)
    heights = np.array([a.shape[0] for a in arrays], np.int32)
    widths = np.array([a.shape[1] for a in arrays], np.int32)
    sizes = heights.astype(np.int64) * widths.astype(np.int64) * 3
    offsets = np.zeros(n, np.int64)
    total = 0
    for i in range(n):
        offsets[i] = total
        total += (int(sizes[i]) + 15) & ~15
    host = torch.empty(max(total, 16), dtype=torch.uint8).pin_memory()
    view = host.numpy()
    for i, a in enumerate(arrays):
        view[offsets[i]:offsets[i] + sizes[i]] = a.reshape(-1)
    return host, offsets, heights, widths


class ImageIngest:
    """Owner of one ``vlmfw_preproc`` handle (mode + output size)."""

    def __init__(self, mode: int, out_size: int, device="cuda"):
        self.device = require_cuda(device)
        self.mode, self.out_size = mode, out_size
        self._lib = _lib.load()
        h = C.c_void_p()
        _lib.check(self._lib.vlmfw_preproc_create(mode, out_size, C.byref(h)))
        self._h = h
        self._ws: Optional[torch.Tensor] = None

    @classmethod
    def for_spec(cls, spec: TowerSpec, device="cuda") -> "ImageIngest":
        return cls(resize_mode_for(spec), spec.image_size, device)

    def close(self) -> None:
        if getattr(self, "_h", None):
            torch.cuda.synchronize(self.device)
            self._lib.vlmfw_preproc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run_packed(self, packed: torch.Tensor, offsets: np.ndarray, heights: np.ndarray, widths: np.ndarray,
                   out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """packed: DEVICE uint8 buffer; returns uint8 [B, S, S, 3] on the device (asynchronous on the current stream)."""
        if not packed.is_cuda or packed.dtype != torch.uint8:
            raise RuntimeError("packed must be a CUDA uint8 tensor: vlmfw_b200 has no CPU fallback")
        b, s = len(heights), self.out_size
        offsets = np.ascontiguousarray(offsets, np.int64)
        heights = np.ascontiguousarray(heights, np.int32)
        widths = np.ascontiguousarray(widths, np.int32)
        if out is None:
            out = torch.empty(b, s, s, 3, dtype=torch.uint8, device=self.device)
        if b == 0:
            return out
        need = int(self._lib.vlmfw_preproc_workspace_bytes(self._h, heights.ctypes.data, widths.ctypes.data, b))
        if need == 0:
            raise _lib.VlmfwError(_lib.E_INVALID, _lib.last_error())
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need + need // 4, dtype=torch.uint8, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(self._lib.vlmfw_preproc_run(
                self._h, packed.data_ptr(), packed.numel(), offsets.ctypes.data, heights.ctypes.data,
                widths.ctypes.data, b, out.data_ptr(), self._ws.data_ptr(), self._ws.numel(),
                torch.cuda.current_stream(self.device).cuda_stream))
        return out

    def __call__(self, images: Sequence, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """List of PIL images / HWC uint8 arrays (host) -> device uint8 [B, S, S, 3]."""
        host, offsets, heights, widths = pack_images(images)
        packed = host.to(self.device, non_blocking=True)
        res = self.run_packed(packed, offsets, heights, widths, out)
        # `host` / `packed` are released to torch's caching allocators, which are stream-ordered for the current stream
        return res

    @property
    def last_launches(self) -> int:
        return int(self._lib.vlmfw_preproc_last_launches(self._h))
