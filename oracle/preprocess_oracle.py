"""TEST INFRASTRUCTURE — CPU restatement of the image preprocessing in front of the vision towers.

The reference builds its processor with ``AutoProcessor.from_pretrained(model_id, use_fast=True)``
(models/base_model.py:31) and calls it in every adapter (models/paligemma.py:108, llava.py:109, blip2.py:91). On
transformers 5.5 that is the torchvision backend (hf:image_processing_backends.py TorchvisionBackend):

    PIL -> RGB -> uint8 CHW tensor -> tvF.resize(uint8, size, BICUBIC, antialias=True) [-> center crop]
        -> fused rescale + normalize in fp32

``tvF.resize`` on a uint8 CPU tensor runs ATen's separable anti-aliased kernel in FIXED POINT, mimicking Pillow
(third-party: PyTorch 2.11 aten/src/ATen/native/cpu/UpSampleKernel.cpp, ``_compute_index_ranges_int16_weights`` and
``basic_loop_aa_horizontal / _vertical<uint8_t>`` — sources not in /root/reference; restated here from the published
algorithm and pinned against the processors' committed and live outputs in tests/test_preprocess_oracle.py):

  * scale = in / out (double); support = 2 * max(scale, 1); for output i: center = scale (i + 0.5),
    xmin = max(int(center - support + 0.5), 0), xsize = min(int(center + support + 0.5), in) - xmin,
    w_j = cubic((j + xmin - center + 0.5) / max(scale, 1)) with a = -0.5, normalised to sum 1 (double)
  * weights -> int16 with the largest precision p < 22 such that round(w_max 2^(p+1)) < 2^15; w < 0 rounds half down
  * horizontal pass over W first, then vertical over H; each pass: (2^(p-1) + sum w_j px_j) >> p, clamped to [0, 255]

Plain numpy, integer arithmetic in int64: exact by construction."""
from __future__ import annotations

from functools import lru_cache
from typing import Sequence, Tuple

import numpy as np


def _cubic_aa(x: np.ndarray) -> np.ndarray:
    a = -0.5
    x = np.abs(x)
    w1 = ((a + 2.0) * x - (a + 3.0)) * x * x + 1.0
    w2 = ((a * x - 5.0 * a) * x + 8.0 * a) * x - 4.0 * a
    return np.where(x < 1.0, w1, np.where(x < 2.0, w2, 0.0))


@lru_cache(maxsize=256)
def aa_weights_int16(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, np.ndarray, int]:
    """(xmin [out], xsize [out], weights int16 [out, max_taps], precision) of one separable pass."""
    scale = float(in_size) / float(out_size)
    support = 2.0 * scale if scale >= 1.0 else 2.0
    max_taps = int(np.ceil(support)) * 2 + 1
    invscale = 1.0 / scale if scale >= 1.0 else 1.0
    xmin = np.zeros(out_size, np.int64)
    xsize = np.zeros(out_size, np.int64)
    wts = np.zeros((out_size, max_taps), np.float64)
    wmax = 0.0
    for i in range(out_size):
        center = scale * (i + 0.5)
        lo = max(int(center - support + 0.5), 0)
        n = min(int(center + support + 0.5), in_size) - lo
        n = min(max(n, 0), max_taps)
        j = np.arange(n, dtype=np.float64)
        w = _cubic_aa((j + lo - center + 0.5) * invscale)
        tot = 0.0  # sequential sum as in ATen (numpy's pairwise sum can differ in the last bit)
        for v in w:
            tot += v
        if tot != 0.0:
            w = w / tot
        xmin[i], xsize[i] = lo, n
        wts[i, :n] = w
        if n:
            wmax = max(wmax, float(w.max()))
    prec = 0
    while prec < 22:
        if int(0.5 + wmax * (1 << (prec + 1))) >= (1 << 15):
            break
        prec += 1
    scaled = wts * float(1 << prec)
    w16 = np.where(wts < 0, np.trunc(-0.5 + scaled), np.trunc(0.5 + scaled)).astype(np.int16)
    return xmin, xsize, w16, prec


def _pass_u8(img: np.ndarray, axis: int, out_size: int) -> np.ndarray:
    """One separable pass along `axis` (0 = H, 1 = W) of an HWC uint8 image."""
    in_size = img.shape[axis]
    if in_size == out_size:
        return img
    xmin, xsize, w16, prec = aa_weights_int16(in_size, out_size)
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((out_size,) + src.shape[1:], np.int64)
    init = 1 << (prec - 1)
    for i in range(out_size):
        n = int(xsize[i])
        acc = np.tensordot(w16[i, :n].astype(np.int64), src[xmin[i]:xmin[i] + n], axes=(0, 0)) + init
        out[i] = acc >> prec
    return np.moveaxis(np.clip(out, 0, 255).astype(np.uint8), 0, axis)


def resize_bicubic_aa_u8(img_hwc: np.ndarray, out_h: int, out_w: int) -> np.ndarray:
    """== tvF.resize(uint8 CHW, [out_h, out_w], BICUBIC, antialias=True) on CPU, HWC in / HWC out. W pass, then H."""
    assert img_hwc.dtype == np.uint8 and img_hwc.ndim == 3
    return _pass_u8(_pass_u8(img_hwc, 1, out_w), 0, out_h)


def shortest_edge_size(h: int, w: int, short: int) -> Tuple[int, int]:
    """hf:image_transforms.get_resize_output_image_size(default_to_square=False): the short side becomes `short`, the
    long one int(short * long / short_side)."""
    if h <= w:
        return short, int(short * w / h)
    return int(short * h / w), short


def center_crop_box(h: int, w: int, ch: int, cw: int) -> Tuple[int, int]:
    """hf:image_processing_backends.py:357-358 (TorchvisionBackend.center_crop): top = int((h - ch) / 2), truncating —
    NOT torchvision's own rounding center_crop. The shortest-edge resize in front guarantees h >= ch and w >= cw."""
    return int((h - ch) / 2.0), int((w - cw) / 2.0)


def preprocess_u8(img_hwc: np.ndarray, mode: str, size: int) -> np.ndarray:
    """uint8 HWC image of any size -> uint8 [size, size, 3] exactly as the tower's processor produces it before the fp32
    rescale + normalize. mode 'square' (SigLIP, BLIP: resize to size x size) or 'shortest_crop' (CLIP)."""
    h, w = img_hwc.shape[:2]
    if mode == "square":
        return resize_bicubic_aa_u8(img_hwc, size, size)
    if mode != "shortest_crop":
        raise ValueError(mode)
    nh, nw = shortest_edge_size(h, w, size)
    r = resize_bicubic_aa_u8(img_hwc, nh, nw)
    top, left = center_crop_box(nh, nw, size, size)
    return np.ascontiguousarray(r[top:top + size, left:left + size])


def rescale_normalize(u8_nhwc: np.ndarray, mean: Sequence[float], std: Sequence[float]) -> np.ndarray:
    """fp32 NCHW pixel_values; hf TorchvisionBackend.rescale_and_normalize fuses the two steps as
    (x - 255 mean) / (255 std) in fp32."""
    x = u8_nhwc.astype(np.float32)
    m = (np.asarray(mean, np.float64) * 255.0).astype(np.float32)
    s = (np.asarray(std, np.float64) * 255.0).astype(np.float32)
    return np.ascontiguousarray(np.moveaxis((x - m) / s, -1, 1))
