"""Kernel-level entry points of the C-ABI on torch tensors (used by the parity tests, the profiler and the probes).

Every function launches on the current CUDA stream of the tensor's device and raises ``VlmfwError`` on failure.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import torch

from . import _lib
from .specs import ACT_NONE

EPI_BIAS_BF16, EPI_BIAS_RESID_F32, EPI_BIAS_RESID_16 = 0, 1, 2


def _st(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _need(t: torch.Tensor, dtype, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: vlmfw_b200 has no CPU fallback")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    if not (t.is_contiguous() or (t.dim() == 2 and t.stride(1) == 1 and t.stride(0) >= t.shape[1])):
        raise ValueError(f"{name} must be contiguous (2-D tensors may have a padded row pitch)")


def _dt(t: torch.Tensor, name: str) -> int:
    if t.dtype == torch.bfloat16:
        return _lib.DT_BF16
    if t.dtype == torch.float16:
        return _lib.DT_F16
    raise TypeError(f"{name} must be bfloat16 or float16, got {t.dtype}")


def pack_16(src: torch.Tensor, ld: Optional[int] = None, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """fp32 [rows, cols] -> bf16 / fp16 [rows, ld] (zero padded), round-to-nearest-even like torch's ``.to(dtype)``."""
    _need(src, torch.float32, "src")
    rows, cols = src.shape
    ld = cols if ld is None else ld
    dst = torch.empty(rows, ld, dtype=dtype, device=src.device)
    with torch.cuda.device(src.device):
        _lib.check(_lib.load().vlmfw_pack_16(src.data_ptr(), rows, cols, dst.data_ptr(), ld, _dt(dst, "dst"), _st(src)))
    return dst


def gemm_tn(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor] = None, *, act: int = ACT_NONE,
              resid: Optional[torch.Tensor] = None, resid_mod: int = 0, out: Optional[torch.Tensor] = None,
              grp: int = 0, grp_out: int = 0, grp_off: int = 0, out_rows: Optional[int] = None,
              block_n: int = 0) -> torch.Tensor:
    """``epilogue(a[M,K] @ w[N,K]^T)`` on tcgen05; a and w share one 16-bit dtype (bf16 or fp16). Without ``resid``:
    ``act(acc + bias)`` in that dtype; with ``resid`` (fp32 [*, N]): fp32 ``acc + bias + resid[row % resid_mod]``."""
    _need(a, a.dtype, "a")
    _need(w, a.dtype, "w")
    dt = _dt(a, "a")
    m, k = a.shape
    n, k2 = w.shape
    if k != k2:
        raise ValueError(f"K mismatch: a {tuple(a.shape)} vs w {tuple(w.shape)}")
    if bias is not None:
        _need(bias, torch.float32, "bias")
    rows = m if out_rows is None else out_rows
    if resid is not None:
        _need(resid, torch.float32, "resid")
        epi = EPI_BIAS_RESID_F32
        if out is None:
            out = torch.empty(rows, n, dtype=torch.float32, device=a.device)
        _need(out, torch.float32, "out")
    else:
        epi = EPI_BIAS_BF16
        if out is None:
            out = torch.empty(rows, n, dtype=a.dtype, device=a.device)
        _need(out, a.dtype, "out")
    with torch.cuda.device(a.device):
        _lib.check(_lib.load().vlmfw_gemm_tn(
            a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), dt, m, n, k,
            None if bias is None else bias.data_ptr(), epi, act, out.data_ptr(), out.stride(0),
            None if resid is None else resid.data_ptr(), 0 if resid is None else resid.stride(0), resid_mod,
            grp, grp_out, grp_off, block_n, _st(a)))
    return out


def gemm_stats_parts(n: int, block_n: int = 0) -> int:
    return int(_lib.load().vlmfw_gemm_stats_parts(n, block_n))


def row_stats_cast(x: torch.Tensor, dtype: torch.dtype = torch.float16, parts: int = 1):
    """fp32 rows -> (16-bit cast, stats [rows, parts, 2] with (sum, sum of squares) in slot 0): head of the folded-LN chain."""
    _need(x, torch.float32, "x")
    rows, dim = x.shape
    out = torch.empty(rows, dim, dtype=dtype, device=x.device)
    stats = torch.empty(rows, parts, 2, dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().vlmfw_row_stats_cast(x.data_ptr(), rows, dim, out.data_ptr(), _dt(out, "out"),
                                                   stats.data_ptr(), parts, _st(x)))
    return out, stats


def ln_fold_weight(w: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, bias: Optional[torch.Tensor] = None,
                   dtype: torch.dtype = torch.float16):
    """(W, LN gamma / beta, bias) -> (W' = round16(gamma . W with centred rows), c2 = W beta + bias)."""
    _need(w, torch.float32, "w")
    n, k = w.shape
    w16 = torch.empty(n, k, dtype=dtype, device=w.device)
    c2 = torch.empty(n, dtype=torch.float32, device=w.device)
    with torch.cuda.device(w.device):
        _lib.check(_lib.load().vlmfw_ln_fold_weight(w.data_ptr(), n, k, gamma.data_ptr(), beta.data_ptr(),
                                                   None if bias is None else bias.data_ptr(), w16.data_ptr(), k,
                                                   _dt(w16, "w16"), c2.data_ptr(), _st(w)))
    return w16, c2


def padded_rows(rows: int, cols: int, dtype: torch.dtype, device, multiple: int = 64) -> torch.Tensor:
    """[rows, cols] view whose row pitch is ``cols`` rounded up to ``multiple`` elements — whole 128-byte lines for
    16-bit data, what the tower uses for its MLP intermediate: TMA box rows then never straddle two lines."""
    ld = (cols + multiple - 1) // multiple * multiple
    return torch.empty(rows, ld, dtype=dtype, device=device)[:, :cols]


def gemm_tn_ln(a16: torch.Tensor, w16: torch.Tensor, c2: torch.Tensor, stats: torch.Tensor,
               ln_dim: int, eps: float, *, act: int = ACT_NONE, block_n: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Consumer side of the folded LayerNorm: ``act(rstd (a16 @ w16^T) + c2)`` in a16's dtype."""
    _need(a16, a16.dtype, "a16")
    _need(w16, a16.dtype, "w16")
    _need(stats, torch.float32, "stats")
    m, k = a16.shape
    n = w16.shape[0]
    if out is None:
        out = torch.empty(m, n, dtype=a16.dtype, device=a16.device)
    _need(out, a16.dtype, "out")
    d = _lib.GemmDesc(a=a16.data_ptr(), lda=a16.stride(0), w=w16.data_ptr(), ldw=w16.stride(0), dtype=_dt(a16, "a16"),
                      m=m, n=n, k=k, bias=c2.data_ptr(), epi=EPI_BIAS_BF16, act=act, out=out.data_ptr(),
                      ldo=out.stride(0), resid=None, ldr=0,
                      block_n=block_n, ln_stats=stats.data_ptr(), ln_np=stats.shape[1],
                      ln_dim=ln_dim, ln_eps=eps, h16_out=None, ld_h16=0, stats_out=None)
    with torch.cuda.device(a16.device):
        _lib.check(_lib.load().vlmfw_gemm_tn_ex(C.byref(d), _st(a16)))
    return out


def gemm_tn_resid_stats(a16: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor], h: torch.Tensor, *,
                        block_n: int = 0):
    """Producer side: ``h += a16 @ w16^T + bias`` in place, returning (16-bit cast of the new h, stats [M, parts, 2])."""
    _need(a16, a16.dtype, "a16")
    _need(w16, a16.dtype, "w16")
    _need(h, torch.float32, "h")
    m, k = a16.shape
    n = w16.shape[0]
    parts = gemm_stats_parts(n, block_n)
    h16 = torch.empty(m, n, dtype=a16.dtype, device=a16.device)
    stats = torch.full((m, parts, 2), float("nan"), dtype=torch.float32, device=a16.device)
    d = _lib.GemmDesc(a=a16.data_ptr(), lda=a16.stride(0), w=w16.data_ptr(), ldw=w16.stride(0), dtype=_dt(a16, "a16"),
                      m=m, n=n, k=k, bias=None if bias is None else bias.data_ptr(), epi=EPI_BIAS_RESID_F32, act=ACT_NONE,
                      out=h.data_ptr(), ldo=n, resid=h.data_ptr(), ldr=n, block_n=block_n, ln_stats=None,
                      ln_np=0, ln_dim=0, ln_eps=0.0, h16_out=h16.data_ptr(), ld_h16=n, stats_out=stats.data_ptr())
    with torch.cuda.device(a16.device):
        _lib.check(_lib.load().vlmfw_gemm_tn_ex(C.byref(d), _st(a16)))
    return h16, stats


def gemm_tn_resid16(a16: torch.Tensor, w16: torch.Tensor, bias: Optional[torch.Tensor], h16: torch.Tensor, *,
                    block_n: int = 0) -> torch.Tensor:
    """16-bit residual stream: ``h16 = round16(h16 + a16 @ w16^T + bias)`` in place; returns the row statistics
    ``[M, parts, 2]`` (sum, sum of squares of the ROUNDED values per column slice) for the LayerNorm folded into the next GEMM."""
    _need(a16, a16.dtype, "a16")
    _need(w16, a16.dtype, "w16")
    _need(h16, a16.dtype, "h16")
    m, k = a16.shape
    n = w16.shape[0]
    parts = gemm_stats_parts(n, block_n)
    stats = torch.full((m, parts, 2), float("nan"), dtype=torch.float32, device=a16.device)
    d = _lib.GemmDesc(a=a16.data_ptr(), lda=a16.stride(0), w=w16.data_ptr(), ldw=w16.stride(0), dtype=_dt(a16, "a16"),
                      m=m, n=n, k=k, bias=None if bias is None else bias.data_ptr(), epi=EPI_BIAS_RESID_16, act=ACT_NONE,
                      out=h16.data_ptr(), ldo=h16.stride(0), resid=h16.data_ptr(), ldr=h16.stride(0), block_n=block_n,
                      ln_stats=None, ln_np=0, ln_dim=0, ln_eps=0.0, h16_out=None, ld_h16=0, stats_out=stats.data_ptr())
    with torch.cuda.device(a16.device):
        _lib.check(_lib.load().vlmfw_gemm_tn_ex(C.byref(d), _st(a16)))
    return stats


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float,
              dtype: torch.dtype = torch.float16, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need(x, torch.float32, "x")
    _need(gamma, torch.float32, "gamma")
    _need(beta, torch.float32, "beta")
    rows, dim = x.shape
    if out is None:
        out = torch.empty(rows, dim, dtype=dtype, device=x.device)
    odt = _lib.DT_F32 if out.dtype == torch.float32 else _dt(out, "out")
    with torch.cuda.device(x.device):
        _lib.check(_lib.load().vlmfw_layernorm(x.data_ptr(), rows, dim, gamma.data_ptr(), beta.data_ptr(), eps,
                                               out.data_ptr(), odt, _st(x)))
    return out


def attention(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int) -> torch.Tensor:
    """qkv 16-bit [B*S, 3*H*dh] (q | k | v thirds) -> softmax(QK^T/sqrt(dh)) V in the same dtype [B*S, H*dh]."""
    _need(qkv, qkv.dtype, "qkv")
    dt = _dt(qkv, "qkv")
    d = heads * head_dim
    if tuple(qkv.shape) != (batch * tokens, 3 * d):
        raise ValueError(f"qkv must be [{batch * tokens}, {3 * d}], got {tuple(qkv.shape)}")
    out = torch.empty(batch * tokens, d, dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().vlmfw_attention(qkv.data_ptr(), dt, batch, tokens, heads, head_dim, out.data_ptr(),
                                               _st(qkv)))
    return out


def attention_cls(qkv: torch.Tensor, batch: int, tokens: int, heads: int, head_dim: int) -> torch.Tensor:
    """Attention of query row 0 of every image over all keys: qkv as for ``attention`` -> [B, H*dh] in the same dtype."""
    _need(qkv, qkv.dtype, "qkv")
    dt = _dt(qkv, "qkv")
    d = heads * head_dim
    if tuple(qkv.shape) != (batch * tokens, 3 * d):
        raise ValueError(f"qkv must be [{batch * tokens}, {3 * d}], got {tuple(qkv.shape)}")
    out = torch.empty(batch, d, dtype=qkv.dtype, device=qkv.device)
    with torch.cuda.device(qkv.device):
        _lib.check(_lib.load().vlmfw_attention_cls(qkv.data_ptr(), dt, batch, tokens, heads, head_dim, out.data_ptr(), d,
                                                   _st(qkv)))
    return out


def im2col(pixels: torch.Tensor, image_size: int, patch: int, kpad: int, mean: Sequence[float] = (0, 0, 0),
           std: Sequence[float] = (1, 1, 1), dtype: torch.dtype = torch.float16) -> torch.Tensor:
    if not pixels.is_cuda or not pixels.is_contiguous():
        raise RuntimeError("pixels must be a contiguous CUDA tensor")
    fmt = 0 if pixels.dtype == torch.uint8 else 1
    b = pixels.shape[0]
    g = image_size // patch
    out = torch.empty(b * g * g, kpad, dtype=dtype, device=pixels.device)
    m3 = (C.c_float * 3)(*mean)
    s3 = (C.c_float * 3)(*std)
    with torch.cuda.device(pixels.device):
        _lib.check(_lib.load().vlmfw_im2col(pixels.data_ptr(), fmt, b, image_size, patch, m3, s3, out.data_ptr(), kpad,
                                            _dt(out, "out"), _st(pixels)))
    return out


def pool(h: torch.Tensor, mode: int, gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None,
         eps: float = 1e-6) -> torch.Tensor:
    """h fp32 [B, S, D] -> pooled fp32 [B, D] (VLMFW_POOL_* mode)."""
    _need(h, torch.float32, "h")
    b, s, d = h.shape
    out = torch.empty(b, d, dtype=torch.float32, device=h.device)
    with torch.cuda.device(h.device):
        _lib.check(_lib.load().vlmfw_pool(h.data_ptr(), b, s, d, mode, None if gamma is None else gamma.data_ptr(),
                                          None if beta is None else beta.data_ptr(), eps, out.data_ptr(), _st(h)))
    return out
