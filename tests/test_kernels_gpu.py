"""Kernel-level parity (-m gpu): every sm_100a kernel, called through the C-ABI, against a plain fp32 torch
restatement of the same op on the same seeded tensors. Tolerances are written next to each check."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from vlmfw_b200 import ops, specs  # noqa: E402


def _rand(shape, seed, scale=1.0, device="cuda"):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(shape, generator=g) * scale).to(device)


def _act_ref(x, act):
    if act == specs.ACT_GELU_TANH:
        return torch.nn.functional.gelu(x, approximate="tanh")
    if act == specs.ACT_QUICK_GELU:
        return x * torch.sigmoid(1.702 * x)
    if act == specs.ACT_GELU_ERF:
        return torch.nn.functional.gelu(x)
    return x


# (M, N, K): tile-exact, ragged M, ragged N (4304 = 16*256+208), K % 64 = 16 (4304, 592), tiny; the CTA-pair kernel runs a
# ragged last N tile as a narrower MMA (width rounded up to 16): 1152 -> 128, 4304 -> 208, 1160 -> 144 (8 zero columns),
# 1288 -> 16 (the narrowest)
GEMM_SHAPES = [(256, 256, 128), (128, 128, 64), (384, 1152, 1152), (1000, 3456, 1152), (520, 4304, 1152),
               (520, 1152, 4304), (777, 1152, 592), (64, 144, 144), (130, 432, 144), (40, 176, 384), (600, 1160, 320),
               (300, 1288, 192)]


# kernel forms: single CTA (128-row tiles) and CTA pair (tcgen05 cta_group::2, 256-row tiles); include/vlmfw_b200.h
FORMS = {"auto": 0, "1cta": 0x1000, "2cta": 0x2000, "1cta-generic-epi": 0x5000, "2cta-generic-epi": 0x6000}


@pytest.mark.parametrize("m,n,k", GEMM_SHAPES)
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
@pytest.mark.parametrize("form", list(FORMS))
def test_gemm_bias_bf16(cuda_device, m, n, k, block_n, form):
    a = _rand((m, k), 1).to(torch.bfloat16)
    w = _rand((n, k), 2, 1 / math.sqrt(k)).to(torch.bfloat16)
    bias = _rand((n,), 3, 0.5)
    out = ops.gemm_tn(a, w, bias, block_n=block_n | FORMS[form])
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    # fp32 accumulate of exact bf16 products; only the final bf16 rounding (rel 2^-8) differs
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    # and it must be the correctly rounded value almost everywhere
    assert (out.float() - ref.to(torch.bfloat16).float()).abs().max() <= 0.04


@pytest.mark.parametrize("act", [specs.ACT_GELU_TANH, specs.ACT_QUICK_GELU, specs.ACT_GELU_ERF])
@pytest.mark.parametrize("form", ["1cta", "2cta"])
def test_gemm_activation(cuda_device, act, form):
    m, n, k = 300, 4304, 1152
    a = _rand((m, k), 4).to(torch.bfloat16)
    w = _rand((n, k), 5, 2 / math.sqrt(k)).to(torch.bfloat16)
    bias = _rand((n,), 6, 0.5)
    out = ops.gemm_tn(a, w, bias, act=act, block_n=FORMS[form])
    ref = _act_ref(a.float() @ w.float().t() + bias, act)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("m,n,k", [(384, 1152, 1152), (777, 1152, 4304), (130, 144, 304), (64, 1408, 6144),
                                   (40000, 1152, 1152)])
@pytest.mark.parametrize("form", ["1cta", "2cta", "1cta-generic-epi", "2cta-generic-epi"])
def test_gemm_residual_f32_inplace(cuda_device, m, n, k, form):
    a = _rand((m, k), 7).to(torch.bfloat16)
    w = _rand((n, k), 8, 1 / math.sqrt(k)).to(torch.bfloat16)
    bias = _rand((n,), 9, 0.5)
    h = _rand((m, n), 10)
    ref = a.float() @ w.float().t() + bias + h
    out = ops.gemm_tn(a, w, bias, resid=h, out=h, block_n=FORMS[form])  # in place, as the tower does for the residual stream
    assert out.data_ptr() == h.data_ptr()
    # fp32 end to end: only accumulation-order error
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4)


@pytest.mark.parametrize("form", ["1cta", "2cta"])
def test_gemm_patch_embed_row_remap(cuda_device, form):
    # rows of image b land at b*S + 1 + p (class-token slot left free), residual = pos[p] broadcast over images
    b, p, s, n, k = 3, 25, 26, 128, 592
    a = _rand((b * p, k), 11).to(torch.bfloat16)
    w = _rand((n, k), 12, 1 / math.sqrt(k)).to(torch.bfloat16)
    pos = _rand((p, n), 13)
    out = torch.full((b * s, n), -7.0, device="cuda")
    ops.gemm_tn(a, w, None, resid=pos, resid_mod=p, out=out, grp=p, grp_out=s, grp_off=1, block_n=FORMS[form])
    ref = (a.float() @ w.float().t()).view(b, p, n) + pos
    got = out.view(b, s, n)
    torch.testing.assert_close(got[:, 1:], ref, rtol=1e-4, atol=2e-4)
    assert (got[:, 0] == -7.0).all()


@pytest.mark.parametrize("form", ["1cta", "2cta", "2cta-generic-epi"])
@pytest.mark.parametrize("m,n,k,mod", [(700, 432, 304, 64), (1024, 1152, 640, 256), (300, 144, 64, 32)])
def test_gemm_broadcast_residual(cuda_device, form, m, n, k, mod):
    """out = A W^T + bias + resid[row % mod] — the position-embedding add of the patch GEMM. With mod a multiple of 32 it
    runs through the TMA residual epilogue (the residual box is fetched at row % mod), otherwise / when forced through the
    generic one; both must agree with the reference."""
    a = _rand((m, k), 24).to(torch.float16)
    w = _rand((n, k), 25, 1 / math.sqrt(k)).to(torch.float16)
    bias = _rand((n,), 26, 0.5)
    pos = _rand((mod, n), 27)
    out = ops.gemm_tn(a, w, bias, resid=pos, resid_mod=mod, block_n=FORMS[form])
    ref = a.float() @ w.float().t() + bias + pos.repeat((m + mod - 1) // mod, 1)[:m]
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


@pytest.mark.parametrize("form", ["1cta", "2cta"])
def test_gemm_residual_f32_out_of_place(cuda_device, form):
    # resid and out are different buffers with different leading dimensions (two TMA maps)
    m, n, k = 700, 432, 304
    a = _rand((m, k), 14).to(torch.float16)
    w = _rand((n, k), 15, 1 / math.sqrt(k)).to(torch.float16)
    bias = _rand((n,), 16, 0.5)
    big = _rand((m, n + 16), 17)
    resid = big[:, :n]
    out = torch.zeros(m, n, device="cuda")
    with torch.cuda.device(0):
        from vlmfw_b200 import _lib
        _lib.check(_lib.load().vlmfw_gemm_tn(a.data_ptr(), k, w.data_ptr(), k, _lib.DT_F16, m, n, k, bias.data_ptr(),
                                            ops.EPI_BIAS_RESID_F32, 0, out.data_ptr(), n, resid.data_ptr(), n + 16, 0,
                                            0, 0, 0, FORMS[form], torch.cuda.current_stream().cuda_stream))
    ref = a.float() @ w.float().t() + bias + resid
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4)


# ------------------------------------------------------------------ LayerNorm folded around the GEMMs
@pytest.mark.parametrize("rows,dim,parts", [(37, 1152, 10), (4096, 1024, 8), (513, 1408, 12), (9, 144, 2)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_row_stats_cast(cuda_device, rows, dim, parts, dtype):
    x = _rand((rows, dim), 40, 3.0) + 1.5
    out, stats = ops.row_stats_cast(x, dtype=dtype, parts=parts)
    assert torch.equal(out, x.to(dtype))  # round-to-nearest-even cast, bit exact
    torch.testing.assert_close(stats[:, 0, 0], x.sum(1), rtol=1e-5, atol=1e-3)
    torch.testing.assert_close(stats[:, 0, 1], (x.double() ** 2).sum(1).float(), rtol=1e-5, atol=1e-3)
    assert (stats[:, 1:] == 0).all()


@pytest.mark.parametrize("n,k", [(3456, 1152), (4304, 1152), (432, 144), (4096, 1024)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_ln_fold_weight(cuda_device, n, k, dtype):
    w = _rand((n, k), 41, 1 / math.sqrt(k))
    gamma, beta, bias = 1 + _rand((k,), 42, 0.2), _rand((k,), 43, 0.2), _rand((n,), 44, 0.5)
    w16, c2 = ops.ln_fold_weight(w, gamma, beta, bias, dtype=dtype)
    gw = w.double() * gamma.double()
    want = (gw - gw.mean(1, keepdim=True)).float()
    # round-to-nearest of the centred row; the fp32 FMA of the kernel may differ from the fp64 restatement by one ulp
    # of the 16-bit type on a rounding boundary, never more
    ulp = 2.0 ** (-10 if dtype == torch.float16 else -7)
    assert ((w16.float() - want).abs() <= ulp * want.abs().clamp_min(2.0 ** -14)).all()
    assert (w16.float() == want.to(dtype).float()).float().mean() > 0.999
    assert w16.double().sum(1).abs().max() < 0.05  # rows are centred up to the 16-bit rounding residue
    torch.testing.assert_close(c2, (w.double() @ beta.double() + bias.double()).float(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("m,n,k", [(384, 1152, 1152), (1000, 3456, 1152), (520, 4304, 1152), (130, 432, 144),
                                   (40, 176, 384), (20000, 1152, 1152)])
@pytest.mark.parametrize("form", ["1cta", "2cta"])
@pytest.mark.parametrize("act", [specs.ACT_NONE, specs.ACT_GELU_TANH])
def test_gemm_folded_layernorm_consumer(cuda_device, m, n, k, form, act):
    """LN(x) W^T + b through (row_stats_cast -> ln_fold_weight -> GEMM with the LN epilogue) == torch's layer_norm + linear."""
    x = _rand((m, k), 45, 2.0) + 0.7  # rows with a non-zero mean: the mu c1 cancellation is exercised
    w = _rand((n, k), 46, 1 / math.sqrt(k))
    gamma, beta, bias = 1 + _rand((k,), 47, 0.2), _rand((k,), 48, 0.2), _rand((n,), 49, 0.5)
    parts = 5
    x16, stats = ops.row_stats_cast(x, dtype=torch.float16, parts=parts)
    w16, c2 = ops.ln_fold_weight(w, gamma, beta, bias, dtype=torch.float16)
    out = ops.gemm_tn_ln(x16, w16, c2, stats, k, 1e-6, act=act, block_n=FORMS[form])
    ref = _act_ref(torch.nn.functional.layer_norm(x, (k,), gamma, beta, 1e-6) @ w.t() + bias, act)
    # same error budget as the unfused path (fp16 operands, fp32 accumulate, fp16 output)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)
    # and against the unfused kernels on the same inputs: the fold must not cost accuracy
    xn = ops.layernorm(x, gamma, beta, 1e-6, dtype=torch.float16)
    unfused = ops.gemm_tn(xn, w.to(torch.float16), bias, act=act, block_n=FORMS[form])
    err_fused = (out.float() - ref).abs().max().item()
    err_unfused = (unfused.float() - ref).abs().max().item()
    assert err_fused <= 2.0 * err_unfused + 2e-3, (err_fused, err_unfused)


@pytest.mark.parametrize("m,n,k", [(384, 1152, 1152), (777, 1152, 4304), (130, 144, 304), (64, 1408, 6144),
                                   (20000, 1024, 1024), (600, 1160, 320), (300, 1288, 192)])
@pytest.mark.parametrize("form", ["1cta", "2cta"])
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_residual_producer_stats_and_cast(cuda_device, m, n, k, form, block_n, dtype):
    a = _rand((m, k), 50).to(dtype)
    w = _rand((n, k), 51, 1 / math.sqrt(k)).to(dtype)
    bias = _rand((n,), 52, 0.5)
    h = _rand((m, n), 53, 2.0) + 0.5
    ref = a.float() @ w.float().t() + bias + h
    h16, stats = ops.gemm_tn_resid_stats(a, w, bias, h, block_n=block_n | FORMS[form])
    torch.testing.assert_close(h, ref, rtol=1e-4, atol=2e-4)
    assert torch.equal(h16, h.to(dtype))  # the cast of exactly what was stored
    assert not torch.isnan(stats).any()  # every partial slot written
    tot = stats.sum(1)
    torch.testing.assert_close(tot[:, 0], h.sum(1), rtol=1e-4, atol=2e-2)
    torch.testing.assert_close(tot[:, 1], (h * h).sum(1), rtol=1e-4, atol=2e-2)


@pytest.mark.parametrize("m,n,k", [(384, 1152, 1152), (777, 1152, 4304), (130, 144, 304), (64, 1408, 6144),
                                   (20000, 1024, 1024), (600, 1160, 320), (300, 1288, 192)])
@pytest.mark.parametrize("form", ["1cta", "2cta"])
@pytest.mark.parametrize("block_n", [0, 128, 192, 256])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_gemm_residual_16bit_stream(cuda_device, m, n, k, form, block_n, dtype):
    """EPI_BIAS_RESID_16: h16 = round16(h16 + A W^T + bias) in place + row statistics of the rounded values."""
    a = _rand((m, k), 40).to(dtype)
    w = _rand((n, k), 41, 1 / math.sqrt(k)).to(dtype)
    bias = _rand((n,), 42, 0.5)
    h0 = (_rand((m, n), 43, 2.0) + 0.3).to(dtype)
    h = h0.clone()
    stats = ops.gemm_tn_resid16(a, w, bias, h, block_n=block_n | FORMS[form])
    exact = h0.float() + a.float() @ w.float().t() + bias
    ulp = 2.0 ** (-10 if dtype == torch.float16 else -7)  # one unit in the last place, relative
    err = (h.float() - exact).abs()
    assert (err <= 0.51 * ulp * exact.abs().clamp_min(2.0 ** -14) + 1e-4).all(), float(err.max())  # correctly rounded +- accumulation order
    tot = stats.sum(dim=1)
    torch.testing.assert_close(tot[:, 0], h.float().sum(1), rtol=1e-4, atol=2e-2)
    torch.testing.assert_close(tot[:, 1], (h.float() ** 2).sum(1), rtol=1e-4, atol=2e-2)


def test_gemm_folded_chain_matches_layernorm_linear(cuda_device):
    """producer -> consumer exactly as the tower chains them: h += a W_o^T + b_o; y = LN(h) W_1^T + b_1."""
    m, d, i = 3000, 1152, 4304
    a = _rand((m, d), 54).to(torch.float16)
    wo = _rand((d, d), 55, 1 / math.sqrt(d)).to(torch.float16)
    bo = _rand((d,), 56, 0.5)
    h = _rand((m, d), 57, 2.0)
    w1 = _rand((i, d), 58, 1 / math.sqrt(d))
    gamma, beta, b1 = 1 + _rand((d,), 59, 0.2), _rand((d,), 60, 0.2), _rand((i,), 61, 0.5)
    href = a.float() @ wo.float().t() + bo + h
    ref = torch.nn.functional.gelu(torch.nn.functional.layer_norm(href, (d,), gamma, beta, 1e-6) @ w1.t() + b1, approximate="tanh")
    h16, stats = ops.gemm_tn_resid_stats(a, wo, bo, h)
    w16, c2 = ops.ln_fold_weight(w1, gamma, beta, b1)
    out = ops.gemm_tn_ln(h16, w16, c2, stats, d, 1e-6, act=specs.ACT_GELU_TANH)
    torch.testing.assert_close(out.float(), ref, rtol=1e-2, atol=1e-2)


def test_gemm_rejects_bad_arguments(cuda_device):
    from vlmfw_b200._lib import VlmfwError

    a = torch.zeros(8, 20, dtype=torch.bfloat16, device="cuda")  # K not a multiple of 8
    w = torch.zeros(8, 20, dtype=torch.bfloat16, device="cuda")
    with pytest.raises(VlmfwError):
        ops.gemm_tn(a, w)


@pytest.mark.parametrize("rows,dim", [(1, 144), (37, 1152), (4096, 1024), (513, 1408), (9, 176), (5, 128)])
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16, torch.float32])
def test_layernorm(cuda_device, rows, dim, dtype):
    x = _rand((rows, dim), 20, 3.0) + 1.5
    g = 1 + _rand((dim,), 21, 0.1)
    b = _rand((dim,), 22, 0.1)
    out = ops.layernorm(x, g, b, 1e-6, dtype=dtype)
    assert out.dtype == dtype
    ref = torch.nn.functional.layer_norm(x, (dim,), g, b, 1e-6)
    tol = {torch.float32: 2e-5, torch.float16: 1e-3, torch.bfloat16: 8e-3}[dtype]  # output rounding
    torch.testing.assert_close(out.float(), ref, rtol=tol, atol=tol)


def test_layernorm_inplace_f32(cuda_device):
    x = _rand((77, 1024), 23, 2.0)
    g, b = 1 + _rand((1024,), 24, 0.1), _rand((1024,), 25, 0.1)
    ref = torch.nn.functional.layer_norm(x, (1024,), g, b, 1e-5)
    ops.layernorm(x, g, b, 1e-5, out=x)
    torch.testing.assert_close(x, ref, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("batch,tokens,heads,dh", [(2, 256, 16, 72), (2, 577, 16, 64), (2, 257, 16, 88), (3, 16, 2, 72),
                                                  (3, 26, 2, 64), (3, 17, 2, 88), (1, 64, 1, 64), (1, 65, 3, 8),
                                                  # 64 k + 1 tokens: a last key block with a single key
                                                  (2, 193, 2, 64), (2, 385, 3, 72), (3, 321, 2, 88), (2, 129, 2, 72)])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_attention(cuda_device, batch, tokens, heads, dh, dtype):
    d = heads * dh
    qkv = _rand((batch * tokens, 3 * d), 30, 1.5).to(dtype)
    out = ops.attention(qkv, batch, tokens, heads, dh)
    q, k, v = (t.view(batch, tokens, heads, dh).transpose(1, 2) for t in qkv.float().split(d, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    ref = (att @ v).transpose(1, 2).reshape(batch * tokens, d)
    # P is rounded to 16 bits before P·V and the output is 16-bit: abs error ~ 2^-8 (bf16) / 2^-11 (fp16) * |v|
    tol = 2e-2 if dtype == torch.bfloat16 else 3e-3
    assert out.dtype == dtype
    torch.testing.assert_close(out.float(), ref, rtol=tol, atol=tol)


def _attn_ref(qkv, batch, tokens, heads, dh):
    d = heads * dh
    q, k, v = (t.view(batch, tokens, heads, dh).transpose(1, 2) for t in qkv.float().split(d, dim=-1))
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(dh), dim=-1)
    return (att @ v).transpose(1, 2).reshape(batch * tokens, d)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("batch,tokens,heads,dh", [(5, 257, 16, 88), (3, 577, 16, 64), (4, 17, 2, 72), (2, 1030, 3, 8),
                                                   (2, 50, 4, 128)])
def test_attention_cls(cuda_device, batch, tokens, heads, dh, dtype):
    """One-query attention (the class-token row of the last layer, hf sdpa_attention.py:92 restricted to query 0): fp32
    probabilities, so only the 16-bit output rounding separates it from the fp32 reference."""
    d = heads * dh
    qkv = _rand((batch * tokens, 3 * d), 33, 1.5).to(dtype)
    out = ops.attention_cls(qkv, batch, tokens, heads, dh)
    ref = _attn_ref(qkv, batch, tokens, heads, dh).view(batch, tokens, d)[:, 0]
    tol = 8e-3 if dtype == torch.bfloat16 else 1e-3
    assert out.dtype == dtype and tuple(out.shape) == (batch, d)
    torch.testing.assert_close(out.float(), ref, rtol=tol, atol=tol)
    # and it agrees with row 0 of the all-queries kernel wherever that kernel covers the shape
    if dh <= 96:
        full = ops.attention(qkv, batch, tokens, heads, dh).view(batch, tokens, d)[:, 0]
        torch.testing.assert_close(out.float(), full.float(), rtol=4 * tol, atol=4 * tol)


@pytest.mark.parametrize("batch,tokens,heads,dh", [(24, 256, 16, 72), (7, 577, 16, 64), (9, 257, 16, 88)])
def test_attention_persistent_many_items(cuda_device, batch, tokens, heads, dh):
    """More work items than SMs: every CTA of the persistent tcgen05 kernel loops over several (image, head, query-pair)
    items, so barrier phases, TMEM reuse and the K/V ring wrap around many times."""
    qkv = _rand((batch * tokens, 3 * heads * dh), 31, 1.2).to(torch.float16)
    out = ops.attention(qkv, batch, tokens, heads, dh)
    torch.testing.assert_close(out.float(), _attn_ref(qkv, batch, tokens, heads, dh), rtol=3e-3, atol=3e-3)


@pytest.mark.parametrize("tokens,dh", [(256, 72), (577, 64), (257, 88)])
def test_attention_running_max_rescale(cuda_device, tokens, dh):
    """Scores that grow along the key axis force the running max up by more than 2^8 between key blocks, i.e. the path
    that rescales the O accumulator inside TMEM."""
    batch, heads = 2, 4
    d = heads * dh
    qkv = _rand((batch * tokens, 3 * d), 32, 1.0)
    ramp = torch.linspace(0.2, 6.0, tokens, device="cuda").repeat(batch)[:, None]
    qkv[:, d:2 * d] *= ramp  # keys of later tokens are much larger
    qkv[:, :d] *= 2.0
    qkv = qkv.to(torch.float16)
    out = ops.attention(qkv, batch, tokens, heads, dh)
    ref = _attn_ref(qkv, batch, tokens, heads, dh)
    assert torch.isfinite(out.float()).all()
    torch.testing.assert_close(out.float(), ref, rtol=5e-3, atol=5e-3)


@pytest.mark.parametrize("fmt", ["u8", "f32"])
@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_im2col(cuda_device, fmt, dtype):
    b, size, p = 3, 56, 14
    kreal, kpad = 3 * p * p, 592
    mean, std = (0.48, 0.45, 0.40), (0.27, 0.26, 0.28)
    g = torch.Generator().manual_seed(40)
    u8 = torch.randint(0, 256, (b, size, size, 3), generator=g, dtype=torch.uint8)
    pv = ((u8.float() / 255.0).permute(0, 3, 1, 2) - torch.tensor(mean).view(1, 3, 1, 1)) / torch.tensor(std).view(1, 3, 1, 1)
    if fmt == "u8":
        out = ops.im2col(u8.cuda(), size, p, kpad, mean, std, dtype=dtype)
    else:
        out = ops.im2col(pv.contiguous().cuda(), size, p, kpad, dtype=dtype)
    gs = size // p
    ref = pv.reshape(b, 3, gs, p, gs, p).permute(0, 2, 4, 1, 3, 5).reshape(b * gs * gs, kreal)
    tol = 8e-3 if dtype == torch.bfloat16 else 1e-3
    torch.testing.assert_close(out[:, :kreal].float().cpu(), ref, rtol=tol, atol=tol)
    assert (out[:, kreal:] == 0).all()


@pytest.mark.parametrize("mode", [specs.POOL_LN_MEAN, specs.POOL_MEAN, specs.POOL_CLS, specs.POOL_LN_CLS_LN])
@pytest.mark.parametrize("b,s,d", [(3, 256, 1152), (2, 577, 1024), (5, 257, 1408), (2, 17, 176)])
def test_pool(cuda_device, mode, b, s, d):
    h = _rand((b, s, d), 50, 2.0) + 0.3
    g, bt = 1 + _rand((d,), 51, 0.1), _rand((d,), 52, 0.1)
    out = ops.pool(h, mode, g, bt, 1e-6)
    ln = lambda t: torch.nn.functional.layer_norm(t, (d,), g, bt, 1e-6)  # noqa: E731
    ref = {specs.POOL_LN_MEAN: lambda: ln(h).mean(1), specs.POOL_MEAN: lambda: h.mean(1), specs.POOL_CLS: lambda: h[:, 0],
           specs.POOL_LN_CLS_LN: lambda: ln(ln(h)[:, 0])}[mode]()
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_pack_16(cuda_device, dtype):
    x = _rand((37, 588), 60)
    out = ops.pack_16(x, 592, dtype)
    assert torch.equal(out[:, :588], x.to(dtype))  # bit-exact RNE
    assert (out[:, 588:] == 0).all()


@pytest.mark.parametrize("m,n,k", [(384, 1152, 1152), (520, 4304, 1152), (777, 1152, 592), (130, 432, 144)])
def test_gemm_fp16_operands(cuda_device, m, n, k):
    """fp16 is the tower's default operand format (DESIGN.md numerics); same kernel, other instruction-descriptor bits."""
    a = _rand((m, k), 70).to(torch.float16)
    w = _rand((n, k), 71, 1 / math.sqrt(k)).to(torch.float16)
    bias = _rand((n,), 72, 0.5)
    h = _rand((m, n), 73)
    ref = a.float() @ w.float().t() + bias + h
    out = ops.gemm_tn(a, w, bias, resid=h.clone())
    torch.testing.assert_close(out, ref, rtol=1e-4, atol=2e-4)  # exact products, fp32 accumulate
    out16 = ops.gemm_tn(a, w, bias, act=specs.ACT_GELU_TANH)
    assert out16.dtype == torch.float16
    torch.testing.assert_close(out16.float(), _act_ref(ref - h, specs.ACT_GELU_TANH), rtol=2e-3, atol=2e-3)


def test_gemm_fp16_output_saturates(cuda_device):
    a = torch.full((128, 64), 200.0, dtype=torch.float16, device="cuda")
    w = torch.full((128, 64), 100.0, dtype=torch.float16, device="cuda")  # acc = 1.28e6 > 65504
    out = ops.gemm_tn(a, w)
    assert torch.isfinite(out.float()).all() and (out.float() == 65504.0).all()


def test_gemm_rejects_mixed_formats(cuda_device):
    """A and W of one tcgen05.mma kind::f16 must share a format (a bf16 x fp16 mix faults as an illegal instruction on
    B200), so the host refuses it."""
    a = torch.zeros(128, 64, dtype=torch.bfloat16, device="cuda")
    w = torch.zeros(128, 64, dtype=torch.float16, device="cuda")
    with pytest.raises(TypeError):
        ops.gemm_tn(a, w)
