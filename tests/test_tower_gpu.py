"""End-to-end parity (-m gpu): TowerEngine (CUDA, through the C-ABI) vs the fp32 CPU oracle on seeded synthetic images
and seeded non-degenerate weights. Tolerances are north_star's: pooled-feature cosine >= 0.999; the logit / argmax
checks live in test_probe_gpu.py."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tower_oracle as O  # noqa: E402
from vlmfw_b200 import specs, weights  # noqa: E402
from vlmfw_b200.engine import TowerEngine  # noqa: E402

COS_MIN = 0.999  # BASELINE.json north_star


def _run(spec, n_img, seed, smooth, fmt, compute_dtype="f16"):
    sd = weights.seeded_hf_state_dict(spec, seed)
    eng = TowerEngine(spec, "cuda:0", compute_dtype=compute_dtype)
    eng.load_hf_state_dict(sd)
    imgs = weights.seeded_images_u8(n_img, spec.image_size, seed, smooth=smooth)
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    dev_in = imgs.cuda() if fmt == "u8" else pv.contiguous().cuda()
    got = eng.forward(dev_in)
    torch.cuda.synchronize()
    if eng.cls_tail:
        # class-token-pooled towers run their last layer for row 0 only: the features above come from that path; the
        # full residual stream the callers compare needs the whole last layer — and must pool to the same features
        eng.set_cls_tail(False)
        full = eng.forward(dev_in)
        torch.cuda.synchronize()
        assert (full - got).abs().max() <= 2e-3 * max(1.0, full.abs().max().item()), (full - got).abs().max()
        hid = eng.hidden_state(n_img).float().cpu()
        eng.set_cls_tail(True)
    else:
        hid = eng.hidden_state(n_img).float().cpu()
    can = weights.hf_to_canonical(spec, sd)
    ref = O.tower_features(spec, can, pv)
    ref_hid = O.last_hidden(spec, {k: v.float() for k, v in can.items()}, pv)
    return got.cpu(), ref, hid, ref_hid, eng


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
@pytest.mark.parametrize("fmt", ["u8", "f32"])
@pytest.mark.parametrize("smooth", [False, True])
def test_tiny_tower_matches_oracle(cuda_device, name, fmt, smooth):
    spec = specs.SPECS[name]
    got, ref, hid, ref_hid, eng = _run(spec, 5, 3, smooth, fmt)
    assert got.shape == ref.shape == (5, spec.hidden)
    cos = torch.nn.functional.cosine_similarity(got, ref, dim=1)
    assert cos.min() >= COS_MIN, cos
    # residual stream itself (stricter than the pooled cosine): relative L2 error per image
    rel = (hid - ref_hid).flatten(1).norm(dim=1) / ref_hid.flatten(1).norm(dim=1)
    assert rel.max() < 5e-3, rel  # fp16 operands; bf16 mode sits near 1e-2
    assert (got - ref).abs().max() < 5e-2 * max(1.0, ref.abs().max().item())
    assert eng.last_launches > 0


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_bf16_compute_mode(cuda_device, name):
    """bf16 operands: the feature tolerance holds; the engine flags that the LOGIT tolerance is not guaranteed (the
    full-size measurement of that error is tests/test_parity_live.py::test_bf16_mode_is_fenced_from_the_logit_tolerance)."""
    with pytest.warns(RuntimeWarning, match="bf16"):
        got, ref, hid, ref_hid, eng = _run(specs.SPECS[name], 4, 7, True, "u8", compute_dtype="bf16")
    assert eng.compute_dtype == "bf16" and not eng.tolerance_ok
    assert torch.nn.functional.cosine_similarity(got, ref, dim=1).min() >= COS_MIN
    assert eng.saturation_count() == 0  # bf16 stores never clamp


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny", "siglip_so400m"])
def test_fp32_residual_stream_mode(cuda_device, monkeypatch, name):
    """fp16 towers keep the residual stream in fp16 by default; VLMFW_RESIDUAL=32 (read when the tower is created) keeps
    it in fp32. Both must meet the tolerance, and they must differ (i.e. the switch really switches)."""
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 5)
    imgs = weights.seeded_images_u8(3, spec.image_size, 8, smooth=True)
    outs = {}
    for mode in ("16", "32"):
        monkeypatch.setenv("VLMFW_RESIDUAL", mode)
        eng = TowerEngine(spec, "cuda:0")
        eng.load_hf_state_dict(sd)
        outs[mode] = eng.forward(imgs.cuda()).cpu()
        assert eng.hidden_state(3).dtype == (torch.float16 if mode == "16" else torch.float32)
        eng.close()
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    ref = O.tower_features(spec, weights.hf_to_canonical(spec, sd), pv) if spec.layers <= 4 else None
    if ref is not None:
        for mode in outs:
            assert torch.nn.functional.cosine_similarity(outs[mode], ref).min() >= COS_MIN
    assert torch.nn.functional.cosine_similarity(outs["16"], outs["32"]).min() >= 0.99999
    assert not torch.equal(outs["16"], outs["32"])


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
@pytest.mark.parametrize("residual", ["16", "32"])
def test_massive_activation_channels(cuda_device, monkeypatch, name, residual):
    """Real ViT checkpoints carry a few residual-stream channels hundreds of times larger than the rest. The fp16 stream
    must hold them (range 65504, relative precision 2^-11 at any magnitude): the first layer's attention output bias
    puts +-300 into three channels (the stream keeps them through every later layer) — against the fp32 oracle, both
    stream modes."""
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 4)
    key = next(k for k in sd if "layers.0.self_attn." in k and k.endswith(("out_proj.bias", "projection.bias")))
    bias = sd[key].clone()
    bias[0] += 300.0
    bias[5] -= 300.0
    bias[spec.hidden - 1] += 150.0
    sd[key] = bias
    monkeypatch.setenv("VLMFW_RESIDUAL", residual)
    eng = TowerEngine(spec, "cuda:0")
    eng.load_hf_state_dict(sd)
    imgs = weights.seeded_images_u8(4, spec.image_size, 9, smooth=True)
    got = eng.forward(imgs.cuda()).cpu()
    if eng.cls_tail:
        eng.set_cls_tail(False)
        eng.forward(imgs.cuda())
    hid = eng.hidden_state(4).float().cpu()
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    can = weights.hf_to_canonical(spec, sd)
    ref = O.tower_features(spec, can, pv)
    ref_hid = O.last_hidden(spec, {k: v.float() for k, v in can.items()}, pv)
    assert ref_hid.abs().max() > 100  # the outlier channels really are in the stream
    assert torch.isfinite(got).all()
    assert torch.nn.functional.cosine_similarity(got, ref).min() >= COS_MIN
    rel = (hid - ref_hid).flatten(1).norm(dim=1) / ref_hid.flatten(1).norm(dim=1)
    assert rel.max() < 5e-3, rel


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_weights_beyond_fp16_range_keep_the_fp16_mode(cuda_device, name):
    """A checkpoint with weights beyond +-65504: the matrix is stored divided by a power of two and its GEMM's accumulator
    multiplied back (exact), so the tower stays in fp16 — the only mode inside the logit tolerance — instead of falling
    back to bf16. One plain matrix (fc2) and one LayerNorm-folded matrix (fc1: its scale comes from the FOLDED values)."""
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 3)
    pre = "" if name == "eva_tiny" else "vision_model."
    k1, k2 = pre + "encoder.layers.0.mlp.fc1.weight", pre + "encoder.layers.1.mlp.fc2.weight"
    kb1, k21 = pre + "encoder.layers.0.mlp.fc1.bias", pre + "encoder.layers.1.mlp.fc1.weight"
    kb21 = pre + "encoder.layers.1.mlp.fc1.bias"
    for k in (k1, k2, kb1, k21, kb21):
        sd[k] = sd[k].clone()
    # fc1 of layer 0: one output row far outside the fp16 range; its bias pushes the unit so far negative that the
    # activation is exactly 0 in fp32 and fp16 alike, so the huge values exercise the operand scale, not the output range
    sd[k1][3, :] *= 3.0e5
    sd[kb1][3] = -1.0e7
    # fc2 of layer 1: a single entry of 1e5 in input column 7, whose activation is exactly 0 (fc1 row 7 and bias zeroed)
    sd[k21][7, :] = 0
    sd[kb21][7] = 0
    sd[k2][0, 7] = 1.0e5
    eng = TowerEngine(spec, "cuda:0")
    eng.load_hf_state_dict(sd)
    assert eng.compute_dtype == "f16" and eng.tolerance_ok
    assert eng.weight_scale("L0.fc1.w") > 1.0 and eng.weight_scale("L1.fc2.w") == 4.0  # 1e5 / 4 <= 32768
    assert eng.weight_scale("L0.o.w") == 1.0 and eng.weight_scale("patch.w") == 1.0
    imgs = weights.seeded_images_u8(3, spec.image_size, 3, smooth=True)
    got = eng.forward(imgs.cuda()).cpu()
    ref = O.tower_features(spec, weights.hf_to_canonical(spec, sd), O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std))
    assert torch.isfinite(got).all() and torch.nn.functional.cosine_similarity(got, ref, dim=1).min() >= COS_MIN
    assert eng.saturation_count() == 0


def test_non_finite_weight_is_refused(cuda_device):
    from vlmfw_b200 import _lib

    spec = specs.SIGLIP_TINY
    sd = weights.seeded_hf_state_dict(spec, 3)
    key = "vision_model.encoder.layers.0.self_attn.out_proj.weight"
    sd[key] = sd[key].clone()
    sd[key][0, 0] = float("inf")
    eng = TowerEngine(spec, "cuda:0")
    with pytest.raises(_lib.VlmfwError) as ei:
        eng.load_hf_state_dict(sd)
    assert ei.value.code == _lib.E_RANGE


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny", "siglip_so400m"])
def test_saturation_counter(cuda_device, name):
    """fp16 stores clamp at +-65504 (cvt.rn.satfinite) instead of overflowing; the tower counts every clamped store so a
    caller can tell that a forward clipped. Massive-activation channels of realistic size (+-300) must NOT trip it; an
    attention-output bias of 1e5 must (the residual stream cannot hold it), and the features stay finite."""
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 4)
    key = next(k for k in sd if "layers.0.self_attn." in k and k.endswith(("out_proj.bias", "projection.bias")))
    imgs = weights.seeded_images_u8(2, spec.image_size, 9, smooth=True).cuda()
    for bump, clipped in ((300.0, False), (1.0e5, True)):
        sd2 = dict(sd)
        bias = sd[key].clone()
        bias[0] += bump
        bias[5] -= bump
        sd2[key] = bias
        eng = TowerEngine(spec, "cuda:0")
        eng.load_hf_state_dict(sd2)
        assert eng.saturation_count() == 0
        got = eng.forward(imgs)
        assert torch.isfinite(got).all()
        n = eng.saturation_count(reset=True)
        assert (n > 0) == clipped, f"{name}: bump {bump} -> {n} clamped stores"
        assert eng.saturation_count() == 0  # reset
        eng.close()


def test_clip_cls_strategy(cuda_device):
    spec = specs.CLIP_TINY.with_pool(specs.POOL_CLS)
    got, ref, *_ = _run(spec, 4, 5, False, "u8")
    assert torch.nn.functional.cosine_similarity(got, ref, dim=1).min() >= COS_MIN


def test_ragged_batches_and_determinism(cuda_device):
    spec = specs.SIGLIP_TINY
    sd = weights.seeded_hf_state_dict(spec, 1)
    eng = TowerEngine(spec, "cuda:0")
    eng.load_hf_state_dict(sd)
    imgs = weights.seeded_images_u8(9, spec.image_size, 1).cuda()
    full = eng.forward(imgs).clone()
    again = eng.forward(imgs).clone()
    assert torch.equal(full, again)  # bit-identical run to run
    one = torch.cat([eng.forward(imgs[i:i + 1]).clone() for i in range(9)])
    assert torch.equal(full, one)  # batch-size independent: images never interact
    assert eng.forward(imgs[:0]).shape == (0, spec.hidden)


@pytest.mark.parametrize("name,batch,golden", [("siglip_so400m", 256, "siglip_so400m"), ("clip_vit_l_336", 128, "clip_vit_l_336"),
                                               ("eva_vit_g", 128, "eva_vit_g")])
def test_full_size_batch_properties(cuda_device, name, batch, golden):
    """BASELINE.json's full sizes (SigLIP-So400m batch 256, CLIP-L/336 and EVA-g batch 128), where the CPU oracle would take
    minutes: size-independent properties instead — run-to-run determinism, batch-composition independence (a 256-image
    forward equals the concatenation of its sub-batches bit for bit: images never interact, tile order never depends
    on M), the CUDA-graph replay equals the eager launches, and the first images (the golden fixture's seed) still match
    the reference's own features."""
    import os

    import numpy as np

    spec = specs.SPECS[name]
    eng = TowerEngine(spec, "cuda:0")
    eng.load_hf_state_dict(weights.seeded_hf_state_dict(spec, 1))
    head = weights.seeded_images_u8(4, spec.image_size, 2, smooth=True)  # the images of tests/golden/<name>.npz
    rest = torch.randint(0, 256, (batch - 4, spec.image_size, spec.image_size, 3), dtype=torch.uint8,
                         generator=torch.Generator().manual_seed(3))
    imgs = torch.cat([head, rest]).cuda()
    full = eng.forward(imgs).clone()
    assert torch.isfinite(full).all()
    assert torch.equal(full, eng.forward(imgs))
    parts = torch.cat([eng.forward(imgs[lo:lo + 37]).clone() for lo in range(0, batch, 37)])
    assert torch.equal(full, parts)
    out = torch.empty_like(full)
    eng.graphed(imgs, out)()
    torch.cuda.synchronize()
    assert torch.equal(full, out)
    if golden is not None:
        g = np.load(os.path.join(os.path.dirname(__file__), "golden", golden + ".npz"))
        ref = torch.from_numpy(g["features"])
        assert torch.nn.functional.cosine_similarity(full[: ref.shape[0]].cpu(), ref).min() >= COS_MIN


def test_missing_weight_is_loud(cuda_device):
    from vlmfw_b200._lib import VlmfwError

    spec = specs.SIGLIP_TINY
    eng = TowerEngine(spec, "cuda:0")
    with pytest.raises(VlmfwError):
        eng.forward(torch.zeros(1, spec.image_size, spec.image_size, 3, dtype=torch.uint8, device="cuda"))


def test_cpu_device_is_refused():
    with pytest.raises(RuntimeError):
        TowerEngine(specs.SIGLIP_TINY, "cpu")
