"""Drop-in parity through the plugin surface (-m gpu): B200 backbones + probes called like the reference's testers do
(lists of PIL images), checked against tests/golden/*.npz — outputs of the reference itself — with north_star's
tolerances: pooled-feature cosine >= 0.999, probe-logit max abs error <= 1e-2, bit-exact argmax labels and parsed
attributes."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import tower_oracle as O  # noqa: E402
from vlmfw_b200 import probes, specs, weights  # noqa: E402
from vlmfw_b200.vision_backbone import backbone_from_spec  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MIVIA_TASKS = {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}
COS_MIN, LOGIT_TOL = 0.999, 1e-2
CASES = {"siglip_tiny": specs.SIGLIP_TINY, "clip_tiny": specs.CLIP_TINY, "eva_tiny": specs.EVA_TINY,
         "siglip_so400m": specs.SIGLIP_SO400M, "clip_vit_l_336": specs.CLIP_VIT_L_336, "eva_vit_g": specs.EVA_VIT_G}


def _check_labels(logits_gpu, ref_logits, ref_pred, pred_gpu):
    """Bit-exact labels against the reference's. On a mismatch the message carries the reference's own fp32 top-2
    margin of the offending rows, so a near-tie (margin below the logit tolerance) is visible as such."""
    top2 = torch.from_numpy(ref_logits).topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    got, want = pred_gpu.cpu(), torch.from_numpy(ref_pred)
    bad = got != want
    assert not bad.any(), f"labels differ at rows {bad.nonzero().flatten().tolist()}, fp32 top-2 margins {margin[bad].tolist()}"


@pytest.mark.parametrize("name", list(CASES))
def test_plugin_matches_reference_golden(cuda_device, name):
    from PIL import Image

    spec = CASES[name]
    g = dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))
    bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, int(g["weight_seed"])), device="cuda")
    imgs = weights.seeded_images_u8(int(g["n"]), spec.image_size, int(g["image_seed"]), smooth=bool(g["smooth"]))
    pil = [Image.fromarray(a.numpy()) for a in imgs]

    lin = probes.LinearProbe(bb, 9).cuda().eval()
    lin.classifier.load_state_dict(O.seeded_head_params(spec.hidden, 9, 0, seed=100))
    feats = lin.extract_features(pil)
    assert feats.shape == (len(pil), spec.hidden) and feats.dtype == torch.float32 and feats.is_cuda
    cos = torch.nn.functional.cosine_similarity(feats.cpu(), torch.from_numpy(g["features"]))
    assert cos.min() >= COS_MIN, cos
    logits = lin(pil)
    err = (logits.cpu() - torch.from_numpy(g["linear_logits"])).abs().max().item()
    assert err <= LOGIT_TOL, f"linear logits max abs err {err}"
    _check_labels(logits, g["linear_logits"], g["linear_pred"], lin.predict(pil))

    deep = probes.LinearProbe(bb, 7, deeper_head=True, hidden_dim=512).cuda().eval()
    deep.classifier.load_state_dict(O.seeded_head_params(spec.hidden, 7, 512, seed=101))
    err = (deep(pil).cpu() - torch.from_numpy(g["deeper_logits"])).abs().max().item()
    assert err <= LOGIT_TOL, f"deeper logits max abs err {err}"
    _check_labels(None, g["deeper_logits"], g["deeper_pred"], deep.predict(pil))

    multi = probes.MultiTaskProbe(bb, MIVIA_TASKS).cuda().eval()
    for i, (t, c) in enumerate(MIVIA_TASKS.items()):
        multi.heads[t].load_state_dict(O.seeded_head_params(spec.hidden, c, 0, seed=200 + i))
    out = multi(pil)["logits"]
    pred = multi.predict(pil)
    for t in MIVIA_TASKS:
        err = (out[t].cpu() - torch.from_numpy(g[f"multi_logits_{t}"])).abs().max().item()
        assert err <= LOGIT_TOL, f"{t} logits max abs err {err}"
        _check_labels(None, g[f"multi_logits_{t}"], g[f"multi_pred_{t}"], pred[t])
    attrs = probes.parsed_attributes(pred)
    assert len(attrs) == len(pil) and set(attrs[0]) == set(MIVIA_TASKS) and all(isinstance(v, int) for v in attrs[0].values())
    if "features_cls" in g:
        cls = bb(pil, strategy="cls").cpu()
        assert torch.nn.functional.cosine_similarity(cls, torch.from_numpy(g["features_cls"])).min() >= COS_MIN


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_non_native_sizes_resized_on_gpu_equal_hf_processor_path(cuda_device, name):
    """Images that are not already S x S: the default path resizes them on the GPU (ingest.ImageIngest), the opt-out
    keeps the unchanged HF processor (CPU resize, fp32 NCHW ingest). The GPU resize is bit-exact with the processor's
    uint8 resize, so both must agree with the oracle on the processor's pixel_values, and with each other almost to the
    bit (only the fused rescale + normalize differs in the last ulp)."""
    from PIL import Image

    from oracle import hf_ref

    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 2)
    proc = hf_ref.ProcessorShim(hf_ref.build_hf_processor(spec))
    bb = backbone_from_spec(spec, sd, device="cuda", processor=proc)
    rng = np.random.default_rng(5)
    pil = [Image.fromarray(rng.integers(0, 256, (h, w, 3), dtype=np.uint8)) for h, w in ((91, 91), (40, 133), (300, 57))]
    pil.append(Image.fromarray(rng.integers(0, 256, (64, 80), dtype=np.uint8), mode="L"))  # converted to RGB like HF does
    assert bb.gpu_ingest
    got = bb(pil).cpu()
    bb.gpu_ingest = False
    via_hf = bb(pil).cpu()
    pv = proc(images=pil, text="x", return_tensors="pt")["pixel_values"]
    ref = O.tower_features(spec, weights.hf_to_canonical(spec, sd), pv)
    assert torch.nn.functional.cosine_similarity(got, ref).min() >= COS_MIN
    assert torch.nn.functional.cosine_similarity(via_hf, ref).min() >= COS_MIN
    assert (got - via_hf).abs().max() <= 2e-3 * ref.abs().max()


def test_checkpoint_roundtrip_repacks_weights(cuda_device):
    """load_state_dict on the probe (reference testers: singletask_tester.py:68-80, strict=False) must reach the
    packed bf16 copy of the tower weights."""
    spec = specs.EVA_TINY
    bb = backbone_from_spec(spec, weights.seeded_hf_state_dict(spec, 1), device="cuda")
    lin = probes.LinearProbe(bb, 4).cuda().eval()
    imgs = weights.seeded_images_u8(2, spec.image_size, 1)
    a = lin.extract_features(imgs).clone()
    other = {"backbone.vision_model." + k: v for k, v in weights.seeded_hf_state_dict(spec, 2).items()}
    lin.load_state_dict(other, strict=False)
    b = lin.extract_features(imgs).clone()
    assert not torch.allclose(a, b)
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    ref = O.tower_features(spec, weights.hf_to_canonical(spec, weights.seeded_hf_state_dict(spec, 2)), pv)
    assert torch.nn.functional.cosine_similarity(b.cpu(), ref).min() >= COS_MIN


def test_host_buffer_cabi_path(cuda_device):
    """vlmfw_tower_forward_host: HOST uint8 in, HOST fp32 out — the call a non-torch host binds."""
    from vlmfw_b200.engine import TowerEngine

    spec = specs.CLIP_TINY
    eng = TowerEngine(spec, "cuda:0")
    eng.load_hf_state_dict(weights.seeded_hf_state_dict(spec, 3))
    imgs = weights.seeded_images_u8(5, spec.image_size, 3)
    host = eng.forward_host(imgs)
    dev = eng.forward(imgs.cuda()).cpu()
    assert host.device.type == "cpu" and torch.equal(host, dev)
