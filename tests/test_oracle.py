"""CPU: pin the oracle restatement against outputs of the reference itself (tests/golden/*.npz, produced by
tests/golden/make_golden.py importing /root/reference) and against the live Hugging Face modules."""
import os

import numpy as np
import pytest
import torch

from oracle import hf_ref, probe_oracle, tower_oracle as O
from vlmfw_b200 import specs, weights

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")
MIVIA_TASKS = {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}
CASES = {"siglip_tiny": specs.SIGLIP_TINY, "clip_tiny": specs.CLIP_TINY, "eva_tiny": specs.EVA_TINY,
         "siglip_so400m": specs.SIGLIP_SO400M}


def _load(name):
    return dict(np.load(os.path.join(GOLDEN, f"{name}.npz")))


def oracle_outputs(spec, g):
    sd = weights.seeded_hf_state_dict(spec, int(g["weight_seed"]))
    imgs = weights.seeded_images_u8(int(g["n"]), spec.image_size, int(g["image_seed"]), smooth=bool(g["smooth"]))
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    can = weights.hf_to_canonical(spec, sd)
    return O.tower_features(spec, can, pv), can, pv


@pytest.mark.parametrize("name", list(CASES))
def test_oracle_matches_reference_golden(name):
    spec, g = CASES[name], _load(name)
    feats, can, pv = oracle_outputs(spec, g)
    ref = torch.from_numpy(g["features"])
    # fp32 vs fp32, different summation orders only
    assert torch.nn.functional.cosine_similarity(feats, ref).min() > 0.999999
    torch.testing.assert_close(feats, ref, rtol=2e-4, atol=2e-4)
    lin = O.head_forward(feats, O.seeded_head_params(spec.hidden, 9, 0, seed=100))
    torch.testing.assert_close(lin, torch.from_numpy(g["linear_logits"]), rtol=5e-4, atol=5e-4)
    assert lin.argmax(1).tolist() == g["linear_pred"].tolist()
    deep = O.head_forward(feats, O.seeded_head_params(spec.hidden, 7, 512, seed=101))
    torch.testing.assert_close(deep, torch.from_numpy(g["deeper_logits"]), rtol=5e-4, atol=5e-4)
    assert deep.argmax(1).tolist() == g["deeper_pred"].tolist()
    for i, (t, c) in enumerate(MIVIA_TASKS.items()):
        lg = O.head_forward(feats, O.seeded_head_params(spec.hidden, c, 0, seed=200 + i))
        torch.testing.assert_close(lg, torch.from_numpy(g[f"multi_logits_{t}"]), rtol=5e-4, atol=5e-4)
        assert lg.argmax(1).tolist() == g[f"multi_pred_{t}"].tolist()
    if "features_cls" in g:
        cls = O.tower_features(spec, can, pv, mode=O.POOL_CLS)
        torch.testing.assert_close(cls, torch.from_numpy(g["features_cls"]), rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_oracle_matches_live_hf_modules(name):
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 5)
    model = hf_ref.build_hf_vision_model(spec, sd)
    imgs = weights.seeded_images_u8(3, spec.image_size, 6)
    pv = O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std)
    torch.testing.assert_close(O.tower_features(spec, weights.hf_to_canonical(spec, sd), pv),
                               hf_ref.hf_features(spec, model, pv), rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_u8_preprocess_equals_hf_processor(name):
    from PIL import Image

    spec = specs.SPECS[name]
    imgs = weights.seeded_images_u8(2, spec.image_size, 9)
    proc = hf_ref.ProcessorShim(hf_ref.build_hf_processor(spec))
    pv = proc(images=[Image.fromarray(a.numpy()) for a in imgs], text="x", return_tensors="pt")["pixel_values"]
    torch.testing.assert_close(O.preprocess_u8(imgs, spec.pix_mean, spec.pix_std), pv, rtol=0, atol=1e-6)


def test_head_forward_equals_torch_modules():
    from vlmfw_b200.probes import make_head, make_head_deeper

    x = torch.randn(5, 64)
    for hidden, mk in ((0, lambda: make_head(64, 4, 0.3)), (32, lambda: make_head_deeper(64, 4, 32, 0.3))):
        p = O.seeded_head_params(64, 4, hidden, seed=3)
        m = mk().eval()
        m.load_state_dict(p)
        torch.testing.assert_close(O.head_forward(x, p), m(x), rtol=1e-5, atol=1e-5)


def test_phase_split_equals_reference_step():
    """sum over shards of the phase outputs, then AdamW == the reference's autograd step on the whole batch."""
    torch.manual_seed(0)
    d, c, n, p = 48, 5, 24, 0.3
    state = O.seeded_head_params(d, c, 0, seed=1)
    x = torch.randn(n, d) * 2 + 0.5
    y = torch.randint(0, c, (n,))
    y[::7] = -1
    keep = (torch.rand(n, d) >= p).to(torch.uint8)
    cw = torch.rand(c) + 0.5
    ref_state, losses = probe_oracle.reference_train_steps(state, [(x, y, keep)], c, p, cw, lr=1e-3, weight_decay=1e-2)
    shards = [(0, 10), (10, 24)]
    stats = sum(probe_oracle.phase1_stats(x[a:b]) for a, b in shards)
    bucket = sum(probe_oracle.phase2_bucket(x[a:b], y[a:b], stats, n, state["0.weight"], state["0.bias"],
                                            state["2.weight"], state["2.bias"], cw, keep[a:b], p) for a, b in shards)
    wsum = bucket[-1]
    assert abs(float(bucket[-2] / wsum) - losses[0]) < 1e-9
    grad = bucket[:-2] / wsum
    params = torch.cat([state["2.weight"].flatten(), state["2.bias"], state["0.weight"], state["0.bias"]]).double()
    lr, wd, b1, b2, eps = 1e-3, 1e-2, 0.9, 0.999, 1e-8
    m, v = (1 - b1) * grad, (1 - b2) * grad * grad
    new = params * (1 - lr * wd) - (lr / (1 - b1)) * m / (v.sqrt() / (1 - b2) ** 0.5 + eps)
    want = torch.cat([ref_state["2.weight"].flatten(), ref_state["2.bias"], ref_state["0.weight"], ref_state["0.bias"]])
    torch.testing.assert_close(new.float(), want, rtol=1e-5, atol=1e-6)
    mu = stats[:d] / n
    var_unb = (stats[d:] / n - mu * mu) * n / (n - 1)
    torch.testing.assert_close((0.9 * state["0.running_mean"].double() + 0.1 * mu).float(), ref_state["0.running_mean"],
                               rtol=1e-5, atol=1e-6)
    torch.testing.assert_close((0.9 * state["0.running_var"].double() + 0.1 * var_unb).float(),
                               ref_state["0.running_var"], rtol=1e-5, atol=1e-6)


KEEP_KAT = [0, 0, 1, 1, 1, 0, 1, 1, 1, 0, 1, 1, 1, 1, 0, 1]  # computed independently with Python integers


def test_keep_mask_generator_properties():
    """The counter-based dropout generator (restated in numpy; the CUDA kernel is compared with it bit for bit on the
    GPU): slices of the global mask, Bernoulli(1 - p) rate, fresh draw per key."""
    n, d, p = 512, 1152, 0.3
    full = probe_oracle.keep_mask_reference(n, d, 0, 12345, p)
    for lo, hi in [(0, 171), (171, 342), (342, 512), (511, 512)]:
        assert torch.equal(probe_oracle.keep_mask_reference(hi - lo, d, lo, 12345, p), full[lo:hi])
    assert abs(full.float().mean().item() - (1 - p)) < 4 * (p * (1 - p) / (n * d)) ** 0.5
    other = probe_oracle.keep_mask_reference(n, d, 0, 12346, p)
    assert abs((other == full).float().mean().item() - (p * p + (1 - p) ** 2)) < 0.01
    assert bool(probe_oracle.keep_mask_reference(4, 7, 0, 1, 0.0).all())
    # known-answer vector: pins the generator itself (splitmix64 constants, key mixing, threshold) against regressions
    assert probe_oracle.keep_mask_reference(2, 8, 3, 42, 0.5).flatten().tolist() == KEEP_KAT
    # splitmix64's published first output for state 0
    assert int(probe_oracle._splitmix64(np.array([0], dtype=np.uint64))[0]) == 0xE220A8397B1DCDAF
