"""CPU: host logic and the C-ABI surface (no compute calls — there is no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

from vlmfw_b200 import _lib, models, parallel, probes, specs, weights
from vlmfw_b200.vision_backbone import backbone_from_spec

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_functions():
    src = open(os.path.join(ROOT, "include", "vlmfw_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vlmfw_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_lib.LIB_PATH), "build the library first (__graft_entry__.build())"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _header_functions()
    assert len(names) >= 24
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/vlmfw_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)
    assert _lib.load().vlmfw_abi_version() == 2


def test_no_device_is_reported_not_faked():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert not _lib.device_ok()
    assert "no CPU fallback" in _lib.last_error() or "CUDA" in _lib.last_error()
    cfg = _lib.TowerCfg(arch=0, image_size=56, patch_size=14, hidden=144, heads=2, mlp=304, layers=1, act=0, pool=0)
    h = ctypes.c_void_p()
    assert _lib.load().vlmfw_tower_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.E_NODEVICE


def test_invalid_config_is_rejected_with_message():
    cfg = _lib.TowerCfg(arch=7, image_size=56, patch_size=14, hidden=144, heads=2, mlp=304, layers=1)
    h = ctypes.c_void_p()
    assert _lib.load().vlmfw_tower_create(ctypes.byref(cfg), ctypes.byref(h)) == _lib.E_INVALID
    assert "arch" in _lib.last_error()


@pytest.mark.parametrize("name", ["siglip_tiny", "clip_tiny", "eva_tiny"])
def test_hf_to_canonical_roundtrip(name):
    spec = specs.SPECS[name]
    sd = weights.seeded_hf_state_dict(spec, 0)
    can = weights.hf_to_canonical(spec, sd)
    assert sorted(can) == sorted(weights.canonical_names(spec))
    shapes = weights.canonical_shapes(spec)
    assert all(tuple(can[k].shape) == shapes[k] for k in can)
    d = spec.hidden
    if spec.arch != specs.ARCH_EVA:
        pre = "vision_model.encoder.layers.1.self_attn."
        assert torch.equal(can["L1.qkv.w"][d:2 * d], sd[pre + "k_proj.weight"])
        assert torch.equal(can["L1.qkv.b"][2 * d:], sd[pre + "v_proj.bias"])
    # checkpoint-style prefixes (probe model.pt: backbone.vision_model.<hf name>) are accepted
    can2 = weights.hf_to_canonical(spec, {"backbone.vision_model." + k: v for k, v in sd.items()})
    assert all(torch.equal(can[k], can2[k]) for k in can)
    bad = dict(sd)
    bad.pop(next(k for k in bad if "fc1.weight" in k))
    with pytest.raises(KeyError):
        weights.hf_to_canonical(spec, bad)


def test_seeded_state_dict_matches_hf_module_parameters():
    from oracle import hf_ref

    for name in ("siglip_tiny", "clip_tiny", "eva_tiny"):
        spec = specs.SPECS[name]
        model = hf_ref.build_hf_vision_model(spec, weights.seeded_hf_state_dict(spec, 0))  # strict inside
        assert {n for n, _ in model.named_parameters()} == {n for n, _ in weights.hf_param_names(spec)}


def test_spec_constants_match_survey():
    assert abs(specs.SIGLIP_SO400M.flops_per_image() / 1e9 - 218.97) < 0.01
    assert abs(specs.CLIP_VIT_L_336.flops_per_image() / 1e9 - 381.92) < 0.01
    assert abs(specs.EVA_VIT_G.flops_per_image() / 1e9 - 520.72) < 0.01
    assert (specs.SIGLIP_SO400M.tokens, specs.CLIP_VIT_L_336.tokens, specs.EVA_VIT_G.tokens) == (256, 577, 257)


def test_shard_range_partitions_in_order():
    for n in (0, 1, 7, 256, 1000):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
    with pytest.raises(ValueError):
        parallel.shard_range(4, 2, 2)


def test_factory_surface_is_the_references():
    f = models.VLMModelFactory
    assert sorted(f.get_available_models()) == ["blip2", "llava", "paligemma"]
    with pytest.raises(ValueError, match="non trovato"):
        f.create_model("nope")
    assert f._registry["paligemma"].OUTPUT_DIM == 1152 and f._registry["llava"].OUTPUT_DIM == 1024
    assert f._registry["blip2"].OUTPUT_DIM == 1408

    class FakeRefFactory:
        _registry = {"blip2": object, "llava": object, "paligemma": object}

    models.install_into_reference(FakeRefFactory)
    assert FakeRefFactory._registry["llava"] is models.LLaVAModel and set(FakeRefFactory._registry) == {"blip2", "llava", "paligemma"}


def test_backbone_and_probe_contract_on_cpu():
    bb = backbone_from_spec(specs.SIGLIP_TINY, device="cpu")
    assert bb.output_dim == 144 and isinstance(bb, torch.nn.Module)
    keys = list(bb.state_dict())
    assert all(k.startswith("vision_model.") for k in keys)
    lin = probes.LinearProbe(bb, n_out_classes=9)
    assert lin._fully_frozen and not any(p.requires_grad for p in bb.parameters())
    assert {k for k in lin.state_dict() if k.startswith("classifier.")} == {
        "classifier.0.weight", "classifier.0.bias", "classifier.0.running_mean", "classifier.0.running_var",
        "classifier.0.num_batches_tracked", "classifier.2.weight", "classifier.2.bias"}
    deep = probes.LinearProbe(bb, 7, deeper_head=True)
    assert "classifier.5.weight" in deep.state_dict() and deep.classifier[2].out_features == 512
    multi = probes.MultiTaskProbe(bb, {"gender": 2, "age": 9})
    assert "heads.age.2.weight" in multi.state_dict()
    lin.unfreeze_last_backbone_k_layers(1, parts="attn", include_embeddings=False)
    assert not lin._fully_frozen
    names = bb.get_lora_target_names({"last_k": 1, "attn_only": True})
    assert names and all("encoder.layers.1.self_attn" in n for n in names)
    # no CPU fallback: the tower path refuses to run off-GPU, loudly
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bb(torch.zeros(1, 56, 56, 3, dtype=torch.uint8))
    with pytest.raises(ValueError):
        backbone_from_spec(specs.CLIP_TINY, device="cpu")(None, strategy="max")


def test_parsed_attributes_shape():
    out = probes.parsed_attributes({"gender": torch.tensor([1, 0]), "hat": torch.tensor([0, 1])})
    assert out == [{"gender": 1, "hat": 0}, {"gender": 0, "hat": 1}]
    assert probes.parsed_attributes(torch.tensor([3, 4]), ["age"]) == [{"age": 3}, {"age": 4}]


def test_sass_of_the_hot_kernels_is_tcgen05_and_tma():
    """Static proof, no GPU needed: the shipped .so carries sm_100a SASS in which every GEMM and tensor-core attention
    instantiation issues tcgen05 MMAs (UTCHMMA), is fed by TMA (UTMALDG), drains TMEM with tcgen05.ld (LDTM), and contains
    no warp-level mma.sync (HMMA) — B200_PROFILING.md's mnemonics. The generic attention fallback is the one mma.sync
    kernel and is not on the three towers' path."""
    import re
    import shutil
    import subprocess

    tool = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(tool):
        pytest.skip("cuobjdump not available")
    so = os.path.join(os.path.dirname(_lib.__file__), "libvlmfw_b200.so")
    out = subprocess.run([tool, "-sass", so], capture_output=True, text=True, timeout=300).stdout
    assert "sm_100a" in out or "EF_CUDA_SM100" in out, "no sm_100a code object in the library"
    per_fn, per_fn_seq = {}, {}
    name = None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            per_fn[name], per_fn_seq[name] = set(), []
        elif name is not None:
            m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
            if m:
                per_fn[name].add(m.group(1))
                per_fn_seq[name].append(m.group(1))
    gemms = {f: ops for f, ops in per_fn.items() if "gemm_bf16_tn_kernel" in f}
    atts = {f: ops for f, ops in per_fn.items() if "attention_tc_kernel" in f or "attention_wb_kernel" in f}
    wbs = {f: ops for f, ops in per_fn.items() if "attention_wb_kernel" in f}
    assert len(gemms) >= 6 and len(atts) >= 8 and len(wbs) == 6, (len(gemms), len(atts), len(wbs))
    for f, ops in wbs.items():  # the three towers' attention: P written to and read from tensor memory
        assert "STTM" in ops, f"{f}: probabilities not stored to TMEM"
    # Single-thread roles (TMA producer, MMA issuer) are warp-uniform loops with elect.sync around the issue: in divergent
    # `if (lane == 0)` code ptxas wraps every UTCHMMA / UTMALDG / UTCBAR in its own ELECT + R2UR.BROADCAST loop, which cost
    # the GEMMs 15-20 % of their tensor-pipe time. In the elect form the MMAs of a k-block
    # (four for the GEMMs' 64-wide stage, four or five for a 64-key score block) sit back to back in the instruction stream.
    def longest_run(seq, op):
        best = cur = 0
        for o in seq:
            cur = cur + 1 if o == op else 0
            best = max(best, cur)
        return best

    for f in list(wbs) + list(gemms):
        assert "ELECT" in per_fn[f], f"{f}: no elect.sync"
        need = 4 if f in gemms else 2  # attention: ptxas weaves the tail tile's descriptor arithmetic between the MMAs
        assert longest_run(per_fn_seq[f], "UTCHMMA") >= need, f"{f}: tcgen05.mma issued one per ELECT / R2UR loop"
    for f, ops in {**gemms, **atts}.items():
        assert "UTCHMMA" in ops, f"{f}: no tcgen05.mma"
        assert "UTMALDG" in ops, f"{f}: operands not loaded by TMA"
        assert "LDTM" in ops, f"{f}: accumulators not read back from TMEM"
        assert not any(op.startswith("HMMA") for op in ops), f"{f}: mma.sync in a tcgen05 kernel"
    assert any("UTMASTG" in ops for ops in gemms.values()), "no TMA store in any GEMM epilogue"
    assert all("UTMASTG" in ops for ops in atts.values()), "attention epilogue does not store through TMA"


def test_train_step_entry_points_validate_arguments_without_a_gpu():
    """Argument checks of the probe-training entry points come before any device work: bad arguments are E_INVALID with
    a message; good arguments on a box without a GPU are E_NODEVICE (reported, not faked)."""
    lib = _lib.load()
    buf = (ctypes.c_uint8 * 64)()
    ctr = (ctypes.c_int64 * 1)()
    addr = ctypes.addressof(buf)
    assert lib.vlmfw_dropout_keep_mask(addr, 4, 0, 0, 1, 0.3, None) == _lib.E_INVALID  # cols = 0
    assert "dropout_keep_mask" in _lib.last_error()
    assert lib.vlmfw_dropout_keep_mask(addr, 4, 16, 0, 1, 1.0, None) == _lib.E_INVALID  # p = 1 keeps nothing
    assert lib.vlmfw_dropout_keep_mask(addr, 4, 16, -1, 1, 0.3, None) == _lib.E_INVALID  # negative global row
    assert lib.vlmfw_dropout_keep_mask_dev(addr, 4, 16, 0, 1, None, 262, 0.3, None) == _lib.E_INVALID
    assert "steps_done" in _lib.last_error()
    assert lib.vlmfw_counter_add(None, 1, None) == _lib.E_INVALID
    assert lib.vlmfw_probe_adamw_dev(addr, addr, addr, addr, 4, 5, None, 1e-4, 0.9, 0.999, 1e-8, 1e-4, None) == _lib.E_INVALID
    assert lib.vlmfw_probe_bn_stats(None, 4, 16, addr, None) == _lib.E_INVALID
    if not torch.cuda.is_available():
        assert lib.vlmfw_dropout_keep_mask(addr, 4, 16, 0, 1, 0.3, None) == _lib.E_NODEVICE
        assert lib.vlmfw_counter_add(ctypes.addressof(ctr), 1, None) == _lib.E_NODEVICE
        assert ctr[0] == 0  # nothing ran on the host instead
