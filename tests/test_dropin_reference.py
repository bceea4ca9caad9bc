"""Drop-in check against the LIVE reference checkout: the reference's own probe classes, trainer-side
bookkeeping and factory are driven over the B200 backbone classes — host logic only, no compute (this container has no
GPU, and the tower refuses to run off-GPU). Skipped where /root/reference does not exist (the GPU box); the compute side
of the same surface is checked there by tests/test_plugin_gpu.py against fixtures the reference itself produced.

Runs in a subprocess: the reference is a set of top-level packages (``models``, ``probing``) that must not leak into this
session's ``sys.modules``. Reference files are imported unmodified; the two harness shims are the ones of
tests/golden/make_golden.py (``huggingface_hub.login`` — import-time network call at models/paligemma.py:17-19 — and a
stub ``matplotlib``).
"""
import json
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "probing")), reason="needs the reference checkout")

CHILD = r"""
import json, sys, types
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT); sys.path.insert(0, REF)
import huggingface_hub
huggingface_hub.login = lambda *a, **k: None
for mod in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(mod, types.ModuleType(mod))
import torch
from models.model_factory import VLMModelFactory as RefFactory          # reference models/model_factory.py:7-50
from models.vision_backbone import VisionBackbone as RefVisionBackbone  # reference models/vision_backbone.py:6-67
from probing.models.linear_probe import LinearProbe                     # reference probing/models/linear_probe.py
from probing.models.multitask_probe import MultiTaskProbe               # reference probing/models/multitask_probe.py
import vlmfw_b200
from models.paligemma import PaliGemmaBackbone as RefPali               # reference models/paligemma.py:88-221
from models.llava import LLaVABackbone as RefLlava                      # reference models/llava.py:86-210
from models.blip2 import BLIP2Backbone as RefBlip                       # reference models/blip2.py:74-190
from oracle import hf_ref
from vlmfw_b200 import models as ours, probes as our_probes, specs, weights
REF_ADAPTERS = {"siglip": RefPali, "clip": RefLlava, "eva": RefBlip}
from vlmfw_b200.vision_backbone import backbone_from_spec

out = {}
# 1. registry rebinding: keys unchanged, reference create_model dispatches to the B200 classes, same ValueError
ref_keys = sorted(RefFactory._registry)
ours.install_into_reference()
out["keys_before"] = ref_keys
out["keys_after"] = sorted(RefFactory._registry)
out["bound"] = all(RefFactory._registry[k] is ours.VLMModelFactory._registry[k] for k in RefFactory._registry)
try:
    RefFactory.create_model("nope")
    out["unknown"] = "no error"
except ValueError as e:
    out["unknown"] = str(e)
calls = []
class Probe(ours.PaLIGemmaModel):
    def __init__(self, model_id, device, quantization):
        calls.append((model_id, str(device), quantization))
RefFactory._registry["paligemma"] = Probe
RefFactory.create_model("PaliGemma", None, torch.device("cpu"), "fp16")
out["ctor_args"] = calls

# 2. the reference's probes built over the B200 backbone: freeze bookkeeping, output_dim, heads, state-dict keys
for name, spec in (("siglip", specs.SIGLIP_TINY), ("clip", specs.CLIP_TINY), ("eva", specs.EVA_TINY)):
    bb = backbone_from_spec(spec, device="cpu")
    r = {"is_module": isinstance(bb, torch.nn.Module), "output_dim": bb.output_dim}
    r["ref_methods"] = all(hasattr(bb, m) for m in dir(RefVisionBackbone) if not m.startswith("_") and
                           callable(getattr(RefVisionBackbone, m)) and m in RefVisionBackbone.__dict__)
    lin = LinearProbe(bb, n_out_classes=9, freeze_backbone=True, dropout_p=0.3)
    mine = our_probes.LinearProbe(backbone_from_spec(spec, device="cpu"), n_out_classes=9, freeze_backbone=True, dropout_p=0.3)
    r["frozen"] = bool(lin._fully_frozen) and not any(p.requires_grad for p in bb.parameters())
    r["same_keys"] = list(lin.state_dict()) == list(mine.state_dict())
    r["same_shapes"] = all(tuple(a.shape) == tuple(b.shape) for a, b in zip(lin.state_dict().values(), mine.state_dict().values()))
    r["head_in"] = lin.classifier[2].in_features
    deep = LinearProbe(bb, n_out_classes=7, freeze_backbone=True, deeper_head=True, hidden_dim=512)
    deep_mine = our_probes.LinearProbe(backbone_from_spec(spec, device="cpu"), 7, deeper_head=True, hidden_dim=512)
    r["deep_same_keys"] = list(deep.state_dict()) == list(deep_mine.state_dict())
    multi = MultiTaskProbe(bb, tasks={"gender": 2, "upper": 12, "hat": 2}, freeze_backbone=True)
    multi_mine = our_probes.MultiTaskProbe(backbone_from_spec(spec, device="cpu"), {"gender": 2, "upper": 12, "hat": 2})
    r["multi_same_keys"] = list(multi.state_dict()) == list(multi_mine.state_dict())
    # a checkpoint written by the reference's probe loads into ours (and back) with strict=True
    mine.load_state_dict(lin.state_dict(), strict=True)
    lin.load_state_dict(mine.state_dict(), strict=True)
    multi_mine.load_state_dict(multi.state_dict(), strict=True)
    # reference BaseProbe.unfreeze_last_backbone_k_layers -> backbone.unfreeze_last_k_layers (base_probe.py:69-82)
    lin.unfreeze_last_backbone_k_layers(1, parts="attn", include_embeddings=False)
    trainable = [n for n, p in bb.named_parameters() if p.requires_grad]
    r["unfrozen_last_layer_only"] = bool(trainable) and all(f"layers.{spec.layers - 1}." in n for n in trainable)
    r["not_frozen_after"] = not lin._fully_frozen
    # 3. host-logic parity sweep: the reference's adapter over the HF module vs the B200 adapter, same arguments ->
    #    same set of trainable parameter names / LoRA target names (relative to the backbone)
    ref_bb = REF_ADAPTERS[name](hf_ref.ProcessorShim(hf_ref.build_hf_processor(spec)),
                                hf_ref.build_hf_vision_model(spec, weights.seeded_hf_state_dict(spec, 3)),
                                spec.hidden, torch.device("cpu"))
    mismatches = []
    for k in (0, 1, 2, 5):
        for parts in ("all", "attn", "mlp", "bogus"):
            for emb in (False, True):
                got = []
                for b in (ref_bb, bb):
                    for q in b.parameters():
                        q.requires_grad = False
                    b.unfreeze_last_k_layers(k, parts=parts, include_embeddings=emb)
                    got.append(sorted(n for n, q in b.named_parameters() if q.requires_grad))
                if got[0] != got[1]:
                    mismatches.append(["unfreeze", k, parts, emb, sorted(set(got[0]) ^ set(got[1]))[:6]])
    for strategy in ({}, {"last_k": 1}, {"last_k": 1, "attn_only": False}, {"last_k": 9, "attn_only": True},
                     {"last_k": 2, "attn_only": False}):
        a, b = ref_bb.get_lora_target_names(strategy), bb.get_lora_target_names(strategy)
        if a != b:
            mismatches.append(["lora", strategy, sorted(set(a) ^ set(b))[:6]])
    r["sweep_mismatches"] = mismatches
    r["n_params"] = [sum(1 for _ in ref_bb.parameters()), sum(1 for _ in bb.parameters())]
    r["param_names_equal"] = [n for n, _ in ref_bb.named_parameters()] == [n for n, _ in bb.named_parameters()]
    # eval()/train() propagate; the reference's extract_features reaches backbone.forward, which refuses to run off-GPU
    lin.eval()
    r["eval_propagates"] = not bb.training
    try:
        lin(torch.zeros(1, spec.image_size, spec.image_size, 3, dtype=torch.uint8))
        r["cpu_forward"] = "ran"
    except RuntimeError as e:
        r["cpu_forward"] = "no CPU fallback" in str(e)
    out[name] = r
print("RESULT " + json.dumps(out))
"""


@pytest.fixture(scope="module")
def result():
    p = subprocess.run([sys.executable, "-c", CHILD, ROOT, REF], capture_output=True, text=True, timeout=600,
                       env={**os.environ, "HF_HUB_OFFLINE": "1"})
    assert p.returncode == 0, p.stderr[-4000:]
    line = [l for l in p.stdout.splitlines() if l.startswith("RESULT ")][-1]
    return json.loads(line[len("RESULT "):])


def test_reference_factory_rebinds_to_b200_classes(result):
    assert result["keys_before"] == result["keys_after"] == ["blip2", "llava", "paligemma"]
    assert result["bound"]
    assert "non trovato" in result["unknown"]  # reference models/model_factory.py:34-37
    # create_model lower-cases the key and passes (model_id, device, quantization) positionally (model_factory.py:33-40)
    assert result["ctor_args"] == [[None, "cpu", "fp16"]]


@pytest.mark.parametrize("tower,dim", [("siglip", 144), ("clip", 128), ("eva", 176)])
def test_reference_probes_accept_the_b200_backbone(result, tower, dim):
    r = result[tower]
    assert r["is_module"] and r["ref_methods"]
    assert r["output_dim"] == dim == r["head_in"]
    assert r["frozen"]
    assert r["same_keys"] and r["same_shapes"] and r["deep_same_keys"] and r["multi_same_keys"]
    assert r["unfrozen_last_layer_only"] and r["not_frozen_after"]
    assert r["param_names_equal"] and r["n_params"][0] == r["n_params"][1]
    assert r["sweep_mismatches"] == []
    assert r["eval_propagates"]
    assert r["cpu_forward"] is True
