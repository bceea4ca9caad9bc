"""``VLMModel`` classes and the factory — the registration surface of the reference, unchanged in shape.

Mirrors reference models/base_model.py:6-100, models/paligemma.py:21-85, models/llava.py:8-83, models/blip2.py:8-71 and
models/model_factory.py:7-50: same class names, same ``(model_id, device, quantization)`` constructor, same registry
keys ``"blip2" | "llava" | "paligemma"``, same ``get_vision_backbone(cleanup=True)`` contract — but the backbone that
comes back runs on the sm_100a kernels. Text generation (``generate_text``) and bitsandbytes quantisation are outside
the hot path: they are passed straight through to Hugging Face when the full model is loaded.
"""
from __future__ import annotations

import gc
from abc import ABC, abstractmethod
from typing import Dict, Optional, Type

import torch

from .vision_backbone import BLIP2Backbone, LLaVABackbone, PaliGemmaBackbone, VisionBackbone


def _find_submodule(root, names):
    """``model.vision_tower`` moved under ``model.model`` in transformers 5; look in both places."""
    for holder in (root, getattr(root, "model", None)):
        if holder is None:
            continue
        for n in names:
            sub = getattr(holder, n, None)
            if sub is not None:
                return sub, holder, n
    raise AttributeError(f"{type(root).__name__} exposes none of {names}")


class VLMModel(ABC):
    """Base of the three VLM wrappers (reference models/base_model.py:6-52)."""

    DEFAULT_ID: str = ""
    HF_CLASS: str = ""
    TOWER_ATTRS = ("vision_tower",)
    BACKBONE: Type[VisionBackbone] = VisionBackbone
    OUTPUT_DIM = 0

    def __init__(self, model_id, device, quantization):
        self.model_id = model_id if model_id is not None else self.DEFAULT_ID
        self.device = device
        self.quantization = quantization
        from transformers import AutoProcessor

        self.processor = AutoProcessor.from_pretrained(self.model_id, use_fast=True)
        self._load_model()
        print(f"Modello {self.model_id} caricato con successo.")

    def get_model_kwargs(self) -> Dict:
        """fp32 / fp16 only: the B200 tower repacks weights to bf16 itself; 8bit/4bit (bitsandbytes) concern the
        language model, which is not on this path (reference models/base_model.py:76-100)."""
        kw: Dict = {"low_cpu_mem_usage": True}
        if self.quantization == "fp16":
            kw["torch_dtype"] = torch.float16
        elif self.quantization in ("8bit", "4bit"):
            raise NotImplementedError("bitsandbytes quantisation of the language model is outside the vlmfw_b200 path")
        return kw

    def _load_model(self):
        import transformers as tf

        self.model = getattr(tf, self.HF_CLASS).from_pretrained(self.model_id, **self.get_model_kwargs()).eval()

    @abstractmethod
    def _prompt(self, prompt: str) -> str:
        ...

    def generate_text(self, image, prompt: str, max_tokens: int = 100) -> str:
        """Pass-through to Hugging Face generation (reference models/base_model.py:54-74); not accelerated here."""
        if self.model is None:
            raise RuntimeError("generate_text after get_vision_backbone(cleanup=True): the language model was released")
        inputs = self.processor(images=image, text=self._prompt(prompt), return_tensors="pt").to(self.device)
        with torch.inference_mode():
            outputs = self.model.generate(**inputs, max_new_tokens=max_tokens)
            response = [o[len(i):] for i, o in zip(inputs.input_ids, outputs)]
        return self.processor.batch_decode(response, skip_special_tokens=True)[0].strip()

    def get_vision_backbone(self, cleanup: bool = True) -> VisionBackbone:
        tower, holder, attr = _find_submodule(self.model, self.TOWER_ATTRS)
        backbone = self.BACKBONE(self.processor, tower.float(), self.OUTPUT_DIM, self.device)
        if cleanup:
            for name in list(holder._modules.keys()):
                if name != attr:
                    delattr(holder, name)
            self.model = None
            self.processor = None
            gc.collect()
            if torch.cuda.is_available():
                torch.cuda.empty_cache()
        return backbone


class PaLIGemmaModel(VLMModel):
    DEFAULT_ID = "google/paligemma-3b-mix-224"
    HF_CLASS = "PaliGemmaForConditionalGeneration"
    BACKBONE = PaliGemmaBackbone
    OUTPUT_DIM = 1152

    def _prompt(self, prompt):
        return f"<image> {prompt}"


class LLaVAModel(VLMModel):
    DEFAULT_ID = "llava-hf/llava-1.5-7b-hf"
    HF_CLASS = "LlavaForConditionalGeneration"
    BACKBONE = LLaVABackbone
    OUTPUT_DIM = 1024

    def _prompt(self, prompt):
        conversation = [{"role": "user", "content": [{"type": "text", "text": prompt}, {"type": "image"}]}]
        return self.processor.apply_chat_template(conversation, add_generation_prompt=True)


class BLIP2OptModel(VLMModel):
    DEFAULT_ID = "Salesforce/blip2-opt-6.7b"
    HF_CLASS = "Blip2ForConditionalGeneration"
    TOWER_ATTRS = ("vision_model",)
    BACKBONE = BLIP2Backbone
    OUTPUT_DIM = 1408

    def _prompt(self, prompt):
        return f"Question: {prompt}. Answer:"


class VLMModelFactory:
    """Same registry keys and ``create_model`` signature as reference models/model_factory.py:7-50."""

    _registry = {"blip2": BLIP2OptModel, "llava": LLaVAModel, "paligemma": PaLIGemmaModel}

    @staticmethod
    def create_model(model_name: str, model_id: Optional[str] = None, device: torch.device = torch.device("cpu"),
                     quantization: str = "fp32") -> VLMModel:
        model_name = model_name.lower()
        if model_name not in VLMModelFactory._registry:
            raise ValueError(f"Modello '{model_name}' non trovato. "
                             f"Disponibili: {list(VLMModelFactory._registry.keys())}")
        return VLMModelFactory._registry[model_name](model_id, device, quantization)

    @staticmethod
    def get_available_models():
        return list(VLMModelFactory._registry.keys())


def install_into_reference(factory_cls=None) -> None:
    """Point an existing reference checkout at the B200 classes without editing its sources: rebinds the three
    ``_registry`` entries of ``models.model_factory.VLMModelFactory`` (keys unchanged) — see INTEGRATION.md."""
    if factory_cls is None:
        from models.model_factory import VLMModelFactory as factory_cls  # the reference's own module
    for key, cls in VLMModelFactory._registry.items():
        factory_cls._registry[key] = cls
