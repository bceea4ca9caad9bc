"""vlmfw_b200 — B200-native vision-tower + linear-probe hot path behind the reference's plugin surface.

Import as ``vlmfw_b200`` (see ``vlmfw_b200/__init__.py``). Nothing here imports CUDA eagerly: ``_lib.load()`` opens
``libvlmfw_b200.so`` on first use and raises if it is missing — there is no CPU fallback.
"""
from .specs import (CLIP_TINY, CLIP_VIT_L_336, EVA_TINY, EVA_VIT_G, FACTORY_TOWERS, SIGLIP_SO400M, SIGLIP_TINY, SPECS,  # noqa: F401
                    TowerSpec)

__version__ = "0.1.0"
