"""Generate tests/golden/preprocess.npz: outputs of the Hugging Face image processors the reference calls
(models/paligemma.py:108, llava.py:109, blip2.py:91 -> transformers TorchvisionBackend -> torchvision resize on uint8)
for seeded images of assorted non-native sizes. Stores, per case, a CRC of the input and the uint8 image (or its CRC
and first rows) the processor produced before its fp32 rescale + normalize (recovered exactly from pixel_values: x = round(pv * 255 std + 255 mean)).

    python tests/golden/make_golden_preprocess.py
"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from PIL import Image  # noqa: E402

from oracle import hf_ref  # noqa: E402
from vlmfw_b200 import specs  # noqa: E402

# (tower spec name, height, width, seed, smooth)
CASES = [
    ("siglip_so400m", 300, 400, 1, False), ("siglip_so400m", 97, 61, 2, True), ("siglip_so400m", 224, 500, 3, False),
    ("siglip_so400m", 1000, 224, 4, True), ("siglip_so400m", 17, 2000, 5, False), ("siglip_so400m", 225, 223, 6, False),
    ("clip_vit_l_336", 480, 640, 7, False), ("clip_vit_l_336", 200, 150, 8, True), ("clip_vit_l_336", 336, 1000, 9, False),
    ("clip_vit_l_336", 1333, 337, 10, True), ("clip_vit_l_336", 50, 50, 11, False),
    ("eva_vit_g", 256, 256, 12, False), ("eva_vit_g", 31, 777, 13, True),
]


def seeded_image(h, w, seed, smooth):
    """uint8 [h, w, 3]; the smooth kind is a bilinear blow-up of an 8x8 field (exercises rounding on slow gradients)."""
    rng = np.random.default_rng(seed)
    if not smooth:
        return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    g = rng.uniform(0, 255, (8, 8, 3))
    yy = np.linspace(0, 7, h)[:, None, None]
    xx = np.linspace(0, 7, w)[None, :, None]
    y0, x0 = np.minimum(yy.astype(int), 6), np.minimum(xx.astype(int), 6)
    fy, fx = yy - y0, xx - x0
    y0, x0 = y0[:, 0, 0], x0[0, :, 0]
    out = (g[y0][:, x0] * (1 - fy) * (1 - fx) + g[y0 + 1][:, x0] * fy * (1 - fx)
           + g[y0][:, x0 + 1] * (1 - fy) * fx + g[y0 + 1][:, x0 + 1] * fy * fx)
    return np.clip(np.rint(out), 0, 255).astype(np.uint8)


def processor_u8(spec, img):
    proc = hf_ref.build_hf_processor(spec)
    pv = proc(images=[Image.fromarray(img)], return_tensors="pt")["pixel_values"][0].numpy().astype(np.float64)
    m = np.asarray(spec.pix_mean)[:, None, None] * 255.0
    s = np.asarray(spec.pix_std)[:, None, None] * 255.0
    x = pv * s + m
    u8 = np.rint(x)
    assert np.abs(x - u8).max() < 1e-3
    return np.ascontiguousarray(np.moveaxis(u8.astype(np.uint8), 0, -1)), pv.astype(np.float32)


def main():
    out = {}
    for i, (name, h, w, seed, smooth) in enumerate(CASES):
        spec = specs.SPECS[name]
        img = seeded_image(h, w, seed, smooth)
        u8, pv = processor_u8(spec, img)
        out[f"c{i}_out_crc"] = np.uint32(zlib.crc32(u8.tobytes()))
        out[f"c{i}_rows"] = u8[:4].copy()
        if i == 1:  # one full output; the others are pinned by CRC + first rows to keep the fixture small
            out[f"c{i}_u8"] = u8
        out[f"c{i}_in_crc"] = np.uint32(zlib.crc32(img.tobytes()))
        out[f"c{i}_pv_row0"] = pv[:, 0, :].copy()  # fp32 pixel_values of the first output row
    np.savez_compressed(os.path.join(HERE, "preprocess.npz"), **out)
    print("wrote preprocess.npz", sum(v.nbytes for v in out.values()))


if __name__ == "__main__":
    main()
