"""CPU: the preprocessing oracle (oracle/preprocess_oracle.py) pinned against (1) the committed outputs of the Hugging
Face processors the reference calls (tests/golden/preprocess.npz, made by tests/golden/make_golden_preprocess.py) and
(2) the live processors, which are installed on every box (transformers + torchvision)."""
import os
import sys
import zlib

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from make_golden_preprocess import CASES, seeded_image  # noqa: E402

from oracle import preprocess_oracle as po  # noqa: E402
from vlmfw_b200 import specs  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))


def _mode(spec):
    return "shortest_crop" if spec.arch == specs.ARCH_CLIP else "square"


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_matches_golden_processor_output(i):
    name, h, w, seed, smooth = CASES[i]
    spec = specs.SPECS[name]
    img = seeded_image(h, w, seed, smooth)
    assert zlib.crc32(img.tobytes()) == int(GOLD[f"c{i}_in_crc"]), "seeded image generator drifted from the fixture"
    u8 = po.preprocess_u8(img, _mode(spec), spec.image_size)
    assert u8.shape == (spec.image_size, spec.image_size, 3)
    np.testing.assert_array_equal(u8[:4], GOLD[f"c{i}_rows"])
    assert zlib.crc32(u8.tobytes()) == int(GOLD[f"c{i}_out_crc"])
    if f"c{i}_u8" in GOLD:
        np.testing.assert_array_equal(u8, GOLD[f"c{i}_u8"])
    pv = po.rescale_normalize(u8[None], spec.pix_mean, spec.pix_std)
    np.testing.assert_allclose(pv[0, :, 0, :], GOLD[f"c{i}_pv_row0"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name,h,w", [("siglip_so400m", 123, 457), ("clip_vit_l_336", 361, 349), ("eva_vit_g", 640, 480),
                                      ("clip_vit_l_336", 336, 336), ("siglip_so400m", 224, 224)])
def test_oracle_matches_live_processor(name, h, w):
    from PIL import Image

    from oracle import hf_ref

    spec = specs.SPECS[name]
    img = seeded_image(h, w, h * 7 + w, False)
    pv = hf_ref.build_hf_processor(spec)(images=[Image.fromarray(img)], return_tensors="pt")["pixel_values"].numpy()
    mine = po.rescale_normalize(po.preprocess_u8(img, _mode(spec), spec.image_size)[None], spec.pix_mean, spec.pix_std)
    np.testing.assert_allclose(mine, pv, rtol=0, atol=1e-6)


def test_weight_table_properties():
    """Taps stay inside the image, int16 weights sum to ~2^precision, identity size needs no table."""
    for n_in, n_out in [(500, 224), (61, 224), (2000, 224), (225, 224), (640, 336)]:
        xmin, xsize, w16, prec = po.aa_weights_int16(n_in, n_out)
        assert (xmin >= 0).all() and (xmin + xsize <= n_in).all() and (xsize > 0).all()
        sums = np.array([int(w16[i, :xsize[i]].astype(np.int64).sum()) for i in range(n_out)])
        assert np.abs(sums - (1 << prec)).max() <= w16.shape[1]
    img = seeded_image(224, 224, 1, False)
    assert po.resize_bicubic_aa_u8(img, 224, 224) is img
