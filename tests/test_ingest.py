"""Image ingest: the host weight tables of libvlmfw_b200.so against the oracle (CPU), and the GPU
resize / crop against the oracle and the committed processor outputs — bit-exact, ragged batches, edge sizes."""
import os
import sys
import zlib

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))

from make_golden_preprocess import CASES, seeded_image  # noqa: E402

from oracle import preprocess_oracle as po  # noqa: E402
from vlmfw_b200 import ingest, specs  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "preprocess.npz"))
MODES = {"square": ingest.RESIZE_SQUARE, "shortest_crop": ingest.RESIZE_SHORTEST_CROP}


# ------------------------------------------------------------------------------------------------ CPU: host logic
@pytest.mark.parametrize("n_in,n_out", [(500, 224), (61, 224), (2000, 224), (225, 224), (223, 224), (640, 336),
                                        (4000, 336), (17, 224), (1, 56), (7, 70), (224, 224)])
def test_host_weight_tables_equal_oracle(n_in, n_out):
    xmin, xsize, w, prec = ingest.resize_weights(n_in, n_out)
    if n_in == n_out:  # identity pass: one tap of weight 2^prec (the oracle skips the pass)
        assert w.shape[1] == 1 and (xsize == 1).all() and (xmin == np.arange(n_out)).all() and (w == 1 << prec).all()
        return
    oxmin, oxsize, ow, oprec = po.aa_weights_int16(n_in, n_out)
    assert prec == oprec
    np.testing.assert_array_equal(xmin, oxmin)
    np.testing.assert_array_equal(xsize, oxsize)
    np.testing.assert_array_equal(w, ow)


def test_host_weight_tables_crop_window_keeps_axis_precision():
    """A crop window uses the precision of the WHOLE axis (ATen picks it from the maximum weight over all outputs)."""
    full = ingest.resize_weights(1000, 448)
    part = ingest.resize_weights(1000, 448, 56, 336)
    assert part[3] == full[3]
    np.testing.assert_array_equal(part[0], full[0][56:392])
    np.testing.assert_array_equal(part[2], full[2][56:392])


def test_pack_images_layout():
    imgs = [np.full((3, 5, 3), 7, np.uint8), np.full((2, 2, 3), 9, np.uint8)]
    if not torch.cuda.is_available():
        pytest.skip("pinned host memory needs a CUDA runtime")
    host, off, h, w = ingest.pack_images(imgs)
    assert off.tolist() == [0, 48] and h.tolist() == [3, 2] and w.tolist() == [5, 2]
    assert host[:45].eq(7).all() and host[48:60].eq(9).all()


def test_ingest_needs_a_gpu():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback|not available"):
        ingest.ImageIngest(ingest.RESIZE_SQUARE, 224, "cuda")


# ------------------------------------------------------------------------------------------------ GPU: kernels
def _mode(spec):
    return "shortest_crop" if spec.arch == specs.ARCH_CLIP else "square"


@pytest.mark.gpu
def test_gpu_resize_equals_golden_processor_outputs(cuda_device):
    """Every committed processor output (tests/golden/preprocess.npz), one ragged batch per tower."""
    for name in ("siglip_so400m", "clip_vit_l_336", "eva_vit_g"):
        spec = specs.SPECS[name]
        idx = [i for i, c in enumerate(CASES) if c[0] == name]
        imgs = [seeded_image(*CASES[i][1:]) for i in idx]
        out = ingest.ImageIngest.for_spec(spec, cuda_device)(imgs).cpu().numpy()
        for k, i in enumerate(idx):
            assert zlib.crc32(out[k].tobytes()) == int(GOLD[f"c{i}_out_crc"]), (name, CASES[i])
            np.testing.assert_array_equal(out[k][:4], GOLD[f"c{i}_rows"])


@pytest.mark.gpu
@pytest.mark.parametrize("mode,size", [("square", 224), ("shortest_crop", 336), ("square", 56), ("shortest_crop", 70)])
def test_gpu_resize_equals_oracle_ragged(cuda_device, mode, size):
    rng = np.random.default_rng(size)
    shapes = [(size, size), (1, 1), (size + 1, size - 1), (1200, 640), (size, 3 * size), (2 * size, size), (5, 5),
              (size * 8 + 3, size * 8 - 5)]
    if mode == "square":  # extreme aspect ratios only where the resized image stays small enough for the oracle
        shapes += [(1, 300), (300, 1), (37, 911)]
    shapes += [(int(rng.integers(2, 700)), int(rng.integers(2, 700))) for _ in range(12)]
    imgs = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for h, w in shapes]
    imgs[1][:] = 255  # saturated single pixel: overshoot of the negative lobes must clamp, not wrap
    ing = ingest.ImageIngest(MODES[mode], size, cuda_device)
    out = ing(imgs).cpu().numpy()
    assert ing.last_launches == 2
    for k, im in enumerate(imgs):
        np.testing.assert_array_equal(out[k], po.preprocess_u8(im, mode, size), err_msg=f"{mode} {im.shape}")


@pytest.mark.gpu
def test_gpu_resize_extremes_and_properties(cuda_device):
    """Size-independent properties at sizes the oracle would take long on: constant images stay constant, an image that
    is already S x S passes through untouched, wide images (shared-memory row staging at its limit) work."""
    ing = ingest.ImageIngest(ingest.RESIZE_SQUARE, 224, cuda_device)
    big = np.full((3000, 4000, 3), 0, np.uint8)
    big[..., 0], big[..., 1], big[..., 2] = 13, 200, 255
    wide = np.full((8, 16384, 3), 77, np.uint8)
    native = np.random.default_rng(0).integers(0, 256, (224, 224, 3), dtype=np.uint8)
    out = ing([big, wide, native]).cpu().numpy()
    assert (out[0] == np.array([13, 200, 255], np.uint8)).all()
    assert (out[1] == 77).all()
    np.testing.assert_array_equal(out[2], native)
    with pytest.raises(Exception, match="unsupported size"):
        ing([np.zeros((4, 16385, 3), np.uint8)])
    assert ing([]).shape == (0, 224, 224, 3)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", [1, 2, 3])
def test_gpu_resize_random_sweep(cuda_device, seed):
    """Random sizes between 1 and ~900 pixels per side, both resize modes, in ONE ragged batch per mode: up- and
    down-scaling on either axis, identity on one axis, heavy aspect ratios (bounded so the oracle stays fast)."""
    rng = np.random.default_rng(seed)
    for mode, size in (("square", 224), ("shortest_crop", 336)):
        imgs = []
        for _ in range(24):
            h = int(rng.choice([1, 2, 7, size - 1, size, size + 1, int(rng.integers(3, 900))]))
            lim = (max(h // 4, 1), h * 4 + 4) if mode == "shortest_crop" else (1, 900)
            w = int(np.clip(rng.choice([1, 3, size - 1, size, size + 1, int(rng.integers(3, 900))]), *lim))
            kind = int(rng.integers(0, 3))
            if kind == 0:
                im = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
            elif kind == 1:  # saturated blocks: the bicubic lobes overshoot on both sides of every edge
                im = (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)
            else:
                im = np.full((h, w, 3), int(rng.integers(0, 256)), np.uint8)
            imgs.append(im)
        out = ingest.ImageIngest(MODES[mode], size, cuda_device)(imgs).cpu().numpy()
        for k, im in enumerate(imgs):
            np.testing.assert_array_equal(out[k], po.preprocess_u8(im, mode, size), err_msg=f"{mode} {im.shape}")
