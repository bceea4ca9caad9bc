"""Throughput of the GPU image ingest (vlmfw_preproc_run) next to the HF processor on the host cores.

    python tools/bench_ingest.py [--batch 256] [--height 480] [--width 640] [--tower siglip_so400m]

Prints one JSON line: device-resident images/s and algorithmic GB/s (raw image bytes read + resized bytes written, against
the measured HBM copy peak), the end-to-end rate with the pinned-host -> device copy of the raw bytes inside the timed
region, and the CPU processor's rate on a bounded sample."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vlmfw_b200 import ingest, specs  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--tower", default="siglip_so400m")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--cpu-images", type=int, default=64)
    a = ap.parse_args()
    spec = specs.SPECS[a.tower]
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (a.height, a.width, 3), dtype=np.uint8) for _ in range(a.batch)]
    ing = ingest.ImageIngest.for_spec(spec, "cuda:0")
    host, off, hs, ws = ingest.pack_images(imgs)
    dev = host.cuda()
    out = torch.empty(a.batch, spec.image_size, spec.image_size, 3, dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        ing.run_packed(dev, off, hs, ws, out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        ing.run_packed(dev, off, hs, ws, out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    e0.record()
    for _ in range(a.iters):
        ing.run_packed(host.to("cuda:0", non_blocking=True), off, hs, ws, out)
    e1.record()
    torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / a.iters
    bytes_alg = host.numel() + out.numel()
    peak = 6549.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    line = {"kernel": "vlmfw_preproc_run (resize_rows + resize_cols)", "tower": a.tower, "batch": a.batch,
            "in_hw": [a.height, a.width], "out": spec.image_size, "ms": round(ms, 4),
            "images_per_s": round(a.batch / ms * 1e3, 1), "alg_gbps": round(bytes_alg / ms / 1e6, 1),
            "hbm_peak_gbps": peak, "e2e_images_per_s_incl_h2d": round(a.batch / ms_e2e * 1e3, 1),
            "h2d_bytes": host.numel()}
    try:
        from PIL import Image

        import transformers as tf

        # the processor the reference's adapters call, built offline with the checkpoint's constants
        size, kw = spec.image_size, dict(image_mean=list(spec.pix_mean), image_std=list(spec.pix_std))
        if spec.arch == specs.ARCH_CLIP:
            proc = tf.CLIPImageProcessor(size={"shortest_edge": size}, crop_size={"height": size, "width": size}, **kw)
        elif spec.arch == specs.ARCH_SIGLIP:
            proc = tf.SiglipImageProcessor(size={"height": size, "width": size}, resample=3, **kw)
        else:
            proc = tf.BlipImageProcessor(size={"height": size, "width": size}, **kw)
        pil = [Image.fromarray(x) for x in imgs[: a.cpu_images]]
        proc(images=pil[:4], return_tensors="pt")
        t0 = time.perf_counter()
        proc(images=pil, return_tensors="pt")
        dt = time.perf_counter() - t0
        line["cpu_hf_processor_images_per_s"] = round(len(pil) / dt, 1)
        line["cpu_cores"] = os.cpu_count()
    except Exception as err:  # the baseline is informative only
        line["cpu_hf_processor_images_per_s"] = None
        line["cpu_error"] = str(err)[:100]
    print(json.dumps(line))


if __name__ == "__main__":
    main()
