"""Timing probe of the GAN contrast-adaptation path (f-3): forward time per batch, tensor-core rate of the 3x3
convolutions against MEASURED_PEAKS.json, config #5 tail (raster u8 -> G image u8) in images/s.

    python tools/gan_probe.py [--batch 32] [--reps 5] [--size 304]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def conv_flops(H, W):
    """Algorithmic flops of the 22 3x3 convolutions for one image (2 * pixels * Cin * Cout * 9), + stem and head."""
    f = 2 * 9 * (H * W * 64 * 128 + (H // 2) * (W // 2) * 128 * 256 + 18 * (H // 4) * (W // 4) * 256 * 256
                 + (H // 2) * (W // 2) * 256 * 128 + H * W * 128 * 64)
    return f, 2 * 49 * 64 * H * W * 2


def main():
    import torch
    from octa_autosegmentation_b200 import gan, _lib

    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--size", type=int, default=304)
    ap.add_argument("--seed", type=int, default=1)
    args = ap.parse_args()
    H = W = args.size
    g = torch.Generator().manual_seed(args.seed)
    sd = {}
    shapes = {"model.1": (64, 1, 7, 7), "model.4": (128, 64, 3, 3), "model.8": (256, 128, 3, 3), "model.22": (128, 256, 3, 3),
              "model.26": (64, 128, 3, 3), "model.30": (1, 64, 7, 7)}
    for b in range(12, 21):
        shapes["model.%d.conv_block.1" % b] = shapes["model.%d.conv_block.5" % b] = (256, 256, 3, 3)
    for k, s in shapes.items():
        sd[k + ".weight"] = torch.randn(*s, generator=g) * (2.0 / (s[1] * s[2] * s[3])) ** 0.5
        sd[k + ".bias"] = torch.zeros(s[0])
    G = gan.ResnetGenerator9(sd, image_size=(H, W), max_images=args.batch)
    n = args.batch
    raster = (torch.rand(n, H, W, generator=g) > 0.8).to(torch.uint8).cuda() * 200
    bg = torch.randint(0, 255, (n, H, W), generator=g).to(torch.uint8).cuda()
    x = gan.prepare_input(raster, bg, gan.speckle_device(list(range(n)), H, W))
    y = torch.empty_like(x)
    for _ in range(2):
        G.forward(x, out=y)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = _lib.launch_count()
    e0.record()
    for _ in range(args.reps):
        G.forward(x, out=y)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.reps
    launches = (_lib.launch_count() - l0) // args.reps
    f3, f7 = conv_flops(H, W)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    tf = n * f3 / ms / 1e9
    print("forward: batch %d  %.2f ms  -> %.1f images/s, %d launches; 3x3 conv flops %.1f GF/image -> %.1f TFLOP/s over the whole forward "
          "(%.1f %% of the measured %.0f TF/s bf16)" % (n, ms, n / ms * 1e3, launches, f3 / 1e9, tf, 100 * tf / peaks["bf16_tflops"], peaks["bf16_tflops"]))
    t0 = time.time()
    e0.record()
    for _ in range(args.reps):
        out = gan.contrast_adapt(G, raster, bg, list(range(n)))
    e1.record()
    torch.cuda.synchronize()
    ms2 = e0.elapsed_time(e1) / args.reps
    print("config #5 tail (u8 raster + background -> speckle, input transform, generator, u8 image): %.2f ms per %d images -> %.1f images/s"
          % (ms2, n, n / ms2 * 1e3))
    G.close()


if __name__ == "__main__":
    main()
