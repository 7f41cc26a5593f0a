#!/usr/bin/env python
"""Timing of the image-I/O row at UHD size on one GPU next to the reference's host path (PIL /
torchvision on the box's CPU): JPEG decode, Resize, ToTensor, save_image quantisation, JPEG encode.  One JSON line to
stdout and gpurun_out/io_bench.json.   python tools/io_bench.py [H W]"""
import io
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from collaborative_distillation_b200 import image_io  # noqa: E402


def gpu_ms(fn, n=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def wall_ms(fn, n=3):
    fn()
    ts = []
    for _ in range(n):
        t = time.perf_counter()
        fn()
        ts.append((time.perf_counter() - t) * 1e3)
    return sorted(ts)[len(ts) // 2]


def main():
    from PIL import Image
    import torchvision.transforms as T
    H, W = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2160, 3840)
    size = H // 2
    y, x = np.mgrid[0:H, 0:W].astype(np.float32)
    rng = np.random.default_rng(0)
    img = np.stack([127 + 100 * np.sin(x * 0.01) * np.cos(y * 0.013), 127 + 100 * np.sin(x * 0.007 + y * 0.009), 255 * x / W * y / H], -1)
    img = np.clip(img + rng.normal(0, 6, img.shape), 0, 255).astype(np.uint8)
    buf = io.BytesIO()
    Image.fromarray(img).save(buf, format="JPEG", quality=90)
    data = buf.getvalue()
    codec = image_io.JpegCodec()
    dev = torch.from_numpy(img).cuda()
    oh, ow = image_io.resized_output_size(H, W, size)
    res = {"image": "%dx%d" % (W, H), "jpeg_bytes": len(data), "resize_to": "%dx%d" % (ow, oh), "gpu_ms": {}, "cpu_ms": {}}
    g, c = res["gpu_ms"], res["cpu_ms"]
    g["jpeg_decode(wall)"] = wall_ms(lambda: codec.decode(data))
    g["resize"] = gpu_ms(lambda: image_io.resize_u8(dev, oh, ow))
    g["to_tensor"] = gpu_ms(lambda: image_io.to_tensor(dev))
    t = image_io.to_tensor(dev)
    g["quantize"] = gpu_ms(lambda: image_io.quantize(t))
    g["jpeg_encode(wall)"] = wall_ms(lambda: codec.encode(dev, 75, "420"))
    mp = H * W / 1e6
    res["gpu_gbs"] = {"to_tensor": mp * 15e-3 / g["to_tensor"] * 1e3, "quantize": mp * 15e-3 / g["quantize"] * 1e3,
                      "resize": (3 * (H * W + H * ow) + 3 * (H * ow + oh * ow)) / 1e9 / g["resize"] * 1e3}
    pil = Image.fromarray(img)
    c["jpeg_decode"] = wall_ms(lambda: Image.open(io.BytesIO(data)).convert("RGB"))
    c["resize"] = wall_ms(lambda: T.Resize(size)(pil))
    c["to_tensor"] = wall_ms(lambda: T.ToTensor()(pil))
    tc = t.cpu()
    c["save_image_quantize"] = wall_ms(lambda: tc[0].mul(255).add_(0.5).clamp_(0, 255).permute(1, 2, 0).to(torch.uint8))
    c["jpeg_encode"] = wall_ms(lambda: pil.save(io.BytesIO(), format="JPEG"))
    res["cpu_threads"] = torch.get_num_threads()
    line = json.dumps(res)
    print(line)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    open(os.path.join(ROOT, "gpurun_out", "io_bench.json"), "w").write(line + "\n")


if __name__ == "__main__":
    main()
