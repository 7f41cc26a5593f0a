#!/usr/bin/env python
"""Drop-in for the reference CLI `PytorchWCT/WCT.py` (flags: WCT.py:15-34; mode->weights tables: :36-75;
log printer: :78-85; 5-stage loop: :120-125; output naming: :127) running the B200-native path.

    cd PytorchWCT && python WCT.py --debug --mode 16x [--UHD] [--alpha a] [--content_size n] [--style_size n] ...

Everything between image decode and image encode stays on the GPU (`WCT.stylize`); additive flags:
  --precision {tf32,fp32}   conv engine (default tf32 tensor cores)       --weights_root DIR  (default ../trained_models)
  --gpu_io                  decode (nvJPEG), Resize, ToTensor, save_image quantisation and JPEG encode on the GPU
                            (collaborative_distillation_b200.image_io; PIL-bit-exact resize / conversions)
"""
import argparse
import os
import sys
import time

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))

FLAGS = [  # (name, kwargs) -- same names, types, defaults and help as the reference parser
    ("--UHD_contentPath", dict(type=str, default="content/UHD_content")),
    ("--UHD_stylePath", dict(type=str, default="style/UHD_style")),
    ("--contentPath", dict(type=str, default="content")),
    ("--stylePath", dict(type=str, default="style")),
    ("--texturePath", dict(type=str, default="style/texture")),
    ("--outf", dict(type=str, default="stylized_results", help="folder to output images")),
    ("--picked_content_mark", dict(type=str, default=".")),
    ("--picked_style_mark", dict(type=str, default=".")),
    ("--mode", dict(type=str, default=None, choices=["original", "16x", "16x_kd2sd"], help="to choose different trained models")),
    ("--UHD", dict(action="store_true", help="if use the UHD images")),
    ("--synthesis", dict(action="store_true", help="for style synthesis")),
    ("--content_size", dict(type=int, default=0, help="resize content, leave it to 0 if not resize")),
    ("--style_size", dict(type=int, default=0, help="resize style, leave it to 0 if not resize")),
    ("--alpha", dict(type=float, default=1, help="hyperparameter to blend wct feature and content feature")),
    ("--log_mark", dict(type=str, default=time.strftime("%Y%m%d-%H%M"))),
    ("--num_run", dict(type=int, default=1, help="you can run WCT for multiple times")),
    ("--debug", dict(action="store_true")),
    ("--numpy", dict(action="store_true", help="use the numpy variant of whiten_and_color (content covariance + I)")),
    # additive
    ("--precision", dict(type=str, default="tf32", choices=["tf32", "fp32"])),
    ("--weights_root", dict(type=str, default="../trained_models")),
    ("--gpu_io", dict(action="store_true", help="image decode / resize / encode on the GPU (nvJPEG + libwctb kernels)")),
]
WEIGHT_DIRS = {  # WCT.py:36-70
    "original": ("original_wct_models/vgg_normalised_conv%d_1.t7", "original_wct_models/feature_invertor_conv%d_1.t7"),
    "16x": ("wct_se_16x_new/%dSE.pth", "wct_se_16x_new_sd/%dSD.pth"),
    "16x_kd2sd": ("wct_se_16x_new/%dSE.pth", "wct_se_16x_new_sd_kd2sd/%dSD.pth"),
}


def parse(argv=None):
    ap = argparse.ArgumentParser(description="WCT Pytorch (B200-native)")
    for name, kw in FLAGS:
        ap.add_argument(name, **kw)
    args = ap.parse_args(argv)
    enc, dec = WEIGHT_DIRS[args.mode or "original"]
    for k in range(1, 6):
        setattr(args, "e%d" % k, os.path.join(args.weights_root, enc % k))
        setattr(args, "d%d" % k, os.path.join(args.weights_root, dec % k))
    return args


class LogPrinter:
    """stdout with --debug, else append to <outf>/log_<mark>_<mode>.txt (WCT.py:78-85)"""

    def __init__(self, debug, f):
        self.log = sys.stdout if debug else open(f, "a+")

    def __call__(self, sth):
        print(str(sth), file=self.log, flush=True)


class _DeviceLoader:
    """--gpu_io: same (content, style, [name]) items as DataLoader(batch_size=1) over data_loader.Dataset, but decoded,
    resized and converted on the GPU (data_loader.py:46-76 -> image_io.load_image)."""

    def __init__(self, dataset, image_io):
        self.ds, self.io = dataset, image_io

    def __len__(self):
        return len(self.ds)

    def __iter__(self):
        for i in range(len(self.ds)):
            c, s, name = self.ds.paths(i)
            if c is None:      # texture synthesis: noise content of the texture's size (data_loader.py:61-76)
                style = self.io.load_image(s, self.ds.style_size, longer_side=True)
                yield torch.rand_like(style), style, [name]
            else:
                yield self.io.load_image(c, self.ds.content_size), self.io.load_image(s, self.ds.style_size), [name]


def main(argv=None):
    args = parse(argv)
    import torchvision.utils as vutils

    import collaborative_distillation_b200 as P
    from data_loader import Dataset

    os.makedirs(args.outf, exist_ok=True)
    log = LogPrinter(args.debug, os.path.join(args.outf, "log_%s_%s.txt" % (args.log_mark, args.mode)))
    log(args._get_kwargs())
    P.set_precision(args.precision)
    dataset = Dataset(args.UHD_contentPath if args.UHD else args.contentPath, args.UHD_stylePath if args.UHD else args.stylePath,
                      args.texturePath, args.content_size, args.style_size, args.picked_content_mark, args.picked_style_mark,
                      args.synthesis)
    if args.gpu_io:
        loader = _DeviceLoader(dataset, P.image_io)
    else:
        loader = torch.utils.data.DataLoader(dataset=dataset, batch_size=1, shuffle=False)
    wct = P.WCT(args).cuda()
    log("Number of content-style pairs: %s" % len(loader))
    total, n = 0.0, 0
    style_cache = {}    # style name -> per-stage style statistics/eigensystems: every style is encoded once, not once per pair
    for i, (cImg, sImg, imname) in enumerate(loader):
        imname = imname[0]
        log("\n" + "*" * 30 + ' #%s: Transferring "%s"' % (i, imname))
        start = time.time()
        skey = imname.rsplit(".", 1)[0].split("+")[-1]
        if skey not in style_cache:
            style_cache[skey] = wct.prepare_style(sImg.cuda())
        out = wct.stylize(cImg.cuda(), None, alpha=args.alpha, num_run=args.num_run, style_cache=style_cache[skey])   # WCT.py:120-125
        out_path = os.path.join(args.outf, "%s_mode=%s_alpha=%s_%s" % (args.log_mark, args.mode, args.alpha, imname))
        if args.gpu_io:
            P.image_io.save_image(out, out_path)               # quantise + JPEG-encode on the device; only the bitstream comes down
        else:
            vutils.save_image(out.cpu(), out_path)                                             # WCT.py:127-128 (timed, like the reference)
        dt = time.time() - start
        total, n = total + dt, n + 1
        log("Elapsed time is: %.4f seconds" % dt)
    log("Processed %d images. Average processing time per pair is: %.4f seconds" % (n, total / max(n, 1)))


if __name__ == "__main__":
    main()
