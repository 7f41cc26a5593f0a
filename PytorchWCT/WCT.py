#!/usr/bin/env python
"""Drop-in for the reference CLI `PytorchWCT/WCT.py` (flags: WCT.py:15-34; mode->weights tables: :36-75;
log printer: :78-85; 5-stage loop: :120-125; output naming: :127) running the B200-native path.

    cd PytorchWCT && python WCT.py --debug --mode 16x [--UHD] [--alpha a] [--content_size n] [--style_size n] ...

Everything between image decode and image encode stays on the GPU (`WCT.stylize`); additive flags:
  --precision {h2,tf32,fp32} conv engine (default h2: fp32-accurate tensor-core convs)   --weights_root DIR  (default ../trained_models)
  --gpus N                  shard every image into N vertical strips, one process per GPU (re-launches itself under torchrun)
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
    ("--precision", dict(type=str, default="h2", choices=["h2", "tf32", "fp32"])),
    ("--gpus", dict(type=int, default=1, help="strip-shard each image over this many GPUs of the node (one process per GPU)")),
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


def _relaunch_under_torchrun(args, argv):
    """--gpus N from a plain `python WCT.py ...`: start N ranks of this script (one process per GPU, NCCL)"""
    import subprocess
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus), "--master-addr", "127.0.0.1",
           "--master-port", os.environ.get("MASTER_PORT", "29517"), os.path.abspath(__file__)] + list(sys.argv[1:] if argv is None else argv)
    return subprocess.call(cmd)


def main(argv=None):
    args = parse(argv)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        sys.exit(_relaunch_under_torchrun(args, argv))
    import torchvision.utils as vutils

    import collaborative_distillation_b200 as P
    from data_loader import Dataset

    rank, grp = 0, None
    if world > 1:
        import torch.distributed as dist
        from collaborative_distillation_b200 import parallel
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        rank = dist.get_rank()
        grp = parallel.StripGroup()
        grp.use_graph = False      # every pair is seen once: a graph per pair would run the path twice

    os.makedirs(args.outf, exist_ok=True)
    log = LogPrinter(args.debug, os.path.join(args.outf, "log_%s_%s.txt" % (args.log_mark, args.mode))) if rank == 0 else (lambda sth: None)
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
    # every content x style pair is seen once: capturing a CUDA graph per pair would run the path twice and pin one
    # activation pool per input shape (graphs pay off for repeated shapes: bench.py, servers)
    wct.use_graph = False
    wct.dist = grp
    log("Number of content-style pairs: %s" % len(loader))
    total, n = 0.0, 0
    # style path -> per-stage style statistics/eigensystems: a style is encoded once, not once per pair.  Keyed on the real
    # file path (names like "a+b.jpg" / "van.1.jpg" would collide if parsed back from the output name); bounded.
    import collections
    style_cache = collections.OrderedDict()
    max_styles = 8
    for i, (cImg, sImg, imname) in enumerate(loader):
        imname = imname[0]
        log("\n" + "*" * 30 + ' #%s: Transferring "%s"' % (i, imname))
        start = time.time()
        if grp is not None:
            # strips: every rank decodes the pair and keeps its own columns; rank 0 gathers the stylized strips for saving
            from collaborative_distillation_b200 import parallel
            c_own = grp.own_slice(cImg, parallel.strip_cuts(cImg.shape[-1], world), rank).cuda()
            s_own = grp.own_slice(sImg, parallel.strip_cuts(sImg.shape[-1], world), rank).cuda()
            own = grp.stylize(wct.style_transfer_stage, args.mode or "original", c_own, s_own, alpha=args.alpha, num_run=args.num_run)
            out = grp.gather_strips(own)
            if rank != 0:
                continue
        else:
            skey = dataset.paths(i)[1]
            if skey not in style_cache:
                style_cache[skey] = wct.prepare_style(sImg.cuda())
                while len(style_cache) > max_styles:
                    style_cache.popitem(last=False)
            else:
                style_cache.move_to_end(skey)
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
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
