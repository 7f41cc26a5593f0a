#!/usr/bin/env python
"""Build libwctb.so in-tree with nvcc for sm_100a (no torch headers; plain C ABI)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SOURCES = ["conv_fp32.cu", "conv_umma.cu", "conv_h2.cu", "conv_h2_fused.cu", "wct_transform.cu", "gram_ring.cu", "gram_alt.cu", "image_io.cu", "halo.cu", "whiten_ns.cu"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
         "-Xcompiler", "-fPIC", "-I", os.path.join(ROOT, "include"), "-I", HERE]


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(verbose=False, force=False):
    out = os.path.join(HERE, "libwctb.so")
    hdrs = [os.path.join(HERE, "common.cuh"), os.path.join(ROOT, "include", "wctb.h")]
    hdrs += [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(".cuh")]
    objs, cmds = [], []
    for src in SOURCES:
        s = os.path.join(HERE, src)
        o = os.path.join(HERE, src.replace(".cu", ".o"))
        if force or _stale(o, [s] + hdrs):
            cmds.append([NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o])
        objs.append(o)
    if cmds:   # translation units are independent: compile them side by side (wct_transform.cu alone takes ~2 min)
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=min(len(cmds), os.cpu_count() or 1)) as pool:
            list(pool.map(subprocess.check_call, cmds))
    if force or _stale(out, objs):
        subprocess.check_call([NVCC, "-shared", "-o", out] + objs + ["-lcudart"])
    build_io(force=force)
    return out


def build_io(force=False):
    """libwctb_io.so: the nvJPEG binding of include/wctb_io.h (host C++ only; separate so that libwctb.so stays
    free of library dependencies beyond cudart)."""
    out = os.path.join(HERE, "libwctb_io.so")
    src = os.path.join(HERE, "jpeg_io.cpp")
    if force or _stale(out, [src, os.path.join(ROOT, "include", "wctb_io.h")]):
        subprocess.check_call([NVCC, "-O2", "-std=c++17", "-Xcompiler", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                               "-o", out, src, "-lnvjpeg", "-lcudart"])
    return out


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
