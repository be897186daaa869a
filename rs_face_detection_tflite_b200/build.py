"""In-tree build of libfdl_b200.so (nvcc, sm_100a only).

``python -m rs_face_detection_tflite_b200.build`` or ``build()``; the shared library lands next to
this file (``rs_face_detection_tflite_b200/libfdl_b200.so``) so that it travels with the repo snapshot.
nvcc cross-compiles without a GPU.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libfdl_b200.so")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden",
          "-Xcompiler", "-Wall", "-Xcompiler", "-Wno-unused-function"]
# translation unit -> extra flags
SOURCES = {
    "tflite_model.cc": [],
    "plan.cc": [],
    "net.cu": [],
    "net_kernels.cu": [],
    "mma_kernels.cu": [],
    "block_ws_kernel.cu": [],
    "stem_kernel.cu": [],
    "stem_tc_kernel.cu": [],
    "conv_tc_kernel.cu": [],
    "pw_kernel.cu": [],
    "pw_tc_kernel.cu": [],
    "chain_plan.cc": [],
    "chain_kernel.cu": [],
    # the glue arithmetic must not contract a*b+c into FMA (the reference's scalar Rust never does)
    "prepost_kernels.cu": ["-fmad=false"],
    "jpeg_kernels.cu": [],
    "jpeg_decode.cu": [],
    "fdl_api.cu": [],
    "pipeline.cu": [],
    "pool.cu": [],
    "render.cu": ["-fmad=false"],
}


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the B200 library cannot be built")


def _deps_mtime() -> float:
    m = 0.0
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in os.listdir(root):
            if f.endswith((".h", ".cuh", ".hpp")):
                m = max(m, os.path.getmtime(os.path.join(root, f)))
    return m


# Variant builds for A/B measurements (never the product library): name -> extra nvcc flags.  They land in build_<name>/ and
# libfdl_b200_<name>.so and are loaded with FDL_LIB=<path>.
VARIANTS = {"trace": ["-DFDL_WS_TRACE"], "trace_c32h32": ["-DFDL_WS_TRACE", "-DFDL_TRACE_C=32", "-DFDL_TRACE_H=32"],
            "trace_c64h16": ["-DFDL_WS_TRACE", "-DFDL_TRACE_C=64", "-DFDL_TRACE_H=16"]}


def build(force: bool = False, verbose: bool = False, variant: str | None = None) -> str:
    global OBJ, LIB
    nvcc = _nvcc()
    extra_all = []
    if variant:
        extra_all = VARIANTS[variant]
        OBJ = os.path.join(HERE, "build_" + variant)
        LIB = os.path.join(HERE, "libfdl_b200_%s.so" % variant)
    os.makedirs(OBJ, exist_ok=True)
    hdr = max(_deps_mtime(), os.path.getmtime(os.path.abspath(__file__)))
    jobs = []
    objs = []
    for src, extra in SOURCES.items():
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        op = os.path.join(OBJ, src.rsplit(".", 1)[0] + ".o")
        objs.append(op)
        if force or not os.path.exists(op) or os.path.getmtime(op) < max(os.path.getmtime(sp), hdr):
            cmd = [nvcc, "-c", sp, "-o", op, "-x", "cu"] + ARCH + COMMON + extra + extra_all
            if verbose:
                cmd += ["-Xptxas", "-v"]
            jobs.append(cmd)

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for cmd, r in ex.map(run, jobs):
                if verbose or r.returncode != 0:
                    sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
                if r.returncode != 0:
                    raise RuntimeError("nvcc failed for " + cmd[2])
    if jobs or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [nvcc, "-shared", "-o", LIB] + objs + ARCH + ["-Xcompiler", "-fPIC"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv,
                variant=sys.argv[sys.argv.index("--variant") + 1] if "--variant" in sys.argv else None))
