"""Builds tests/hostcheck/hostcheck.cc (g++) and exposes it through ctypes.  Test infrastructure only."""
import ctypes as C
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "_hostcheck.so")


def load():
    src = os.path.join(HERE, "hostcheck.cc")
    csrc = os.path.join(HERE, "..", "..", "rs_face_detection_tflite_b200", "csrc")
    newest = max(os.path.getmtime(f) for f in (src, os.path.join(csrc, "glue_math.h"), os.path.join(csrc, "jpeg_math.h"), os.path.join(csrc, "jpeg_parse.h")))
    if not os.path.exists(SO) or os.path.getmtime(SO) < newest:
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-o", SO, src])
    return C.CDLL(SO)
