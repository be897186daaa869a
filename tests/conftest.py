"""pytest configuration: the `gpu` marker and shared fixtures.

`-m "not gpu"`: oracle vs golden vectors, host logic (flatbuffer reader, planner, glue arithmetic compiled
for the host), C-ABI load/export checks.  `-m gpu`: parity of the CUDA path with the oracle through the C ABI.
"""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

MODELS = os.path.join(ROOT, "models")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


@pytest.fixture(scope="session")
def fdl():
    """The product package with its native library built (never a fallback)."""
    from rs_face_detection_tflite_b200 import build as _build
    _build.build()
    import rs_face_detection_tflite_b200 as m
    return m


@pytest.fixture(scope="session")
def gpu(fdl):
    if fdl.device_count() < 1:
        pytest.fail("a test marked gpu ran on a box without a CUDA device")
    return 0


@pytest.fixture(scope="session")
def man():
    import synth_frames
    return synth_frames.load_rgb("man.jpg")


@pytest.fixture(scope="session")
def oracle_pipeline():
    from oracle import glue, pipeline
    return {m: pipeline.Pipeline(m, MODELS) for m in (glue.BACK_CAMERA,)}


def rng(seed):
    return np.random.default_rng(seed)
