"""Parity step 2: the planned CUDA graphs vs the oracle's TFLite-semantics executor on
identical input tensors.  Replaces `interpreter.invoke()` at face_detection.rs:235, face_landmark.rs:265,
iris_landmark.rs:203.

Tolerances (fp32 on both sides, different summation order): regressors atol 1e-3*S px, logits
atol 1e-3 + rtol 1e-4, landmark / iris raw outputs atol 0.05 (tensor-pixel units).
"""
import os

import numpy as np
import pytest

from conftest import MODELS, rng

pytestmark = pytest.mark.gpu

NETS = {
    "face_detection_short_range": 128,
    "face_detection_front": 128,
    "face_detection_back": 256,
    "face_detection_full_range": 192,
    "face_landmark": 192,
    "iris_landmark": 64,
    "face_detection_full_range_sparse": 192,     # DENSIFY / spatial PAD / fused RELU / DEPTH_TO_SPACE
}


def _inputs(name, size, batch, seed, man):
    """Half random tensors in the model's input range, half real crops of man.jpg."""
    from oracle import glue
    lo = -1.0 if "detection" in name else 0.0
    r = rng(seed)
    x = r.uniform(lo, 1.0, (batch, size, size, 3)).astype(np.float32)
    it = glue.image_to_tensor(man, None, (size, size), True, (lo, 1.0), False)
    x[0] = it.tensor_data
    return x


def _check(name, ours, ref, size):
    for i, (a, b) in enumerate(zip(ours, ref)):
        a = a.reshape(b.shape)
        if "detection" in name and i == 0:
            np.testing.assert_allclose(a, b, atol=1e-3 * size, rtol=0)
        elif "detection" in name:
            np.testing.assert_allclose(a, b, atol=1e-3, rtol=1e-4)
        else:
            np.testing.assert_allclose(a, b, atol=0.05, rtol=1e-4)


@pytest.mark.parametrize("name", list(NETS))
@pytest.mark.parametrize("mode", [0, 1, 2])
def test_forward_matches_oracle(fdl, gpu, man, name, mode):
    from oracle.graph_exec import GraphExecutor
    size = NETS[name]
    path = os.path.join(MODELS, name + ".tflite")
    net = fdl.Net(path, device=gpu)
    net.set_mode(mode)
    ref = GraphExecutor(path)
    for batch, seed in ((1, 0), (5, 1)):
        x = _inputs(name, size, batch, seed, man)
        _check(name, net.forward(x), ref.run(x), size)
    net.close()


@pytest.mark.parametrize("name,batch,pick", [("face_detection_back", 7, 3), ("face_landmark", 41, 38), ("iris_landmark", 41, 39)])
def test_batch_independence(fdl, gpu, man, name, batch, pick):
    """Items of a batch do not influence each other and results do not depend on the batch size: the same item alone, in a small and
    in a large batch gives the same BITS.  (Kernels are chosen by layer shape, never by batch size; the tail chain groups 2 eyes /
    3 faces per CTA, so the picked item also changes its position in a group and the group changes from full to partial.)"""
    size = NETS[name]
    net = fdl.Net(os.path.join(MODELS, name + ".tflite"), device=gpu)
    x = _inputs(name, size, batch, 3, man)
    full = net.forward(x)
    one = net.forward(x[pick:pick + 1])
    some = net.forward(x[pick - 1:pick + 2])
    for a, b, c in zip(full, one, some):
        np.testing.assert_array_equal(a[pick:pick + 1], b)
        np.testing.assert_array_equal(a[pick - 1:pick + 2], c)
    net.close()


def test_f64_reference_bounds_the_oracle(man):
    """The oracle's own fp32 rounding, bounded by re-running the graph in fp64 (oracle-of-an-oracle check)."""
    import torch
    from oracle.graph_exec import GraphExecutor
    from oracle import glue
    path = os.path.join(MODELS, "face_detection_back.tflite")
    it = glue.image_to_tensor(man, None, (256, 256), True, (-1.0, 1.0), False)
    a = GraphExecutor(path).run(it.tensor_data[None])
    b = GraphExecutor(path, dtype=torch.float64).run(it.tensor_data[None])
    assert np.abs(a[0] - b[0]).max() < 1e-2
    keep = np.abs(b[1]) < 20
    assert np.abs(a[1] - b[1])[keep].max() < 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(NETS))
def test_pipelined_block_kernel_matches_the_serial_one(fdl, gpu, man, name):
    """block_ws_kernel (mode 1: warp-specialised, several tiles in flight per CTA, bias added on the tensor cores) against
    blaze_block_tc_kernel (mode 2) on a batch large enough to wrap every mbarrier ring many times: same results up to the
    rounding of the bias term (2^-22 relative), and bit-identical from run to run (any race would show up here)."""
    size = NETS[name]
    net = fdl.Net(os.path.join(MODELS, name + ".tflite"), device=gpu)
    x = _inputs(name, size, 96 if size <= 192 else 48, 7, man)
    net.set_mode(2)
    ref = net.forward(x)
    net.set_mode(1)
    first = net.forward(x)
    for a, b in zip(first, ref):
        np.testing.assert_allclose(a, b, atol=2e-4 * max(1.0, float(np.abs(b).max())), rtol=0)
    for _ in range(3):
        for a, b in zip(net.forward(x), first):
            np.testing.assert_array_equal(a, b)
    net.close()
