"""torch-CPU fp32 executor for the .tflite graphs (five dense ones + the sparse full-range detector) -- oracle side.

Test infrastructure; see ``oracle/__init__.py``.  Restates what
``interpreter.invoke()`` computes at face_detection.rs:235, face_landmark.rs:265
and iris_landmark.rs:203 using the TFLite *float reference* semantics of the
11 builtin ops involved:

* CONV_2D weights OHWI, DEPTHWISE_CONV_2D weights [1,kh,kw,C], fp32 accumulate;
* SAME padding: out=ceil(in/stride), total=max(0,(out-1)*stride+k-in),
  before=total//2, after=total-before (asymmetric for stride 2);
* PAD constant 0, ADD same-shape, RELU, PRELU (alpha [1,1,C]),
* DEQUANTIZE f16->f32 exact widening, MAX_POOL_2D, RESHAPE, CONCATENATION,
* RESIZE_BILINEAR with half_pixel_centers (== torch interpolate
  align_corners=False).

All tensors are kept NHWC like TFLite; the batch dimension (1 in the files) is
generalised to B.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import tflite_reader as T


def _same_pad(in_size, k, stride):
    out = -(-in_size // stride)
    total = max(0, (out - 1) * stride + k - in_size)
    return total // 2, total - total // 2


def _nchw(x):
    return x.permute(0, 3, 1, 2)


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


class GraphExecutor:
    def __init__(self, path: str, dtype=torch.float32):
        self.model = T.load(path)
        self.dtype = dtype
        m = self.model
        self.const = {}
        for t in m.tensors:
            if t.data is not None:
                self.const[t.index] = t.data
        # fold DENSIFY (sparse -> dense, same type) and DEQUANTIZE (f16 -> f32) of constants
        self.ops = []
        for op in m.ops:
            if op.code == T.DENSIFY and op.inputs[0] in self.const:
                self.const[op.outputs[0]] = T.densify(m.tensors[op.inputs[0]])
            elif op.code == T.DEQUANTIZE and op.inputs[0] in self.const:
                self.const[op.outputs[0]] = self.const[op.inputs[0]].astype(np.float32)
            else:
                self.ops.append(op)
        self._tconst = {}
        self.input_shape = m.tensors[m.inputs[0]].shape
        self.output_shapes = [m.tensors[i].shape for i in m.outputs]

    def _c(self, idx):
        if idx not in self._tconst:
            a = self.const[idx]
            if a.dtype == np.int32:
                self._tconst[idx] = torch.from_numpy(a.copy())
            else:
                self._tconst[idx] = torch.from_numpy(a.astype(np.float32)).to(self.dtype)
        return self._tconst[idx]

    @torch.no_grad()
    def run(self, x, keep=None):
        """x: [B,H,W,3] float array/tensor (NHWC). Returns list of numpy outputs in
        graph output order (== ``interpreter.outputs()`` order).  ``keep``: optional
        iterable of tensor indices whose values are also returned (dict) for
        layer-wise debugging."""
        m = self.model
        x = torch.as_tensor(np.asarray(x)).to(self.dtype)
        assert list(x.shape[1:]) == self.input_shape[1:], (x.shape, self.input_shape)
        B = x.shape[0]
        vals = {m.inputs[0]: x}
        kept = {}

        def get(i):
            return vals[i] if i in vals else self._c(i)

        def fused(y, act):
            # fused_activation_function: 0 NONE, 1 RELU (the only ones the reference's models use)
            assert act in (0, 1), act
            return torch.relu(y) if act == 1 else y

        for op in self.ops:
            o = op.opts
            c = op.code
            if c == T.CONV_2D:
                inp, w, b = get(op.inputs[0]), get(op.inputs[1]), get(op.inputs[2])
                kh, kw = w.shape[1], w.shape[2]
                xi = _nchw(inp)
                if o["padding"] == 0:
                    pt, pb = _same_pad(inp.shape[1], kh, o["stride_h"])
                    pl, pr = _same_pad(inp.shape[2], kw, o["stride_w"])
                    xi = F.pad(xi, (pl, pr, pt, pb))
                y = F.conv2d(xi, w.permute(0, 3, 1, 2).contiguous(), b,
                             stride=(o["stride_h"], o["stride_w"]))
                r = fused(_nhwc(y), o["act"])
            elif c == T.DEPTHWISE_CONV_2D:
                inp, w, b = get(op.inputs[0]), get(op.inputs[1]), get(op.inputs[2])
                kh, kw, C = w.shape[1], w.shape[2], w.shape[3]
                assert o["depth_multiplier"] == 1
                xi = _nchw(inp)
                if o["padding"] == 0:
                    pt, pb = _same_pad(inp.shape[1], kh, o["stride_h"])
                    pl, pr = _same_pad(inp.shape[2], kw, o["stride_w"])
                    xi = F.pad(xi, (pl, pr, pt, pb))
                wt = w[0].permute(2, 0, 1).unsqueeze(1).contiguous()  # [C,1,kh,kw]
                y = F.conv2d(xi, wt, b, stride=(o["stride_h"], o["stride_w"]), groups=C)
                r = fused(_nhwc(y), o["act"])
            elif c == T.MAX_POOL_2D:
                inp = get(op.inputs[0])
                xi = _nchw(inp)
                if o["padding"] == 0:
                    pt, pb = _same_pad(inp.shape[1], o["filter_h"], o["stride_h"])
                    pl, pr = _same_pad(inp.shape[2], o["filter_w"], o["stride_w"])
                    if pt or pb or pl or pr:
                        xi = F.pad(xi, (pl, pr, pt, pb), value=float("-inf"))
                y = F.max_pool2d(xi, (o["filter_h"], o["filter_w"]), (o["stride_h"], o["stride_w"]))
                r = _nhwc(y)
            elif c == T.ADD:
                a, b2 = get(op.inputs[0]), get(op.inputs[1])
                assert a.shape == b2.shape
                r = fused(a + b2, o["act"])
            elif c == T.RELU:
                r = torch.relu(get(op.inputs[0]))
            elif c == T.PRELU:
                a, alpha = get(op.inputs[0]), get(op.inputs[1])
                r = torch.where(a >= 0, a, a * alpha.reshape(1, 1, 1, -1))
            elif c == T.PAD:
                a = get(op.inputs[0])
                p = self.const[op.inputs[1]].reshape(-1, 2)
                assert p.shape[0] == a.dim()
                flat = []
                for d in range(a.dim() - 1, -1, -1):
                    flat += [int(p[d, 0]), int(p[d, 1])]
                r = F.pad(a, flat)
            elif c == T.RESHAPE:
                a = get(op.inputs[0])
                ns = list(o.get("new_shape") or [])
                if not ns and len(op.inputs) > 1:
                    ns = [int(v) for v in self.const[op.inputs[1]].reshape(-1)]
                assert ns[0] == 1
                ns[0] = B
                r = a.reshape(ns)
            elif c == T.CONCATENATION:
                r = torch.cat([get(i) for i in op.inputs], dim=o["axis"])
            elif c == T.RESIZE_BILINEAR:
                a = get(op.inputs[0])
                size = [int(v) for v in self.const[op.inputs[1]].reshape(-1)]
                assert o["half_pixel_centers"] == 1 and o["align_corners"] == 0
                r = _nhwc(F.interpolate(_nchw(a), size=size, mode="bilinear", align_corners=False))
            elif c == T.DEPTH_TO_SPACE:
                # out[n, h*b + i, w*b + j, c] = in[n, h, w, (i*b + j)*Cout + c]  (tensorflow/lite/kernels/internal/reference/depth_to_space.h)
                a = get(op.inputs[0])
                bs = o["block_size"]
                n_, h_, w_, c_ = a.shape
                co = c_ // (bs * bs)
                r = a.reshape(n_, h_, w_, bs, bs, co).permute(0, 1, 3, 2, 4, 5).reshape(n_, h_ * bs, w_ * bs, co).contiguous()
            else:
                raise NotImplementedError(op.name)
            vals[op.outputs[0]] = r
            if keep is not None and op.outputs[0] in keep:
                kept[op.outputs[0]] = r.numpy().copy()
        outs = [vals[i].to(torch.float32).numpy() if vals[i].dtype != torch.float64
                else vals[i].numpy() for i in m.outputs]
        return (outs, kept) if keep is not None else outs
