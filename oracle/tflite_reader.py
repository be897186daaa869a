"""Minimal TFLite (schema v3) flatbuffer reader -- oracle side (numpy only).

Test infrastructure; see ``oracle/__init__.py``.  The product has its own,
independent C++ reader (``csrc/tflite_model.cc``); the two are cross-checked in
``tests/test_tflite_reader.py``.

Follows the public TFLite schema (tensorflow/lite/schema/schema.fbs, v3) for the
subset the reference's models use; the reference loads the same files through
``FlatBufferModel::build_from_file`` (face_detection.rs:188,
face_landmark.rs:216, iris_landmark.rs:150).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

import numpy as np

# builtin operator codes used by the five dense graphs
ADD, CONCATENATION, CONV_2D, DEPTHWISE_CONV_2D, DEQUANTIZE = 0, 2, 3, 4, 6
MAX_POOL_2D, RELU, RESHAPE, RESIZE_BILINEAR, PAD, PRELU = 17, 19, 22, 23, 34, 54
# the sparse full-range detector adds two
DEPTH_TO_SPACE, DENSIFY = 5, 124
OP_NAMES = {
    ADD: "ADD", CONCATENATION: "CONCATENATION", CONV_2D: "CONV_2D",
    DEPTHWISE_CONV_2D: "DEPTHWISE_CONV_2D", DEQUANTIZE: "DEQUANTIZE",
    MAX_POOL_2D: "MAX_POOL_2D", RELU: "RELU", RESHAPE: "RESHAPE",
    RESIZE_BILINEAR: "RESIZE_BILINEAR", PAD: "PAD", PRELU: "PRELU",
    5: "DEPTH_TO_SPACE", 124: "DENSIFY",
}
TENSOR_TYPES = {0: np.float32, 1: np.float16, 2: np.int32, 3: np.uint8, 4: np.int64}


class _FB:
    """Tiny flatbuffer cursor: tables, vectors, scalars by field id."""

    def __init__(self, buf: bytes):
        self.b = buf

    def u8(self, p): return self.b[p]
    def i8(self, p): return struct.unpack_from("<b", self.b, p)[0]
    def u16(self, p): return struct.unpack_from("<H", self.b, p)[0]
    def i32(self, p): return struct.unpack_from("<i", self.b, p)[0]
    def u32(self, p): return struct.unpack_from("<I", self.b, p)[0]

    def root(self):
        return self.u32(0)

    def field(self, table, fid):
        """Absolute position of field ``fid`` in ``table`` or None if absent."""
        vt = table - self.i32(table)
        vsize = self.u16(vt)
        slot = 4 + 2 * fid
        if slot >= vsize:
            return None
        off = self.u16(vt + slot)
        return table + off if off else None

    def scalar(self, table, fid, kind, default=0):
        p = self.field(table, fid)
        if p is None:
            return default
        return getattr(self, kind)(p)

    def indirect(self, p):
        return p + self.u32(p)

    def table(self, table, fid):
        p = self.field(table, fid)
        return None if p is None else self.indirect(p)

    def vector(self, table, fid):
        """(start, length) of a vector field, or (None, 0)."""
        p = self.field(table, fid)
        if p is None:
            return None, 0
        v = self.indirect(p)
        return v + 4, self.u32(v)

    def vec_i32(self, table, fid):
        s, n = self.vector(table, fid)
        if s is None:
            return []
        return list(struct.unpack_from("<%di" % n, self.b, s))

    def vec_tables(self, table, fid):
        s, n = self.vector(table, fid)
        return [self.indirect(s + 4 * i) for i in range(n)] if s is not None else []

    def string(self, table, fid):
        s, n = self.vector(table, fid)
        return "" if s is None else self.b[s:s + n].decode("utf-8", "replace")


@dataclass
class Tensor:
    index: int
    name: str
    shape: list
    dtype: type
    buffer: int
    data: np.ndarray | None = None  # constant payload, if any (the stored VALUES only when `sparsity` is set)
    sparsity: dict | None = None    # SparsityParameters: traversal_order, block_map, dims=[(format, dense_size, segments, indices)]


@dataclass
class Op:
    code: int
    inputs: list
    outputs: list
    opts: dict = field(default_factory=dict)

    @property
    def name(self):
        return OP_NAMES.get(self.code, "OP_%d" % self.code)


@dataclass
class Model:
    version: int
    tensors: list
    ops: list
    inputs: list
    outputs: list


def _parse_options(fb: _FB, code: int, t) -> dict:
    if t is None:
        return {}
    if code == CONV_2D:
        return dict(padding=fb.scalar(t, 0, "i8"), stride_w=fb.scalar(t, 1, "i32"),
                    stride_h=fb.scalar(t, 2, "i32"), act=fb.scalar(t, 3, "i8"),
                    dil_w=fb.scalar(t, 4, "i32", 1), dil_h=fb.scalar(t, 5, "i32", 1))
    if code == DEPTHWISE_CONV_2D:
        return dict(padding=fb.scalar(t, 0, "i8"), stride_w=fb.scalar(t, 1, "i32"),
                    stride_h=fb.scalar(t, 2, "i32"), depth_multiplier=fb.scalar(t, 3, "i32"),
                    act=fb.scalar(t, 4, "i8"), dil_w=fb.scalar(t, 5, "i32", 1),
                    dil_h=fb.scalar(t, 6, "i32", 1))
    if code == MAX_POOL_2D:
        return dict(padding=fb.scalar(t, 0, "i8"), stride_w=fb.scalar(t, 1, "i32"),
                    stride_h=fb.scalar(t, 2, "i32"), filter_w=fb.scalar(t, 3, "i32"),
                    filter_h=fb.scalar(t, 4, "i32"), act=fb.scalar(t, 5, "i8"))
    if code == ADD:
        return dict(act=fb.scalar(t, 0, "i8"))
    if code == CONCATENATION:
        return dict(axis=fb.scalar(t, 0, "i32"), act=fb.scalar(t, 1, "i8"))
    if code == RESHAPE:
        return dict(new_shape=fb.vec_i32(t, 0))
    if code == RESIZE_BILINEAR:
        return dict(align_corners=fb.scalar(t, 2, "u8"), half_pixel_centers=fb.scalar(t, 3, "u8"))
    if code == DEPTH_TO_SPACE:
        return dict(block_size=fb.scalar(t, 0, "i32"))
    return {}


def _sparse_index_vector(fb: _FB, table, type_fid, value_fid):
    """SparseIndexVector union (schema.fbs): 1 Int32Vector, 2 Uint16Vector, 3 Uint8Vector, each {0: values}."""
    kind = fb.scalar(table, type_fid, "u8")
    tb = fb.table(table, value_fid)
    if kind == 0 or tb is None:
        return None
    s, n = fb.vector(tb, 0)
    fmt = {1: "<%di", 2: "<%dH", 3: "<%dB"}[kind]
    return np.array(struct.unpack_from(fmt % n, fb.b, s), np.int64)


def _parse_sparsity(fb: _FB, t):
    """Tensor.sparsity (field 6): SparsityParameters{0: traversal_order, 1: block_map, 2: dim_metadata[]},
    DimensionMetadata{0: format (0 DENSE, 1 SPARSE_CSR), 1: dense_size, 2/3: array_segments, 4/5: array_indices}."""
    sp = fb.table(t, 6)
    if sp is None:
        return None
    dims = []
    for d in fb.vec_tables(sp, 2):
        dims.append((fb.scalar(d, 0, "i8"), fb.scalar(d, 1, "i32"), _sparse_index_vector(fb, d, 2, 3), _sparse_index_vector(fb, d, 4, 5)))
    return dict(traversal_order=fb.vec_i32(sp, 0), block_map=fb.vec_i32(sp, 1), dims=dims)


def densify(t: "Tensor") -> np.ndarray:
    """What the DENSIFY op computes (tensorflow/lite/kernels/densify.cc -> FormatConverter::SparseToDense) for the layout
    the reference's sparse model uses: identity traversal order, no block map, every dimension DENSE except the last,
    which is SPARSE_CSR (segments over the flattened outer dimensions, indices = positions in the last dimension)."""
    sp = t.sparsity
    rank = len(t.shape)
    if sp["traversal_order"] != list(range(rank)) or sp["block_map"] or len(sp["dims"]) != rank:
        raise NotImplementedError("sparse layout other than row-major CSR on the last dimension")
    for d, (fmt, size, _, _) in enumerate(sp["dims"][:-1]):
        if fmt != 0 or size != t.shape[d]:
            raise NotImplementedError("sparse layout other than row-major CSR on the last dimension")
    fmt, _, seg, idx = sp["dims"][-1]
    if fmt != 1 or seg is None or idx is None:
        raise NotImplementedError("last dimension is not SPARSE_CSR")
    rows = int(np.prod(t.shape[:-1]))
    values = t.data.reshape(-1)
    if len(seg) != rows + 1 or seg[-1] != len(idx) or len(values) < len(idx):
        raise ValueError("inconsistent sparsity metadata")
    out = np.zeros((rows, t.shape[-1]), t.dtype)
    for r in range(rows):
        a, b = int(seg[r]), int(seg[r + 1])
        out[r, idx[a:b]] = values[a:b]
    return out.reshape(t.shape)


def load(path: str) -> Model:
    with open(path, "rb") as f:
        buf = f.read()
    fb = _FB(buf)
    root = fb.root()
    version = fb.scalar(root, 0, "u32")
    codes = []
    for oc in fb.vec_tables(root, 1):
        dep = fb.scalar(oc, 0, "i8")
        new = fb.scalar(oc, 3, "i32")
        codes.append(max(dep, new))
    buffers = []
    for b in fb.vec_tables(root, 4):
        s, n = fb.vector(b, 0)
        buffers.append((s, n))
    sub = fb.vec_tables(root, 2)[0]
    tensors = []
    for i, t in enumerate(fb.vec_tables(sub, 0)):
        shape = fb.vec_i32(t, 0)
        ttype = fb.scalar(t, 1, "i8")
        bidx = fb.scalar(t, 2, "u32")
        dtype = TENSOR_TYPES[ttype]
        data = None
        s, n = buffers[bidx] if bidx < len(buffers) else (None, 0)
        sparsity = _parse_sparsity(fb, t)
        if s is not None and n > 0:
            data = np.frombuffer(buf, dtype=dtype, count=n // np.dtype(dtype).itemsize, offset=s).copy()
            if sparsity is None:
                data = data.reshape(shape)
        tensors.append(Tensor(i, fb.string(t, 3), shape, dtype, bidx, data, sparsity))
    ops = []
    for o in fb.vec_tables(sub, 3):
        code = codes[fb.scalar(o, 0, "u32")]
        ops.append(Op(code, fb.vec_i32(o, 1), fb.vec_i32(o, 2),
                      _parse_options(fb, code, fb.table(o, 4))))
    return Model(version, tensors, ops, fb.vec_i32(sub, 1), fb.vec_i32(sub, 2))


def summary(m: Model) -> str:
    from collections import Counter
    c = Counter(op.name for op in m.ops)
    lines = ["version %d, %d tensors, %d ops" % (m.version, len(m.tensors), len(m.ops))]
    lines.append("inputs  " + ", ".join("%s%s" % (m.tensors[i].name, m.tensors[i].shape) for i in m.inputs))
    lines.append("outputs " + ", ".join("%s%s" % (m.tensors[i].name, m.tensors[i].shape) for i in m.outputs))
    lines.append(", ".join("%s %d" % kv for kv in sorted(c.items())))
    return "\n".join(lines)


if __name__ == "__main__":
    import sys
    for p in sys.argv[1:]:
        print(p)
        print(summary(load(p)))
