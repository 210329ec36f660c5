"""ORACLE (test infrastructure, not product code).

Minimal TFLite flatbuffer (schema v3) reader written against the public TFLite
schema; no `flatbuffers`/`tensorflow` dependency.  Used only by tests, by
`__graft_entry__.smoke()` and by `bench.py`'s cpu_baseline / reference arm.

The reference loads the same files through flutter_litert's
`Interpreter.fromBuffer` (reference: lib/src/models/face_detection_model.dart:156-191,
lib/src/models/face_landmark.dart:148-191).  The schema field slots used here are
listed in SURVEY.md section 7.4.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np

# builtin operator codes (tensorflow/lite/schema/schema.fbs, BuiltinOperator)
OP_ADD = 0
OP_AVERAGE_POOL_2D = 1
OP_CONCATENATION = 2
OP_CONV_2D = 3
OP_DEPTHWISE_CONV_2D = 4
OP_DEQUANTIZE = 6
OP_MAX_POOL_2D = 17
OP_RELU = 19
OP_RESHAPE = 22
OP_RESIZE_BILINEAR = 23
OP_PAD = 34
OP_PRELU = 54

OP_NAMES = {
    0: "ADD", 1: "AVERAGE_POOL_2D", 2: "CONCATENATION", 3: "CONV_2D", 4: "DEPTHWISE_CONV_2D",
    6: "DEQUANTIZE", 17: "MAX_POOL_2D", 19: "RELU", 22: "RESHAPE", 23: "RESIZE_BILINEAR",
    34: "PAD", 54: "PRELU",
}

TENSOR_F32, TENSOR_F16, TENSOR_I32 = 0, 1, 2


class _FB:
    """Raw flatbuffer accessors."""

    def __init__(self, buf: bytes):
        self.b = memoryview(buf)

    def u8(self, o):
        return self.b[o]

    def i8(self, o):
        return struct.unpack_from("<b", self.b, o)[0]

    def u16(self, o):
        return struct.unpack_from("<H", self.b, o)[0]

    def i32(self, o):
        return struct.unpack_from("<i", self.b, o)[0]

    def u32(self, o):
        return struct.unpack_from("<I", self.b, o)[0]

    def indirect(self, o):
        return o + self.u32(o)

    def field(self, table, slot) -> int:
        """Absolute position of field `slot` of `table`, or 0 if absent."""
        vt = table - self.i32(table)
        vsize = self.u16(vt)
        e = 4 + 2 * slot
        if e >= vsize:
            return 0
        off = self.u16(vt + e)
        return table + off if off else 0

    def vec(self, table, slot):
        """(start, length) of a vector field, or (0, 0)."""
        p = self.field(table, slot)
        if not p:
            return 0, 0
        v = self.indirect(p)
        return v + 4, self.u32(v)

    def vec_tables(self, table, slot):
        s, n = self.vec(table, slot)
        return [self.indirect(s + 4 * i) for i in range(n)]

    def vec_i32(self, table, slot):
        s, n = self.vec(table, slot)
        return list(struct.unpack_from("<%di" % n, self.b, s)) if n else []

    def string(self, table, slot):
        s, n = self.vec(table, slot)
        return bytes(self.b[s:s + n]).decode("utf-8", "replace") if s else ""

    def scalar(self, table, slot, fmt, default=0):
        p = self.field(table, slot)
        return struct.unpack_from(fmt, self.b, p)[0] if p else default


@dataclass
class Tensor:
    index: int
    name: str
    shape: List[int]
    dtype: int
    buffer: int
    data: Optional[np.ndarray] = None  # constant payload in its stored dtype


@dataclass
class Op:
    code: int
    inputs: List[int]
    outputs: List[int]
    opts: Dict[str, int] = field(default_factory=dict)

    @property
    def name(self):
        return OP_NAMES.get(self.code, "OP%d" % self.code)


@dataclass
class Model:
    tensors: List[Tensor]
    ops: List[Op]
    inputs: List[int]
    outputs: List[int]


def read_tflite(path_or_bytes) -> Model:
    if isinstance(path_or_bytes, (bytes, bytearray, memoryview)):
        buf = bytes(path_or_bytes)
    else:
        with open(path_or_bytes, "rb") as f:
            buf = f.read()
    fb = _FB(buf)
    root = fb.indirect(0)
    # Model: 1 operator_codes, 2 subgraphs, 4 buffers
    opcodes = []
    for t in fb.vec_tables(root, 1):
        dep = fb.scalar(t, 0, "<b", 0)
        new = fb.scalar(t, 3, "<i", 0)
        opcodes.append(max(dep, new))
    buffers = []
    for t in fb.vec_tables(root, 4):
        s, n = fb.vec(t, 0)
        buffers.append((s, n))
    sg = fb.vec_tables(root, 2)[0]
    tensors = []
    for i, t in enumerate(fb.vec_tables(sg, 0)):
        shape = fb.vec_i32(t, 0)
        dtype = fb.scalar(t, 1, "<b", 0)
        bidx = fb.scalar(t, 2, "<I", 0)
        name = fb.string(t, 3)
        data = None
        s, n = buffers[bidx] if bidx < len(buffers) else (0, 0)
        if n:
            np_dt = {TENSOR_F32: np.float32, TENSOR_F16: np.float16, TENSOR_I32: np.int32}.get(dtype)
            if np_dt is not None:
                data = np.frombuffer(buf, dtype=np_dt, count=n // np.dtype(np_dt).itemsize, offset=s)
                data = data.reshape(shape) if shape else data
        tensors.append(Tensor(i, name, shape, dtype, bidx, data))
    ops = []
    for t in fb.vec_tables(sg, 3):
        code = opcodes[fb.scalar(t, 0, "<I", 0)]
        ins = fb.vec_i32(t, 1)
        outs = fb.vec_i32(t, 2)
        opts: Dict[str, int] = {}
        p = fb.field(t, 4)
        if p:
            o = fb.indirect(p)
            if code == OP_CONV_2D:
                opts = dict(padding=fb.scalar(o, 0, "<b"), stride_w=fb.scalar(o, 1, "<i"),
                            stride_h=fb.scalar(o, 2, "<i"), act=fb.scalar(o, 3, "<b"),
                            dil_w=fb.scalar(o, 4, "<i", 1), dil_h=fb.scalar(o, 5, "<i", 1))
            elif code == OP_DEPTHWISE_CONV_2D:
                opts = dict(padding=fb.scalar(o, 0, "<b"), stride_w=fb.scalar(o, 1, "<i"),
                            stride_h=fb.scalar(o, 2, "<i"), depth_mult=fb.scalar(o, 3, "<i"),
                            act=fb.scalar(o, 4, "<b"), dil_w=fb.scalar(o, 5, "<i", 1),
                            dil_h=fb.scalar(o, 6, "<i", 1))
            elif code in (OP_MAX_POOL_2D, OP_AVERAGE_POOL_2D):
                opts = dict(padding=fb.scalar(o, 0, "<b"), stride_w=fb.scalar(o, 1, "<i"),
                            stride_h=fb.scalar(o, 2, "<i"), filter_w=fb.scalar(o, 3, "<i"),
                            filter_h=fb.scalar(o, 4, "<i"), act=fb.scalar(o, 5, "<b"))
            elif code == OP_RESIZE_BILINEAR:
                opts = dict(align_corners=fb.scalar(o, 2, "<b"), half_pixel=fb.scalar(o, 3, "<b"))
            elif code == OP_ADD:
                opts = dict(act=fb.scalar(o, 0, "<b"))
            elif code == OP_CONCATENATION:
                opts = dict(axis=fb.scalar(o, 0, "<i"), act=fb.scalar(o, 1, "<b"))
        ops.append(Op(code, ins, outs, opts))
    return Model(tensors, ops, fb.vec_i32(sg, 1), fb.vec_i32(sg, 2))


def summarize(m: Model) -> str:
    lines = []
    for op in m.ops:
        if op.code == OP_DEQUANTIZE:
            continue
        ins = ["%d%s" % (i, m.tensors[i].shape) for i in op.inputs if i >= 0]
        outs = ["%d%s" % (i, m.tensors[i].shape) for i in op.outputs]
        lines.append("%-18s %s -> %s %s" % (op.name, ins, outs, op.opts))
    return "\n".join(lines)


if __name__ == "__main__":
    import sys
    mdl = read_tflite(sys.argv[1])
    print("inputs", [(i, mdl.tensors[i].shape) for i in mdl.inputs])
    print("outputs", [(i, mdl.tensors[i].name, mdl.tensors[i].shape) for i in mdl.outputs])
    print(summarize(mdl))
