"""ORACLE (test infrastructure, not product code).

CPU executor of a parsed TFLite graph with TFLite op semantics, in fp64 (ground
truth that brackets every fp32 implementation, XNNPACK included) or fp32.
It restates what `Interpreter.invoke()` computes for the reference
(lib/src/models/face_detection_model.dart:391, lib/src/models/face_landmark.dart:302);
the arithmetic itself lives in flutter_litert 3.8.0 (TFLite + XNNPACK, un-vendored), so
the op definitions follow the published TFLite kernels:
  * CONV_2D / DEPTHWISE_CONV_2D: NHWC, weights OHWI / [1,KH,KW,C], SAME padding with
    pad_before = floor(pad_total / 2);
  * MAX_POOL_2D, PAD (constant 0), ADD, RELU, PRELU (alpha broadcast), RESHAPE,
    CONCATENATION, RESIZE_BILINEAR (align_corners / half_pixel_centers), DEQUANTIZE (f16->f32).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import tflite_reader as tr


def _same_pads(size, k, s, d=1):
    out = -(-size // s)
    eff = (k - 1) * d + 1
    total = max((out - 1) * s + eff - size, 0)
    return total // 2, total - total // 2


def _act(x, act):
    if act == 0:
        return x
    if act == 1:
        return torch.relu(x)
    if act == 3:
        return torch.clamp(x, 0, 6)
    raise NotImplementedError("fused activation %d" % act)


class GraphExecutor:
    def __init__(self, model: tr.Model, dtype=torch.float64):
        self.m = model
        self.dtype = dtype
        self.consts: Dict[int, torch.Tensor] = {}
        for t in model.tensors:
            if t.data is not None and t.dtype in (tr.TENSOR_F32, tr.TENSOR_F16):
                self.consts[t.index] = torch.from_numpy(np.array(t.data, dtype=np.float32)).to(dtype)

    def run(self, x_nhwc: np.ndarray, taps: Optional[Iterable[int]] = None) -> Dict[int, np.ndarray]:
        """Run the graph on `x_nhwc` [B,H,W,C]; returns {tensor_index: ndarray} for the
        graph outputs plus any `taps` (all activations if taps == 'all')."""
        m = self.m
        vals: Dict[int, torch.Tensor] = dict(self.consts)
        vals[m.inputs[0]] = torch.from_numpy(np.ascontiguousarray(x_nhwc)).to(self.dtype)
        B = x_nhwc.shape[0]
        for op in m.ops:
            c = op.code
            o = op.opts
            if c == tr.OP_DEQUANTIZE:
                vals[op.outputs[0]] = vals[op.inputs[0]]
                continue
            a = vals[op.inputs[0]]
            if c in (tr.OP_CONV_2D, tr.OP_DEPTHWISE_CONV_2D):
                w = vals[op.inputs[1]]
                b = vals[op.inputs[2]] if len(op.inputs) > 2 and op.inputs[2] >= 0 else None
                x = a.permute(0, 3, 1, 2)
                kh, kw = w.shape[1], w.shape[2]
                sh, sw = o["stride_h"], o["stride_w"]
                if o["padding"] == 0:
                    pt, pb = _same_pads(x.shape[2], kh, sh, o.get("dil_h", 1))
                    pl, pr = _same_pads(x.shape[3], kw, sw, o.get("dil_w", 1))
                    x = F.pad(x, (pl, pr, pt, pb))
                if c == tr.OP_CONV_2D:
                    y = F.conv2d(x, w.permute(0, 3, 1, 2), b, stride=(sh, sw),
                                 dilation=(o.get("dil_h", 1), o.get("dil_w", 1)))
                else:
                    C = w.shape[3]
                    y = F.conv2d(x, w.permute(3, 0, 1, 2), b, stride=(sh, sw), groups=C,
                                 dilation=(o.get("dil_h", 1), o.get("dil_w", 1)))
                r = _act(y.permute(0, 2, 3, 1), o["act"])
            elif c == tr.OP_MAX_POOL_2D:
                x = a.permute(0, 3, 1, 2)
                if o["padding"] == 0:
                    pt, pb = _same_pads(x.shape[2], o["filter_h"], o["stride_h"])
                    pl, pr = _same_pads(x.shape[3], o["filter_w"], o["stride_w"])
                    x = F.pad(x, (pl, pr, pt, pb), value=float("-inf"))
                y = F.max_pool2d(x, (o["filter_h"], o["filter_w"]), (o["stride_h"], o["stride_w"]))
                r = _act(y.permute(0, 2, 3, 1), o["act"])
            elif c == tr.OP_RELU:
                r = torch.relu(a)
            elif c == tr.OP_PRELU:
                alpha = vals[op.inputs[1]]
                r = torch.where(a >= 0, a, a * alpha)
            elif c == tr.OP_ADD:
                r = _act(a + vals[op.inputs[1]], o.get("act", 0))
            elif c == tr.OP_PAD:
                pads = np.array(m.tensors[op.inputs[1]].data).reshape(-1, 2)
                flat = []
                for d in range(pads.shape[0] - 1, -1, -1):
                    flat += [int(pads[d, 0]), int(pads[d, 1])]
                r = F.pad(a, flat)
            elif c == tr.OP_RESHAPE:
                shape = list(m.tensors[op.outputs[0]].shape)
                shape[0] = B
                r = a.reshape(shape)
            elif c == tr.OP_CONCATENATION:
                r = _act(torch.cat([vals[i] for i in op.inputs], dim=o["axis"]), o.get("act", 0))
            elif c == tr.OP_RESIZE_BILINEAR:
                oh, ow = m.tensors[op.outputs[0]].shape[1:3]
                r = _resize_bilinear(a, oh, ow, bool(o["align_corners"]), bool(o["half_pixel"]))
            else:
                raise NotImplementedError(op.name)
            vals[op.outputs[0]] = r
        want = list(m.outputs)
        if taps == "all":
            want = [t for t in vals if t not in self.consts]
        elif taps is not None:
            want += list(taps)
        return {i: vals[i].detach().numpy() for i in want}


def _resize_bilinear(a, oh, ow, align_corners, half_pixel):
    """TFLite reference RESIZE_BILINEAR (tensorflow/lite/kernels/internal/reference/resize_bilinear.h)."""
    B, ih, iw, C = a.shape

    def axis(out, inp):
        if align_corners and out > 1:
            scale = (inp - 1) / (out - 1)
        else:
            scale = inp / out
        idx = torch.arange(out, dtype=torch.float64)
        src = (idx + 0.5) * scale - 0.5 if half_pixel else idx * scale
        f = torch.floor(src)
        i0 = torch.clamp(f, 0, inp - 1).long()
        i1 = torch.clamp(f + 1, 0, inp - 1).long()
        return i0, i1, (src - f)

    y0, y1, fy = axis(oh, ih)
    x0, x1, fx = axis(ow, iw)
    fy = fy.to(a.dtype).view(1, oh, 1, 1)
    fx = fx.to(a.dtype).view(1, 1, ow, 1)
    r0 = a[:, y0]
    r1 = a[:, y1]
    top = r0[:, :, x0] * (1 - fx) + r0[:, :, x1] * fx
    bot = r1[:, :, x0] * (1 - fx) + r1[:, :, x1] * fx
    return top * (1 - fy) + bot * fy
