// Layer program of k_tail_ws (kernels_tail.cu): the small-resolution BlazeBlocks and head pairs of a detector,
// executed by ONE launch with the activations of an image resident in shared memory.  Built by plan.cpp, uploaded
// by engine.cu, interpreted by the kernel.
#pragma once

namespace fdt {

struct TailLayerD {
  int kind;            // 0: BlazeBlock (depthwise 3x3 -> pointwise -> + residual -> ReLU), 1: head pair (pointwise only -> graph outputs)
  int src, dst;        // activation buffer read / written: 0 = A (first resolution), 1 = B (after the stride-2 block)
  int stride, pad;     // depthwise stride, SAME pad-before
  int IH, IW, OH, OW;  // input / output spatial size
  int Cin, Cout, K16, Npad;
  int res;             // 0 none, 1 the same pixel of src, 2 2x2 max-pool of src (zero channel pad in both)
  int rec_off;         // float offset in the weight blob of the record [W fp16 Npad x K16 | dww 9 x K16 | dwb K16]
  int rec_bytes;       // its size (multiple of 16): one bulk copy per layer
  int bias_off;        // float offset of the pointwise bias [Npad]
  int c1, c2;          // heads: columns [0, c1) -> output o1, [c1, c1 + c2) -> output o2
  int o1, o2;          // heads: indices into TailP::outs (-1: none)
  int relu;            // 1: ReLU epilogue
  int pad_[10];
};
static_assert(sizeof(TailLayerD) == 128, "TailLayerD is copied as 16-byte words");

constexpr int kTailMaxLayers = 16;

}  // namespace fdt
