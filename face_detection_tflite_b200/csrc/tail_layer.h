// Layer program of k_tail_ws (kernels_tail.cu): a chain of small-resolution layers executed by ONE launch with the
// activations of an image resident in shared memory - the 16x16 / 8x8 BlazeBlocks and head pairs of a detector, or the
// 12x12 ... 3x3 trunk and both branches of the face-landmark net.  Built by plan.cpp, uploaded by engine.cu,
// interpreted by the kernel.
#pragma once

namespace fdt {

struct TailLayerD {
  int kind;            // 0: block (depthwise 3x3 -> pointwise -> + residual -> activation), 1: head pair (pointwise only -> graph outputs),
                       // 2: pointwise only -> activation -> buffer, 3: dot product of the whole map with one filter -> one scalar per image
  int src, dst;        // shared-memory activation buffer read / written (dst -1: none)
  int stride, pad;     // depthwise stride, SAME pad-before
  int IH, IW, OH, OW;  // input / output spatial size
  int Cin, Cout, K16, Npad;
  int res;             // 0 none, 1 the same pixel of buffer rbuf, 2 2x2 max-pool of buffer rbuf (zero channel pad in both);
                       // k_chain_wide only: 3 the same pixel / 4 2x2 max-pool of the HBM tensor TailP::rsrc[rbuf]
  int rec_off;         // float offset in the weight blob of the pointwise record [W fp16 w_parts x Npad x K16]
  int rec_bytes;       // its size (multiple of 16): one bulk copy per layer
  int bias_off;        // float offset of the pointwise bias [Npad]
  int c1, c2;          // heads: columns [0, c1) -> output o1, [c1, c1 + c2) -> output o2
  int o1, o2;          // indices into TailP::outs (-1: none); kinds 0 / 2: the layer's result is ALSO written to HBM tensor o1
  int act;             // 0 none, 1 ReLU, 2 PReLU (slopes at alpha_off)
  int rbuf;            // residual buffer
  int w_parts;         // 1: weights exact in fp16; 2: W = hi + lo (fp32-origin weights, a third MMA pass A_hi x W_lo)
  int tap_off;         // float offset of the depthwise record [taps 9 x K16 | bias K16] (kind 3: the filter [OH*OW*Cin] | bias)
  int tap_bytes;
  int alpha_off;       // float offset of the PReLU slopes [Npad]
  float wscale;        // the stored weights are W * 2^s; the epilogue multiplies the accumulator by wscale = 2^-s
  // k_chain_wide (layers wider than 128 channels): the pointwise product runs as ng column groups x nk K chunks (<= 128 each);
  // W arrives as blocks [<= 128 columns][<= 128 K] in consumption order (group, K chunk, block of the group)
  int blk0;            // index of the layer's first block in the chain's block table
  int nk;              // K chunks
  int ng;              // column groups | blocks per group << 8  (two-tile maps: groups of one block; one-tile maps: one group of <= 3 blocks)
  int bias_s;          // float offset of the layer's bias / slopes inside the kernel's shared-memory copies
};
static_assert(sizeof(TailLayerD) == 128, "TailLayerD is copied as 16-byte words");

// one W block of k_chain_wide: [w_parts][ncols][kw] fp16 in UMMA K-major core matrices, kw = min(128, K16 - 128 * chunk)
struct TailBlk {
  int off;             // float offset in the weight blob
  int bytes;           // multiple of 16: one bulk copy
  int n0, ncols;       // accumulator columns [n0, n0 + ncols) of the layer
};
static_assert(sizeof(TailBlk) == 16, "TailBlk is copied as 16-byte words");

constexpr int kTailMaxLayers = 16;
constexpr int kTailMaxBlks = 128;
constexpr int kTailMaxBufs = 8;

}  // namespace fdt
