// Execution plan: the TFLite graph lowered to a list of kernel launches over a chunk of images.
// fuse_level 0: one step per TFLite op (parity taps for every tensor);
// fuse_level 1: BlazeBlock fusion (DW3x3 + PW1x1 + residual + activation in one kernel), stems as
//               im2col GEMM reading the u8 letterboxed image, liveness-based activation reuse;
// fuse_level 2: as 1 but without buffer reuse (every materialised tensor stays readable).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "tail_layer.h"
#include "tflite_model.h"

namespace fdt {

struct PTensor {
  int tf = -1;
  int H = 1, W = 1, C = 1, Cs = 1;
  long long istride = 0;     // floats per image
  int root = -1;             // >= 0: dense view into graph output #root
  long long view_off = 0;    // float offset inside that output's per-image row
  long long arena_off = -1;  // per-image float offset inside the activation arena
  int def_step = -1, last_use = -1;
  bool materialized = false;
  bool is_input = false;
};

enum StepKind { kStepNormalize, kStepNaiveConv, kStepGemmConv, kStepDwPw, kStepAdd, kStepAct, kStepPadC,
                kStepMaxPool, kStepResize, kStepStem, kStepBlockWs, kStepStemWs, kStepTailWs, kStepFcTc, kStepBlockTs };

struct PStep {
  StepKind kind = kStepAct;
  std::string name;
  int in = -1, in2 = -1, out = -1;
  int kh = 1, kw = 1, sh = 1, sw = 1, pt = 0, pl = 0, act = 0, depthwise = 0;
  bool in_u8 = false;
  bool has_dw = false;
  int dws = 1, dpt = 0, dpl = 0, res_pool = 0, res_mode = 0;
  long long w = -1, bias = -1, alpha = -1, dww = -1, dwb = -1;  // float offsets into Plan::blob
  int K = 0, KP = 0, KS = 0, Cout = 0, CoutP = 0, NC = 0, nchunks = 0, NPG = 0, TM = 0;
  int TH = 0, TW = 0, G = 1, IH = 0, IW = 0, tilesX = 1, tilesY = 1;
  size_t smem = 0;
  int K8 = 0, Npad = 0, a_rows = 0, RS = 1, tmem_cols = 0, w_parts = 1, nbuf = 1;   // tensor-core variant
  double out_scale = 1.0;    // k_stem_ws: D * out_scale + bias
  int ns = 0, na = 0, nd = 8, nt = 2, in_stage_floats = 0;
  int no = 0, KSo = 0, out_stage_floats = 0;           // TMA-store epilogue (k_block_ws)
  int nr = 0, KSr = 0, res_stage_floats = 0;           // k_block_ws: ring of TMA-staged residual tiles (residual in another HBM tensor)
  int deint = 0, plane_floats = 0, PW = 0;             // k_block_ws stride 2: even / odd input columns staged as two planes of PW columns
  int out2 = -1, c1 = 0, c2 = 0;                        // two heads in one launch: columns [0,c1) -> out, [c1,c1+c2) -> out2   // warp-specialised variant: input stages, A buffers, depthwise warps
  int fh = 1, fw = 1, align = 0, half = 0;
  std::vector<int> extra_out;   // further tensors the step materialises (k_tail_ws: every graph output of the fused tail)
  // k_block_ts (kernels_ts.cu): fp16 weight record, staged pixel stride, input ring
  long long ts_rec = -1, ts_bias = -1;
  int ts_rec_bytes = 0, ts_k16 = 0, ts_npad = 0, ts_ks = 0, ts_ns = 0, ts_stage_bytes = 0, ts_res = 0, ts_relu = 0;
  // k_tail_ws: the layer program and its shared-memory geometry (see kernels_tail.cu)
  std::vector<TailLayerD> tail;
  std::vector<int> tail_outs;   // PTensor ids the heads write, indexed by TailLayerD::o1 / o2
  int tail_nbuf = 0, tail_buf_off[kTailMaxBufs] = {0}, tail_buf_ks[kTailMaxBufs] = {0}, tail_buf_px[kTailMaxBufs] = {0};   // shared-memory activation buffers
  // k_chain_wide (layers wider than 128 channels): W block table, HBM residual tensors, shared-memory bias area
  std::vector<TailBlk> tail_blks;
  std::vector<int> tail_rsrc;
  int tail_wide = 0, tail_bias_floats = 0;
  int tail_act_floats = 0, tail_in_bytes = 0, tail_last_a = 0, tail_wbuf = 0, tail_wdepth = 2, tail_tbuf = 0;
  double macs = 0;  // per image
};

struct Plan {
  std::vector<PTensor> tensors;
  std::vector<PStep> steps;
  std::vector<float> blob;
  std::map<int, int> tf2pt;
  int input = -1, in_h = 0, in_w = 0;
  std::vector<int> outputs;            // PTensor id per graph output
  std::vector<long long> out_elems;    // floats per image per graph output
  long long arena_per_image = 0;       // floats
  int fuse_level = 1;

  bool build(const TfModel& m, int fuse_level, std::string* err, bool use_tc = true);
  std::string describe() const;
};

}  // namespace fdt
