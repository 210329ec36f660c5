#include "plan.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>

namespace fdt {
namespace {

constexpr size_t kSmemLimit = 200 * 1024;
constexpr int kActNone = 0, kActRelu = 1, kActPrelu = 2;

inline int ru(int v, int m) { return (v + m - 1) / m * m; }
inline int smem_stride(int KP) { return KP + (((KP / 4) % 2 == 0) ? 4 : 0); }

void same_pad(int in, int k, int s, int* before) {
  int out = (in + s - 1) / s;
  int total = (out - 1) * s + k - in;
  if (total < 0) total = 0;
  *before = total / 2;
}


// fp32 -> fp16 bit pattern, round to nearest even, subnormals handled; |v| must be below the fp16 overflow threshold
uint16_t f32_to_f16(float v) {
  uint32_t x;
  std::memcpy(&x, &v, 4);
  const uint32_t sign = (x >> 16) & 0x8000u;
  const int32_t e = (int32_t)((x >> 23) & 0xFF) - 127 + 15;
  uint32_t mant = x & 0x7FFFFFu;
  if (((x >> 23) & 0xFF) == 0) return (uint16_t)sign;          // fp32 zero / subnormal -> 0
  if (e >= 31) return (uint16_t)(sign | 0x7BFFu);              // clamp (callers scale weights to avoid this)
  if (e <= 0) {
    if (e < -10) return (uint16_t)sign;
    mant |= 0x800000u;
    const int shift = 14 - e;                                   // 14..24
    uint32_t h = mant >> shift;
    const uint32_t rem = mant & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (rem > half || (rem == half && (h & 1))) ++h;
    return (uint16_t)(sign | h);
  }
  uint32_t h = ((uint32_t)e << 10) | (mant >> 13);
  const uint32_t rem = mant & 0x1FFFu;
  if (rem > 0x1000u || (rem == 0x1000u && (h & 1))) ++h;       // may carry into the exponent: still correct
  return (uint16_t)(sign | h);
}
float f16_to_f32(uint16_t h) {
  const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  int e = (h >> 10) & 0x1F;
  uint32_t mant = h & 0x3FFu;
  float out;
  uint32_t x;
  if (e == 0) {
    if (mant == 0) { x = sign; std::memcpy(&out, &x, 4); return out; }
    float f = (float)mant * 5.9604644775390625e-8f;            // 2^-24
    return (h & 0x8000u) ? -f : f;
  }
  x = sign | ((uint32_t)(e - 15 + 127) << 23) | (mant << 13);
  std::memcpy(&out, &x, 4);
  return out;
}

struct View { int root = -1; long long off = 0; };

struct Fused {       // analysis result for one CONV_2D
  bool ok = false;
  PStep st;
  int src_tf = -1, res_tf = -1, out_tf = -1, out2_tf = -1;
};

struct Builder {
  const TfModel& m;
  Plan& P;
  int fuse;
  std::string err;
  std::vector<char> done;
  std::vector<int> ncons;
  std::vector<View> views;
  std::map<int, Fused> fused;  // conv op index -> fused step

  bool use_tc = true;
  Builder(const TfModel& mm, Plan& pp, int f) : m(mm), P(pp), fuse(f) {}

  bool fail(const std::string& e) { err = e; return false; }

  long long numel(int t) const { return m.tensors[t].numel(); }

  bool is_view(int t) const { return views[t].root >= 0; }

  // ---- output views ---------------------------------------------------------------------------
  bool assign_view(int t, int root, long long off) {
    views[t].root = root;
    views[t].off = off;
    int p = m.producer(t);
    if (p < 0) return true;
    const TfOp& op = m.ops[p];
    if (op.code == kOpReshape) {
      done[p] = 1;
      return assign_view(op.in[0], root, off);
    }
    if (op.code == kOpConcat) {
      const TfTensor& ot = m.tensors[t];
      int axis = op.axis < 0 ? op.axis + (int)ot.shape.size() : op.axis;
      for (int d = 0; d < axis; ++d)
        if (ot.shape[d] != 1) return fail("CONCATENATION over a non-leading axis is not supported");
      done[p] = 1;
      long long cur = off;
      for (int i : op.in) {
        if (!assign_view(i, root, cur)) return false;
        cur += numel(i);
      }
    }
    return true;
  }

  // Channel stride of an arena tensor: channels rounded up to a quad (float4 loads / stores everywhere), pad channels exact zeros.
  // Measured alternative (round 2): rounding up to 8 channels, so that every pixel record is 32-byte aligned and every epilogue uses
  // 256-bit stores, LOSES on the only layers it changes (short-range 28 / 36 / 42 channels): block 2 273 -> 295 us per 1024 frames
  // (14 % more bytes written), block 3 126 -> 131, blocks 4-6 -8 us in total.
  int cs(int C) const { return ru(C, 4); }

  int pt(int tf) {
    auto it = P.tf2pt.find(tf);
    if (it != P.tf2pt.end()) return it->second;
    const TfTensor& t = m.tensors[tf];
    PTensor x;
    x.tf = tf;
    int r = (int)t.shape.size();
    if (r == 4) { x.H = t.shape[1]; x.W = t.shape[2]; x.C = t.shape[3]; }
    else if (r == 3) { x.H = 1; x.W = t.shape[1]; x.C = t.shape[2]; }
    else if (r == 2) { x.H = 1; x.W = 1; x.C = t.shape[1]; }
    else { x.H = 1; x.W = 1; x.C = (int)t.numel(); }
    if (is_view(tf)) {
      x.Cs = x.C;
      x.root = views[tf].root;
      x.view_off = views[tf].off;
      x.istride = P.out_elems[x.root];
    } else {
      x.Cs = cs(x.C);
      x.istride = (long long)x.H * x.W * x.Cs;
    }
    P.tensors.push_back(x);
    P.tf2pt[tf] = (int)P.tensors.size() - 1;
    return (int)P.tensors.size() - 1;
  }

  long long push(const std::vector<float>& v, size_t padded) {
    while (P.blob.size() % 4) P.blob.push_back(0.f);  // 16-byte alignment of every array
    long long off = (long long)P.blob.size();
    P.blob.insert(P.blob.end(), v.begin(), v.end());
    if (padded > v.size()) P.blob.insert(P.blob.end(), padded - v.size(), 0.f);
    return off;
  }

  // single consumer op of tensor t (or -1)
  int sole_consumer(int t) const {
    if (ncons[t] != 1) return -1;
    auto c = m.consumers(t);
    return c.size() == 1 ? c[0] : -1;
  }

  // Follows T -> [RELU | PRELU]; returns the activation to fuse and the final tensor.
  void fuse_activation(int* cur, int* act, int* alpha_tf, std::vector<int>* absorbed) {
    if (is_view(*cur)) return;
    int c = sole_consumer(*cur);
    if (c < 0 || done[c]) return;
    const TfOp& op = m.ops[c];
    if (op.code == kOpRelu) {
      *act = kActRelu;
    } else if (op.code == kOpPrelu) {
      std::vector<float> a;
      if (!m.const_f32(op.in[1], &a) || (int)a.size() != m.tensors[*cur].shape.back()) return;
      *act = kActPrelu;
      *alpha_tf = op.in[1];
    } else {
      return;
    }
    absorbed->push_back(c);
    *cur = op.out[0];
  }

  bool pack_alpha(int alpha_tf, size_t padded, long long* off) {
    std::vector<float> a;
    if (!m.const_f32(alpha_tf, &a)) return fail("PRELU alpha is not constant");
    *off = push(a, padded);
    return true;
  }

  // ---- tiling of the GEMM part ----------------------------------------------------------------
  // Output-channel chunking and thread tile: quads of 4 channels, <= 24 quads per chunk; 4 pixels per
  // thread while that keeps the CTA <= 384 threads, else 8.
  static void plan_columns(PStep* s) {
    s->CoutP = ru(s->Cout, 4);
    int quads = s->CoutP / 4;
    s->nchunks = (quads + 23) / 24;
    int nq = (quads + s->nchunks - 1) / s->nchunks;
    s->NC = nq * 4;
    s->TM = nq <= 12 ? 4 : 8;
    s->NPG = 128 / s->TM;
  }

  // ---- graph pattern of one pointwise conv: [DW3x3 ->] PW [-> ADD(residual | PAD / MAXPOOL2x2 of it)] [-> RELU | PRELU] --------
  struct PwMatch {
    bool has_dw = false;
    int dws = 1, dpt = 0, dpl = 0;
    int src = -1, cur = -1, res = -1, pool = 0, act = kActNone, alpha_tf = -1;
    std::vector<int> absorbed;     // ops folded into the step besides the CONV_2D itself (the depthwise first when has_dw)
  };
  // graph_only: the residual source need not be produced by an earlier STEP (the caller checks where it lives)
  bool match_pointwise(int ci, PwMatch* M, bool graph_only) {
    const TfOp& conv = m.ops[ci];
    if (conv.code != kOpConv2D || conv.in.size() < 2) return false;
    const TfTensor& wt = m.tensors[conv.in[1]];
    if (wt.shape.size() != 4 || wt.shape[1] != 1 || wt.shape[2] != 1) return false;
    if (conv.stride_h != 1 || conv.stride_w != 1 || (conv.act != 0 && conv.act != 1)) return false;
    int X = conv.in[0];
    M->src = X;
    int dwi = m.producer(X);
    if (dwi >= 0 && !done[dwi] && m.ops[dwi].code == kOpDwConv2D && ncons[X] == 1 && !is_view(X)) {
      const TfOp& dw = m.ops[dwi];
      const TfTensor& dwt = m.tensors[dw.in[1]];
      bool ok = dwt.shape.size() == 4 && dwt.shape[1] == 3 && dwt.shape[2] == 3 && dw.depth_mult == 1 &&
                dw.dil_h == 1 && dw.dil_w == 1 && dw.padding == 0 && dw.stride_h == dw.stride_w &&
                (dw.stride_h == 1 || dw.stride_h == 2) && dw.act == 0 && dw.in.size() >= 3;
      if (ok) {
        M->has_dw = true;
        M->dws = dw.stride_h;
        M->src = dw.in[0];
        const TfTensor& it = m.tensors[M->src];
        same_pad(it.dim(1), 3, M->dws, &M->dpt);
        same_pad(it.dim(2), 3, M->dws, &M->dpl);
        M->absorbed.push_back(dwi);
      }
    }
    const int Cout = wt.shape[0];
    int cur = conv.out[0];
    M->act = conv.act == 1 ? kActRelu : kActNone;
    if (M->act == kActNone && !is_view(cur)) {
      int c = sole_consumer(cur);
      if (c >= 0 && !done[c] && m.ops[c].code == kOpAdd && (m.ops[c].act == 0 || m.ops[c].act == 1) &&
          m.ops[c].in.size() == 2) {
        const TfOp& add = m.ops[c];
        int R = add.in[0] == cur ? add.in[1] : add.in[0];
        std::vector<int> abs2;
        int r = R;
        int pool = 0;
        bool good = R != cur;
        int pp = m.producer(r);
        if (good && pp >= 0 && m.ops[pp].code == kOpPad && ncons[r] == 1 && !is_view(r) && !done[pp]) {
          std::vector<int> pads;
          const TfTensor& pi = m.tensors[m.ops[pp].in[0]];
          if (m.const_i32(m.ops[pp].in[1], &pads) && pads.size() == 8 && pads[0] == 0 && pads[1] == 0 &&
              pads[2] == 0 && pads[3] == 0 && pads[4] == 0 && pads[5] == 0 && pads[6] == 0 && pi.shape.size() == 4) {
            abs2.push_back(pp);
            r = m.ops[pp].in[0];
            pp = m.producer(r);
          }
        }
        if (good && pp >= 0 && m.ops[pp].code == kOpMaxPool && ncons[r] == 1 && !is_view(r) && !done[pp]) {
          const TfOp& mp = m.ops[pp];
          if (mp.filter_h == 2 && mp.filter_w == 2 && mp.stride_h == 2 && mp.stride_w == 2 && mp.act == 0) {
            abs2.push_back(pp);
            r = mp.in[0];
            pool = 1;
          }
        }
        const TfTensor& rt = m.tensors[r];
        const TfTensor& ot = m.tensors[cur];
        int rp = m.producer(r);
        // the residual source must be materialised by an earlier step: either a stand-alone op or
        // the final tensor of an already fused chain (never an absorbed intermediate)
        bool produced_before = graph_only || (rp >= 0 && rp < ci && (!done[rp] || is_output_of_fused(r)));
        good = good && rt.shape.size() == 4 && rt.dim(3) <= Cout && produced_before;
        if (good) {
          if (pool) good = (rt.dim(1) + 1) / 2 == ot.dim(1) && (rt.dim(2) + 1) / 2 == ot.dim(2);
          else good = rt.dim(1) == ot.dim(1) && rt.dim(2) == ot.dim(2);
        }
        if (good) {
          M->res = r;
          M->pool = pool;
          M->absorbed.push_back(c);
          for (int a : abs2) M->absorbed.push_back(a);
          cur = add.out[0];
          if (add.act == 1) M->act = kActRelu;
        }
      }
      if (M->act == kActNone) fuse_activation(&cur, &M->act, &M->alpha_tf, &M->absorbed);
    }
    M->cur = cur;
    return true;
  }

  // ---- analysis of one pointwise conv: DW3x3 -> PW -> (+res) -> act ---------------------------
  bool analyse_pointwise(int ci, Fused* F) {
    const TfOp& conv = m.ops[ci];
    PwMatch M;
    if (!match_pointwise(ci, &M, false)) return false;
    const TfTensor& wt = m.tensors[conv.in[1]];
    PStep st;
    st.kind = kStepDwPw;
    std::vector<int> absorbed = M.absorbed;
    const int src = M.src;
    st.has_dw = M.has_dw; st.dws = M.dws; st.dpt = M.dpt; st.dpl = M.dpl;
    const TfTensor& st_in = m.tensors[src];
    if (st_in.shape.size() != 4) return false;
    if (is_view(src) && (st_in.dim(3) % 4 != 0 || views[src].off % 4 != 0 || P.out_elems[views[src].root] % 4 != 0))
      return false;
    int Cin = st_in.dim(3);
    st.Cout = wt.shape[0];
    st.K = Cin;
    st.KP = ru(Cin, 4);
    st.KS = smem_stride(st.KP);
    plan_columns(&st);
    // epilogue chain
    int cur = M.cur;
    int res = M.res, alpha_tf = M.alpha_tf;
    st.act = M.act;
    if (res >= 0) {
      const TfTensor& rt = m.tensors[res];
      st.res_pool = M.pool;
      // the residual is the block input itself: take it from the staged tile when it is there
      bool even = rt.dim(1) % 2 == 0 && rt.dim(2) % 2 == 0;
      st.res_mode = (st.has_dw && res == src && (!M.pool || (even && st.dws == 2))) ? 1 : 2;
    }
    // spatial tiling
    const TfTensor& ot = m.tensors[conv.out[0]];
    int OH = ot.dim(1), OW = ot.dim(2);
    if (!plan_spatial(&st, OH, OW)) return false;
    // commit
    for (int a : absorbed) done[a] = 1;
    done[ci] = 1;
    F->ok = true;
    F->st = st;
    F->src_tf = src;
    F->res_tf = res;
    F->out_tf = cur;
    F->st.name = m.tensors[cur].name;
    fused_outputs.push_back(cur);
    // weights
    std::vector<float> w, b;
    if (!m.const_f32(conv.in[1], &w)) return fail("conv weights are not constant");
    if (conv.in.size() > 2 && conv.in[2] >= 0) { if (!m.const_f32(conv.in[2], &b)) return fail("conv bias is not constant"); }
    b.resize(st.Cout, 0.f);
    // Two heads of one scale (classificator + regressor read the same tensor and write dense graph outputs): one launch.
    // The head whose channel count is a multiple of 4 comes first so that its columns keep their float4 stores; the kernel
    // writes columns [0, c1) to its output and [c1, c1 + c2) to the other head's.
    if (use_tc && !st.has_dw && res < 0 && st.act == kActNone && is_view(cur)) {
      int sib = -1;
      for (size_t j = 0; j < m.ops.size() && sib < 0; ++j) {
        const TfOp& o2 = m.ops[j];
        if ((int)j == ci || done[j] || o2.code != kOpConv2D || o2.in[0] != conv.in[0]) continue;
        const TfTensor& w2t = m.tensors[o2.in[1]];
        if (w2t.shape.size() != 4 || w2t.shape[1] != 1 || w2t.shape[2] != 1 || w2t.shape[3] != Cin) continue;
        if (o2.stride_h != 1 || o2.stride_w != 1 || o2.act != 0 || !is_view(o2.out[0])) continue;
        sib = (int)j;
      }
      if (sib >= 0) {
        const TfOp& o2 = m.ops[sib];
        const int sib_out = o2.out[0];
        std::vector<float> w2, b2;
        bool ok2 = m.const_f32(o2.in[1], &w2);
        if (ok2 && o2.in.size() > 2 && o2.in[2] >= 0) ok2 = m.const_f32(o2.in[2], &b2);
        const int C2n = m.tensors[o2.in[1]].shape[0];
        b2.resize(C2n, 0.f);
        const bool this_primary = st.Cout % 4 == 0 && (C2n % 4 != 0 || st.Cout >= C2n);
        const int Cp = this_primary ? st.Cout : C2n, Cq = this_primary ? C2n : st.Cout;
        const int p_out = this_primary ? cur : sib_out, q_out = this_primary ? sib_out : cur;
        const bool vec_ok = views[p_out].off % 4 == 0 && P.out_elems[views[p_out].root] % 4 == 0;
        if (ok2 && Cp % 4 == 0 && vec_ok) {
          std::vector<float> wm, bm;
          const std::vector<float>& wp = this_primary ? w : w2;
          const std::vector<float>& wq = this_primary ? w2 : w;
          const std::vector<float>& bp = this_primary ? b : b2;
          const std::vector<float>& bq = this_primary ? b2 : b;
          wm.insert(wm.end(), wp.begin(), wp.end()); wm.insert(wm.end(), wq.begin(), wq.end());
          bm.insert(bm.end(), bp.begin(), bp.end()); bm.insert(bm.end(), bq.begin(), bq.end());
          PStep t = st;
          t.Cout = Cp + Cq; t.c1 = Cp; t.c2 = Cq;
          if (plan_tc(&t, Cin, OH, OW, tf32_exact(wm) ? 1 : 2) && plan_ws(&t, OH, OW)) {
            st.Cout = Cp + Cq; st.c1 = Cp; st.c2 = Cq;
            plan_columns(&st);
            const std::string nm = m.tensors[p_out].name + "+" + m.tensors[q_out].name;
            F->st = st;
            F->st.name = nm;
            F->out_tf = p_out; F->out2_tf = q_out;
            done[sib] = 1;
            fused_outputs.push_back(sib_out);
            w.swap(wm); b.swap(bm);
          }
        }
      }
    }
    const int w_parts = tf32_exact(w) ? 1 : 2;
    bool tc = false;
    if (use_tc) {
      PStep t = F->st;                       // shapes the warp-specialised kernel cannot take stay on the CUDA-core kernel
      res_c_hint = res >= 0 ? m.tensors[res].dim(3) : 0;
      tc = plan_tc(&t, Cin, OH, OW, w_parts) && plan_ws(&t, OH, OW);
      res_c_hint = 0;
      if (tc) F->st = t;
    }
    if (F->st.c2 > 0 && F->st.kind != kStepBlockWs) return fail("internal: merged heads need the warp-specialised kernel");
    const PStep& fs = F->st;
    if (tc) {
      // B operand [Npad x K8] in the UMMA K-major core-matrix layout (8 rows x 16 bytes per core matrix)
      const size_t SBO = (size_t)(fs.K8 / 4) * 128, LBO = 128;
      std::vector<float> wb((size_t)w_parts * fs.Npad * fs.K8, 0.f);
      for (int n = 0; n < fs.Cout; ++n)
        for (int k = 0; k < Cin; ++k) {
          size_t o = ((size_t)(n >> 3) * SBO + (size_t)(k >> 2) * LBO + (size_t)(n & 7) * 16 + (size_t)(k & 3) * 4) / 4;
          split_w(w[(size_t)n * Cin + k], w_parts, &wb[o], &wb[o + (w_parts > 1 ? (size_t)fs.Npad * fs.K8 : 0)]);
        }
      F->st.w = push(wb, wb.size());
      F->st.bias = push(b, (size_t)fs.Npad + 8);
      if (alpha_tf >= 0 && !pack_alpha(alpha_tf, (size_t)fs.Npad + 8, &F->st.alpha)) return false;
    } else {
      std::vector<float> wk((size_t)st.KP * st.CoutP, 0.f);
      for (int co = 0; co < st.Cout; ++co)
        for (int c = 0; c < Cin; ++c) wk[(size_t)c * st.CoutP + co] = w[(size_t)co * Cin + c];
      size_t padded = (size_t)st.nchunks * st.NC + 8;
      F->st.w = push(wk, wk.size());
      F->st.bias = push(b, padded);
      if (alpha_tf >= 0 && !pack_alpha(alpha_tf, padded, &F->st.alpha)) return false;
    }
    if (st.has_dw) {
      const TfOp& dw = m.ops[absorbed[0]];
      std::vector<float> dwv, dbv;
      if (!m.const_f32(dw.in[1], &dwv) || !m.const_f32(dw.in[2], &dbv)) return fail("depthwise weights are not constant");
      const int kp = tc ? fs.K8 : st.KP;
      std::vector<float> dk((size_t)9 * kp, 0.f);
      for (int t = 0; t < 9; ++t)
        for (int c = 0; c < Cin; ++c) dk[(size_t)t * kp + c] = dwv[(size_t)t * Cin + c];
      F->st.dww = push(dk, dk.size());
      F->st.dwb = push(dbv, kp);
    }
    F->st.macs = (double)OH * OW * ((double)Cin * st.Cout + (st.has_dw ? 9.0 * Cin : 0.0));
    if (tc && F->st.kind == kStepBlockWs) plan_ts(&F->st, st_in, OH, OW, alpha_tf >= 0);
    return true;
  }

  // Re-plans a kStepBlockWs BlazeBlock for k_block_ts (kernels_ts.cu) when the layer qualifies: fp16-exact weights, at most
  // 64 input / output channels (two tile slots of operand + accumulator fit the 512 TMEM columns), output tiled by 16 x 16
  // (stride 1) or 8 x 16 (stride 2), residual = the block's own input.  The k_block_ws record stays in the blob (the tail
  // fusion and FDT_TS=0 use it).
  void plan_ts(PStep* st, const TfTensor& in, int OH, int OW, bool prelu) {
    static const int want = [] { const char* e = std::getenv("FDT_TS"); return e ? std::atoi(e) : 1; }();      // FDT_TS=0: A/B against k_block_ws
    if (!want || !st->has_dw || st->c2 > 0 || st->w_parts != 1 || prelu) return;
    // (Round 1 kept blocks of up to 24 input channels on k_block_ws, whose threads hold their depthwise taps in registers for the whole
    // kernel.  With the taps read from the kernel parameters - constant bank, no shared-memory traffic - and two column groups per
    // TMEM round trip this kernel wins there too: 64x64x24 -> 24: 215 vs 257 us per 1024 frames, 24 -> 28: 249 vs 277.)
    if (st->act != kActRelu && st->act != kActNone) return;
    const int Cin = in.dim(3), K16 = ru(Cin, 16), Npad = ru(st->Cout, 16);
    if (K16 > 64 || Npad > 64) return;
    if (st->dws == 1 ? (OH % 16 || OW % 16 || st->dpt != 1 || st->dpl != 1)
                     : (OH % 8 || OW % 16 || st->dpt != 0 || st->dpl != 0 || in.dim(1) != 2 * OH || in.dim(2) != 2 * OW)) return;
    int res = 0;
    if (st->res_mode != 0) {
      if (st->res_mode != 1) return;
      res = st->res_pool ? 2 : 1;
      if ((res == 2) != (st->dws == 2)) return;
    }
    // weights back from the k_block_ws operand layout (exact values) -> fp16 UMMA K-major core matrices
    const size_t SBO8 = (size_t)(st->K8 / 4) * 128, SBO16 = (size_t)(K16 / 8) * 128;
    std::vector<uint16_t> wh((size_t)Npad * K16, 0);
    for (int n = 0; n < st->Cout; ++n)
      for (int k = 0; k < Cin; ++k) {
        const float v = P.blob[(size_t)st->w + ((size_t)(n >> 3) * SBO8 + (size_t)(k >> 2) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 3) * 4) / 4];
        const uint16_t hbits = f32_to_f16(v);
        if (f16_to_f32(hbits) != v) return;
        wh[((size_t)(n >> 3) * SBO16 + (size_t)(k >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2) / 2] = hbits;
      }
    std::vector<float> rec(wh.size() / 2 + (size_t)10 * K16, 0.f);
    std::memcpy(rec.data(), wh.data(), wh.size() * 2);
    float* dw = rec.data() + wh.size() / 2;
    for (int k = 0; k < 9; ++k)
      for (int c = 0; c < Cin; ++c) dw[(size_t)k * K16 + c] = P.blob[(size_t)st->dww + (size_t)k * st->K8 + c];
    for (int c = 0; c < Cin; ++c) dw[(size_t)9 * K16 + c] = P.blob[(size_t)st->dwb + c];
    std::vector<float> bias((size_t)Npad, 0.f);
    for (int n = 0; n < st->Cout; ++n) bias[n] = P.blob[(size_t)st->bias + n];
    const int KS = odd_quads(ru(Cin, 4));
    // stride 2: the even- and the odd-column plane of 17 x 17 pixels, each 128-byte aligned (one TMA each)
    const int stage_bytes = st->dws == 1 ? (18 * 18 * KS * 4 + 127) / 128 * 128 : 2 * ((17 * 17 * KS * 4 + 127) / 128 * 128);
    const size_t head = rec.size() * 4 + (size_t)Npad * 4 + 32 * 8 + 256;
    int ns = (int)(((size_t)220 * 1024 - head) / stage_bytes);
    if (ns < 2) return;
    if (ns > 4) ns = 4;
    st->ts_rec = push(rec, rec.size());
    st->ts_bias = push(bias, bias.size());
    st->ts_rec_bytes = (int)(rec.size() * 4);
    st->ts_k16 = K16; st->ts_npad = Npad; st->ts_ks = KS; st->ts_ns = ns; st->ts_stage_bytes = stage_bytes;
    st->ts_res = res; st->ts_relu = st->act == kActRelu ? 1 : 0;
    st->TH = st->dws == 1 ? 16 : 8; st->TW = 16; st->G = 1;
    st->smem = head + (size_t)ns * stage_bytes;
    st->kind = kStepBlockTs;
  }

  // ---- tensor-core variant of a fused pointwise step -------------------------------------------
  // W = hi + lo with hi exactly representable in TF32 (low 13 mantissa bits cleared)
  static void split_w(float v, int parts, float* hi, float* lo) {
    if (parts == 1) { *hi = v; return; }
    uint32_t u;
    std::memcpy(&u, &v, 4);
    u &= 0xFFFFE000u;
    float h;
    std::memcpy(&h, &u, 4);
    *hi = h;
    *lo = v - h;
  }

  static bool tf32_exact(const std::vector<float>& w) {
    for (float v : w) {
      uint32_t u;
      std::memcpy(&u, &v, 4);
      if (u & 0x1FFFu) return false;
    }
    return true;
  }

  // Re-plans `st` (a kStepDwPw) for k_dwpw_tc: M = 128 pixel slots, N = Cout padded to 16, K padded to 8.
  bool plan_tc(PStep* st, int Cin, int OH, int OW, int w_parts) {
    PStep s = *st;
    s.w_parts = w_parts;
    s.K8 = ru(Cin, 8);
    s.Npad = ru(s.Cout, 16);
    if (s.Npad > 128 || s.K8 > 256) return false;
    s.KS = s.K8 + (((s.K8 / 4) % 2 == 0) ? 4 : 0);
    s.tmem_cols = 32;
    while (s.tmem_cols < s.Npad) s.tmem_cols *= 2;
    for (int Pn = 128; Pn >= 32; Pn /= 2) {
      const size_t cap = (size_t)220 * 1024;
      s.TM = 1; s.NPG = Pn;                       // slot budget of this candidate
      int bestTH = 0, bestTW = 0, bestG = 1;
      double best = -1;
      if (OH * OW <= Pn && OW <= 30) {
        bestTH = OH; bestTW = OW; bestG = std::min(63, Pn / (OH * OW));
      } else {
        for (int TW = std::min(std::min(OW, Pn), 30); TW >= 1; --TW) {
          int TH = std::min(std::min(OH, Pn / TW), 30);
          if (TH < 1) continue;
          double tiles = (double)((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
          double util = (double)OH * OW / (tiles * Pn);
          double halo = (double)((TH - 1) * s.dws + 3) * ((TW - 1) * s.dws + 3) / ((double)TH * TW * s.dws * s.dws);
          double score = util / (s.has_dw ? halo : 1.0);
          if (score > best + 1e-9) { best = score; bestTH = TH; bestTW = TW; }
        }
      }
      s.TH = bestTH; s.TW = bestTW; s.G = bestG;
      s.IH = s.has_dw ? (s.TH - 1) * s.dws + 3 : s.TH;
      s.IW = s.has_dw ? (s.TW - 1) * s.dws + 3 : s.TW;
      s.tilesY = (OH + s.TH - 1) / s.TH;
      s.tilesX = (OW + s.TW - 1) / s.TW;
      int nslots = s.G * s.TH * s.TW;
      s.a_rows = ru(nslots, 8);
      s.RS = 1;
      for (int rs : {8, 4, 2}) if (s.TH % rs == 0 && s.G * (s.K8 / 4) * (s.TH / rs) * s.TW >= 256) { s.RS = rs; break; }
      if (s.IH > 63 || s.IW > 63 || s.G > 63 || nslots > 255) continue;     // staging-table field widths
      size_t head = (size_t)s.w_parts * s.Npad * s.K8 + 2 * (size_t)s.Npad + (s.has_dw ? 10 * (size_t)s.K8 : 0);
      size_t a = 2 * (size_t)s.a_rows * s.K8;
      size_t in = (size_t)s.G * s.IH * s.IW * s.KS;
      // the MMA always reads 128 rows of each A tile: keep those addresses inside the allocation
      size_t need_tail = (size_t)(128 - s.a_rows) * s.K8;      // floats past the end of A_lo
      if (in < need_tail) in = need_tail;
      in = (in + 3) / 4 * 4;
      size_t n_chunks = (size_t)s.G * s.IH * s.IW * (s.K8 / 4);
      size_t n_items = s.has_dw ? (size_t)s.G * (s.K8 / 4) * (s.TH / s.RS) * s.TW : 0;
      // double-buffer the input tile when that does not cost residency: the kernel's ~110 registers allow
      // two CTAs per SM, i.e. <= 113 KB each; a CTA that is alone on its SM anyway may use up to 220 KB
      size_t smem1 = (head + a + in) * 4 + (n_chunks + n_items) * 8 + 128;
      size_t smem2 = smem1 + in * 4;
      bool dbl = smem2 <= 113 * 1024 || (smem1 > 113 * 1024 && smem2 <= 220 * 1024);
      s.nbuf = dbl ? 2 : 1;
      s.smem = dbl ? smem2 : smem1;
      if (s.smem <= cap) { *st = s; return true; }      // (plan_ws sets the kind)
    }
    return false;
  }

  // Re-plans a kStepDwPwTc step for k_block_ws (kernels_ws.cu): one CTA per SM, the input tile arrives by TMA
  // into a ring of `ns` stages, the A operand (hi + lo, always 128 rows) is double-buffered when it fits.
  // Shared-memory formula mirrors the kernel's carve-up.
  int res_c_hint = 0;          // channels of the residual tensor of the step being planned (0: unknown / none)
  bool plan_ws(PStep* st, int OH, int OW) {
    constexpr int max_ns = 6, max_na = 4;
    PStep s = *st;
    // staged pixel stride: >= K8 (and >= CoutS when the residual is read from the stage, so the epilogue needs no
    // channel guard: the TMA zero-fills everything past CinS), an odd number of 16-byte quads (conflict-free LDS.128)
    {
      int ks = s.K8;
      if (s.res_mode == 1 && cs(s.Cout) > ks) ks = cs(s.Cout);
      if ((ks / 4) % 2 == 0) ks += 4;
      s.KS = ks;
    }
    s.nt = 4 * s.Npad <= 512 ? 4 : 2;                  // TMEM accumulator ring
    s.tmem_cols = 32;
    while (s.tmem_cols < s.nt * s.Npad) s.tmem_cols *= 2;
    if (s.tmem_cols > 512) return false;
    const size_t cap = (size_t)225 * 1024;
    for (int Pn = 128; Pn >= 32; Pn /= 2) {
      s.TM = 1; s.NPG = Pn;
      int bestTH = 0, bestTW = 0, bestG = 1;
      double best = -1;
      if (OH * OW <= Pn) {
        bestTH = OH; bestTW = OW; bestG = Pn / (OH * OW);
      } else {
        for (int TW = std::min(OW, Pn); TW >= 1; --TW) {
          int TH = std::min(OH, Pn / TW);
          if (TH < 1) continue;
          double tiles = (double)((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
          double util = (double)OH * OW / (tiles * Pn);
          double halo = (double)((TH - 1) * s.dws + 3) * ((TW - 1) * s.dws + 3) / ((double)TH * TW * s.dws * s.dws);
          double score = util / (s.has_dw ? halo : 1.0);
          if (score > best + 1e-9) { best = score; bestTH = TH; bestTW = TW; }
        }
      }
      s.TH = bestTH; s.TW = bestTW; s.G = bestG;
      s.IH = s.has_dw ? (s.TH - 1) * s.dws + 3 : s.TH;
      s.IW = s.has_dw ? (s.TW - 1) * s.dws + 3 : s.TW;
      s.tilesY = (OH + s.TH - 1) / s.TH;
      s.tilesX = (OW + s.TW - 1) / s.TW;
      if (s.IH > 256 || s.IW > 256 || s.G > 256 || s.KS > 256) continue;      // TMA box limits
      s.a_rows = std::min(128, ru(s.G * s.TH * s.TW, 8));      // operand rows really written (the MMA's over-read lands in what follows)
      // Depthwise work split: ND warps (8 with 128 registers, 12 with 96) and RS output rows per item, chosen by an
      // instruction-count estimate per thread and tile; one item per thread lets the taps stay in registers.
      s.RS = 1; s.nd = 8;
      if (s.has_dw) {
        double best_cost = 1e30;
        for (int nd : {12, 8}) {
          for (int rs : {1, 2, 4}) {
            if (s.TH % rs || rs > (nd == 12 ? 2 : 4)) continue;
            int items = s.G * (s.K8 / 4) * (s.TH / rs) * s.TW;
            int m = (items + nd * 32 - 1) / (nd * 32);
            int rows = s.dws == 1 ? rs + 2 : 2 * rs + 1;
            double cost = m * ((m > 1 ? 14 : 0) + 3.0 * rows + 50.0 * rs + 10);
            if (cost < best_cost - 1e-9) { best_cost = cost; s.RS = rs; s.nd = nd; }
          }
        }
      }
      // expand blocks (K <= 16, residual in another HBM tensor): eight depthwise warps are plenty for two to four channel quads
      if (s.has_dw && s.dws == 1 && s.res_mode == 2 && !s.res_pool && ru(s.Cout, 4) <= 32 && s.K8 <= 16) {
        s.nd = 8;
        int items = s.G * (s.K8 / 4) * s.TH * s.TW;
        s.RS = (items > 256 && s.TH % 2 == 0) ? 2 : 1;
        if (s.G * (s.K8 / 4) * (s.TH / s.RS) * s.TW > 256 && s.TH % 4 == 0) s.RS = 4;
      }
      size_t n_items = s.has_dw ? (size_t)s.G * (s.K8 / 4) * (s.TH / s.RS) * s.TW : 0;
      size_t head = ((size_t)s.w_parts * s.Npad * s.K8 + 2 * (size_t)s.Npad + (s.has_dw ? 10 * (size_t)s.K8 : 0)) * 4 + (n_items + 1) / 2 * 8 + 48 * 8 + 128;
      size_t a_bytes = (size_t)2 * s.a_rows * s.K8 * 4;
      size_t in_bytes = ((size_t)s.G * s.IH * s.IW * s.KS * 4 + 127) / 128 * 128;
      // stride 2 (even input, no padding before): even and odd columns as two planes, one TMA each (element stride 2 along x), so
      // that neighbouring output columns read neighbouring records - conflict-free LDS.128 (see k_block_ts)
      s.deint = (s.has_dw && s.dws == 2 && s.dpt == 0 && s.dpl == 0 && (s.IW & 1)) ? 1 : 0;
      s.PW = (s.IW + 1) / 2;
      if (s.deint) {
        const size_t plane = ((size_t)s.G * s.IH * s.PW * s.KS * 4 + 127) / 128 * 128;
        s.plane_floats = (int)(plane / 4);
        in_bytes = 2 * plane;
      }
      // Output tile for the TMA-store epilogue: pixel stride KSo = CoutS rounded to an odd number of quads (conflict-free
      // STS.128), one or two buffers.  Ring choice: (A buffers, input stages) with even A rings preferred (the two MMA
      // issuers alternate tiles); output buffers first, since the direct-store epilogue saturates the LSU.
      static const int want_no = [] { const char* e = std::getenv("FDT_WS_NO"); return e ? std::atoi(e) : 0; }();   // measured: direct stores are as fast; TMA store via FDT_WS_NO=2
      const int couts = cs(s.Cout);
      s.KSo = ((couts / 4) % 2 == 0) ? couts + 4 : couts;
      const size_t out_bytes = ((size_t)s.G * s.TH * s.TW * s.KSo * 4 + 127) / 128 * 128;
      const bool can_tma_out = s.KSo <= 256;
      static const int combos[10][2] = {{4, 6}, {4, 5}, {4, 4}, {2, 5}, {2, 4}, {2, 3}, {2, 2}, {1, 2}, {1, 1}, {0, 0}};
      // preference: two output tiles + prefetching input ring (ns >= 2), then one output tile + prefetching ring, then any
      // ring with an output tile, then direct stores
      static const int prefs[5][2] = {{2, 2}, {1, 2}, {2, 1}, {1, 1}, {0, 1}};     // (output tiles, min input stages)
      // Residual in ANOTHER HBM tensor (the expand blocks of the full-range / iris nets): staged by TMA into a ring of its own, so that
      // several tiles' worth of residual are in flight (ncu: with direct loads the four epilogue warps keep 8 KB per SM on the way and
      // the kernel idles on long scoreboard at 23 % SM throughput).  Same-size or exactly 2x2-pooled residuals only.
      size_t res_bytes = 0;
      int want_nr = 0;
      s.nr = 0; s.KSr = 0; s.res_stage_floats = 0;
      if (s.res_mode == 2 && s.c2 == 0) {
        const int m = s.res_pool ? 2 : 1;
        // record stride: the output's channels (no channel guard in the epilogue); a pooled tile holds four times the pixels, so it
        // keeps only the residual's own channels and the epilogue guards the rest (res_C arrives with the launch: KSr is fixed up there)
        s.KSr = odd_quads(ru(s.Cout, 4));
        if (s.res_pool && res_c_hint > 0) s.KSr = odd_quads(ru(res_c_hint, 4));
        if (s.KSr <= 256 && m * s.TW <= 256 && m * s.TH <= 256) {
          res_bytes = ((size_t)s.G * (m * s.TH) * (m * s.TW) * s.KSr * 4 + 127) / 128 * 128;
          want_nr = 3;
        }
      }
      for (const auto& pr : prefs) {
        const int no = pr[0];
        if (no > want_no || (no > 0 && (!can_tma_out || s.c2 > 0))) continue;
        for (const auto& c : combos) {
          if (c[0] == 0 || c[0] > max_na || c[1] > max_ns || c[1] < pr[1]) continue;
          size_t total = head + c[0] * a_bytes + c[1] * in_bytes + no * out_bytes;
          int nr = 0;
          if (want_nr && no == 0) {
            // the ring takes what the input / operand rings leave, two to four stages; rings one step smaller are preferred over none
            nr = (int)std::min<size_t>(4, total < cap ? (cap - total) / res_bytes : 0);
            if (nr < 2) { if (c[0] > 1 || c[1] > 2) continue; nr = 0; }
            total += (size_t)nr * res_bytes;
          }
          // the MMA reads 128 operand rows: its over-read past the last A buffer must stay inside the allocation
          const size_t over = (size_t)(128 - s.a_rows) * s.K8 * 4, after = c[1] * in_bytes + no * out_bytes;
          if (over > after) total += over - after;
          if (total > cap) continue;
          s.na = c[0]; s.ns = c[1]; s.no = no;
          s.nr = nr; s.res_stage_floats = (int)(res_bytes / 4);
          s.in_stage_floats = (int)(in_bytes / 4);
          s.out_stage_floats = (int)(out_bytes / 4);
          s.smem = total;
          s.kind = kStepBlockWs;
          *st = s;
          return true;
        }
      }
    }
    return false;
  }

  std::vector<int> fused_outputs;
  bool is_output_of_fused(int tf) const {
    return std::find(fused_outputs.begin(), fused_outputs.end(), tf) != fused_outputs.end();
  }

  bool plan_spatial(PStep* s, int OH, int OW) {
    for (int attempt = 0; attempt < 2; ++attempt) {
      int Pn = s->TM * s->NPG;
      int bestTH = 0, bestTW = 0, bestG = 1;
      double best = -1;
      if (OH * OW <= Pn) {
        bestTH = OH; bestTW = OW; bestG = Pn / (OH * OW);
      } else {
        for (int TW = std::min(OW, Pn); TW >= 1; --TW) {
          int TH = std::min(OH, Pn / TW);
          if (TH < 1) continue;
          double tiles = (double)((OH + TH - 1) / TH) * ((OW + TW - 1) / TW);
          double util = (double)OH * OW / (tiles * Pn);
          double halo = (double)((TH - 1) * s->dws + 3) * ((TW - 1) * s->dws + 3) / ((double)TH * TW * s->dws * s->dws);
          double score = util / (s->has_dw ? halo : 1.0);
          if (score > best + 1e-9) { best = score; bestTH = TH; bestTW = TW; }
        }
      }
      s->TH = bestTH; s->TW = bestTW; s->G = bestG;
      s->IH = s->has_dw ? (s->TH - 1) * s->dws + 3 : s->TH;
      s->IW = s->has_dw ? (s->TW - 1) * s->dws + 3 : s->TW;
      s->tilesY = (OH + s->TH - 1) / s->TH;
      s->tilesX = (OW + s->TW - 1) / s->TW;
      size_t in_elems = (size_t)s->G * s->IH * s->IW * s->KS;
      size_t a_elems = (size_t)Pn * s->KS;
      size_t floats = s->has_dw ? in_elems + a_elems : std::max(in_elems, a_elems);
      floats += (size_t)s->KP * s->NC;
      s->smem = floats * 4;
      if (s->smem <= kSmemLimit) return true;
      if (s->NPG % 2 == 0 && attempt == 0) s->NPG /= 2; else break;   // retry with 64-pixel tiles
    }
    return false;
  }

  // ---- dense k x k conv as im2col GEMM ---------------------------------------------------------
  bool analyse_dense(int ci, Fused* F) {
    const TfOp& conv = m.ops[ci];
    const TfTensor& wt = m.tensors[conv.in[1]];
    const TfTensor& it = m.tensors[conv.in[0]];
    const TfTensor& ot = m.tensors[conv.out[0]];
    if (wt.shape.size() != 4 || it.shape.size() != 4 || conv.dil_h != 1 || conv.dil_w != 1) return false;
    if (conv.act != 0 && conv.act != 1) return false;
    PStep st;
    st.kind = kStepGemmConv;
    st.kh = wt.shape[1]; st.kw = wt.shape[2];
    st.sh = conv.stride_h; st.sw = conv.stride_w;
    int Cin = wt.shape[3];
    st.Cout = wt.shape[0];
    // the window is the whole map (VALID, 1x1 output): one dense contraction per image -> k_fc_tc (tcgen05 GEMM over the chunk)
    if (use_tc && conv.padding == 1 && st.kh == it.dim(1) && st.kw == it.dim(2) && ot.numel() == st.Cout && conv.act == 0 &&
        !is_view(conv.in[0]) && Cin % 4 == 0 && Cin == it.dim(3) && (st.kh * st.kw * Cin) % 16 == 0 && st.kh * st.kw * Cin <= 384) {
      std::vector<float> w, b;
      if (!m.const_f32(conv.in[1], &w)) return fail("conv weights are not constant");
      if (conv.in.size() > 2 && conv.in[2] >= 0) { if (!m.const_f32(conv.in[2], &b)) return fail("conv bias is not constant"); }
      b.resize(st.Cout, 0.f);
      st.kind = kStepFcTc;
      st.K = st.kh * st.kw * Cin;
      st.K8 = st.K;                                  // K16
      st.w_parts = f16_exact(w) ? 1 : 2;
      float scale = 1.f, wsc = 1.f;
      if (st.w_parts == 2) {
        float mx = 0.f;
        for (float v : w) mx = std::max(mx, std::fabs(v));
        if (mx > 0.f) { int e; std::frexp(mx, &e); scale = std::ldexp(1.f, 14 - e); }
      }
      const int ntile = (st.Cout + 127) / 128;
      std::vector<float> rec;
      for (int tI = 0; tI < ntile; ++tI) {
        std::vector<float> r1 = pack_w_f16(w, std::min(128, st.Cout - tI * 128), st.K, 128, st.K, st.w_parts, &wsc, scale, tI * 128);
        rec.insert(rec.end(), r1.begin(), r1.end());
      }
      st.ts_rec_bytes = st.w_parts * 128 * st.K * 2;
      st.out_scale = 1.0 / (double)scale;
      st.w = push(rec, rec.size());
      st.bias = push(b, b.size() + 8);
      st.smem = (size_t)st.ts_rec_bytes + 256;
      st.act = kActNone;
      done[ci] = 1;
      st.macs = (double)st.K * st.Cout;
      st.name = m.tensors[conv.out[0]].name;
      F->ok = true;
      F->st = st;
      F->src_tf = conv.in[0];
      F->out_tf = conv.out[0];
      fused_outputs.push_back(conv.out[0]);
      return true;
    }
    if (conv.padding == 0) { same_pad(it.dim(1), st.kh, st.sh, &st.pt); same_pad(it.dim(2), st.kw, st.sw, &st.pl); }
    st.K = st.kh * st.kw * Cin;
    st.KP = ru(st.K, 4);
    st.KS = smem_stride(st.KP);
    plan_columns(&st);
    // A (P x KS) + W (KP x NC) must fit in shared memory: every weight chunk re-stages the im2col tile,
    // so shrink the pixel tile (128 -> 64 -> 32) before splitting the output channels any further.
    {
      bool fit = false;
      PStep best = st;
      int best_chunks = 1 << 30;
      for (int P = 128; P >= 32; P /= 2) {
        PStep c = st;
        plan_columns(&c);
        for (;;) {
          c.TM = (c.NC / 4) <= 12 && P >= 128 ? 4 : (P >= 128 ? 8 : 4);
          c.NPG = P / c.TM;
          c.smem = ((size_t)P * c.KS + (size_t)c.KP * c.NC) * 4;
          if (c.smem <= kSmemLimit && c.NPG * (c.NC / 4) <= 384) break;
          int nq = c.NC / 4;
          if (nq <= 1) { c.smem = 0; break; }
          nq = (nq + 1) / 2;
          c.NC = nq * 4;
          c.nchunks = (c.CoutP + c.NC - 1) / c.NC;
        }
        if (c.smem == 0) continue;
        if (c.nchunks < best_chunks) { best_chunks = c.nchunks; best = c; fit = true; }
      }
      if (!fit) return false;
      st = best;
    }
    std::vector<int> absorbed;
    int cur = conv.out[0], alpha_tf = -1;
    st.act = conv.act == 1 ? kActRelu : kActNone;
    if (st.act == kActNone) fuse_activation(&cur, &st.act, &alpha_tf, &absorbed);
    st.in_u8 = conv.in[0] == m.inputs[0] && Cin == 3;
    if (st.in_u8 && st.kh == st.kw && (st.kw == 3 || st.kw == 5) && st.sh == 2 && st.sw == 2 && ru(st.Cout, 4) <= 64) {
      st.kind = kStepStem;                       // direct strip convolution from the staged patch
      st.CoutP = ru(st.Cout, 4);
      st.NC = st.CoutP;
      st.nchunks = 1;
      int PH = 14 + st.kw, PW = 30 + st.kw;
      st.smem = (size_t)PH * PW * 16 + (size_t)st.KP * st.NC * 4;
    }
    for (int a : absorbed) done[a] = 1;
    done[ci] = 1;
    std::vector<float> w, b;
    if (!m.const_f32(conv.in[1], &w)) return fail("conv weights are not constant");
    if (conv.in.size() > 2 && conv.in[2] >= 0) { if (!m.const_f32(conv.in[2], &b)) return fail("conv bias is not constant"); }
    b.resize(st.Cout, 0.f);
    if (st.kind == kStepStem && use_tc && ru(st.Cout, 16) <= 64) {
      // k_stem_ws: im2col GEMM with exact fp16 operands.  A holds (u8 - 127.5) (half-integers, exact in fp16), laid out
      // per tap row ky as SEGP pixels x {B,G,R,X} halves; W is scaled by a power of two and split into w_parts fp16
      // terms (1 when the weights are fp16-origin, 3 for fp32 weights); out = D * out_scale + bias.
      st.kind = kStepStemWs;
      const int SEGP = st.kw == 5 ? 6 : 4, CPK = SEGP / 2;       // pixels / 16-byte chunks per tap row
      const int nchunk = ru(st.kw * CPK, 2);                       // K chunks of 8 halves, even (MMA K = 16)
      st.K8 = nchunk * 8;
      st.Npad = ru(st.Cout, 16);
      st.tmem_cols = 32;
      while (st.tmem_cols < 4 * st.Npad) st.tmem_cols *= 2;
      float wmax = 0.f;
      for (float v : w) wmax = std::max(wmax, std::fabs(v));
      float wscale = 1.f;
      while (wmax * wscale * 2.f <= 16384.f && wscale < 65536.f) wscale *= 2.f;
      // parts: how many fp16 terms reproduce every scaled weight exactly (up to 3)
      int parts = 1;
      for (float v : w) {
        float r = v * wscale;
        int need = 0;
        for (int t = 0; t < 3 && r != 0.f; ++t) { r -= f16_to_f32(f32_to_f16(r)); need = t + 1; }
        parts = std::max(parts, std::max(need, 1));
      }
      st.w_parts = parts;
      st.out_scale = 1.0 / (127.5 * (double)wscale);
      const size_t SBO = (size_t)nchunk * 128, LBO = 128;          // bytes
      std::vector<uint16_t> hb((size_t)parts * st.Npad * st.K8, 0);
      for (int n = 0; n < st.Cout; ++n)
        for (int ky = 0; ky < st.kw; ++ky)
          for (int kx = 0; kx < st.kw; ++kx)
            for (int c = 0; c < 3; ++c) {                          // c: 0=B 1=G 2=R in the letterboxed bytes; weights are RGB
              const int k = (ky * CPK + kx / 2) * 8 + (kx & 1) * 4 + c;
              float r = w[(((size_t)n * st.kw + ky) * st.kw + kx) * 3 + (2 - c)] * wscale;
              const size_t o = ((size_t)(n >> 3) * SBO + (size_t)(k >> 3) * LBO + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2) / 2;
              for (int t = 0; t < parts; ++t) {
                const uint16_t hbits = f32_to_f16(r);
                hb[(size_t)t * st.Npad * st.K8 + o] = hbits;
                r -= f16_to_f32(hbits);
              }
            }
      std::vector<float> wb((hb.size() + 1) / 2, 0.f);
      std::memcpy(wb.data(), hb.data(), hb.size() * 2);
      const int PH = 14 + st.kw;
      st.ns = 6; st.na = 2;   // must match FDT_STEM_NA in kernels_ws.cu
      st.smem = (size_t)parts * st.Npad * st.K8 * 2 + 2 * (size_t)st.Npad * 4 + 32 * 8 + 128     // W, bias, alpha, barriers
                + (size_t)st.na * 128 * st.K8 * 2 + 2 * (size_t)PH * 36 * 8 + (size_t)st.ns * ((PH * 160 + 127) / 128 * 128) + 256;
      st.w = push(wb, wb.size());
      st.bias = push(b, (size_t)st.Npad + 8);
      if (alpha_tf >= 0 && !pack_alpha(alpha_tf, (size_t)st.Npad + 8, &st.alpha)) return false;
    } else {
      std::vector<float> wk((size_t)st.KP * st.CoutP, 0.f);
      for (int co = 0; co < st.Cout; ++co)
        for (int k = 0; k < st.K; ++k) wk[(size_t)k * st.CoutP + co] = w[(size_t)co * st.K + k];
      size_t padded = (size_t)st.nchunks * st.NC + 8;
      st.w = push(wk, wk.size());
      st.bias = push(b, padded);
      if (alpha_tf >= 0 && !pack_alpha(alpha_tf, padded, &st.alpha)) return false;
    }
    st.macs = (double)ot.dim(1) * ot.dim(2) * st.K * st.Cout;
    st.name = m.tensors[cur].name;
    F->ok = true;
    F->st = st;
    F->src_tf = conv.in[0];
    F->out_tf = cur;
    fused_outputs.push_back(cur);
    return true;
  }

  // ---- image-resident tail ---------------------------------------------------------------------------
  // The maximal suffix of the plan made of warp-specialised BlazeBlock / head steps on feature maps of at most 256
  // pixels becomes ONE k_tail_ws step (kernels_tail.cu): activations stay in shared memory from the first block to
  // the heads, weights (exact in fp16: the detectors ship fp16 weights) stream in per layer.  Anything the kernel
  // does not cover (fp32-only weights, residuals from other tensors, branching chains, > 128 channels, more than
  // kTailMaxLayers layers) leaves the plan as it is.
  static int odd_quads(int c) { int q = ru(c, 4) / 4; if (q % 2 == 0) ++q; return q * 4; }

  // fp16 K-major core-matrix image of W [N][K] (8 rows x 16 bytes per core matrix, row groups SBO = (K16 / 8) * 128 bytes apart), as
  // floats; parts == 2: W * 2^s = hi + lo, both fp16 (the second image follows the first), s chosen so that the largest weight sits
  // near 2^13 (no fp16 subnormals among the significant ones); *wscale = 2^-s.  parts == 1 requires exact fp16 weights (s = 0).
  std::vector<float> pack_w_f16(const std::vector<float>& w, int N, int K, int Npad, int K16, int parts, float* wscale, float force_scale = 0.f, int n_first = 0) {
    const size_t SBO16 = (size_t)(K16 / 8) * 128;
    std::vector<uint16_t> img((size_t)parts * Npad * K16, 0);
    float scale = 1.f;
    if (parts == 2) {
      float mx = 0.f;
      for (float v : w) mx = std::max(mx, std::fabs(v));
      if (mx > 0.f) { int e; std::frexp(mx, &e); scale = std::ldexp(1.f, 14 - e); }
      if (force_scale > 0.f) scale = force_scale;
    }
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < K; ++k) {
        const float v = w[(size_t)(n_first + n) * K + k] * scale;
        const size_t o = ((size_t)(n >> 3) * SBO16 + (size_t)(k >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2) / 2;
        const uint16_t hb = f32_to_f16(v);
        img[o] = hb;
        if (parts == 2) img[(size_t)Npad * K16 + o] = f32_to_f16(v - f16_to_f32(hb));
      }
    *wscale = 1.f / scale;
    std::vector<float> rec((img.size() + 1) / 2, 0.f);
    std::memcpy(rec.data(), img.data(), img.size() * 2);
    return rec;
  }
  static bool f16_exact(const std::vector<float>& w) {
    for (float v : w) if (f16_to_f32(f32_to_f16(v)) != v) return false;
    return true;
  }

  // Shared-memory geometry of a k_tail_ws step from its layers and buffers (pixels, widest tenant); false: does not fit.
  bool finish_tail(PStep* t, const std::vector<int>& buf_px, const std::vector<int>& buf_c) {
    if (buf_px.empty() || buf_px.size() > (size_t)kTailMaxBufs || t->tail.empty() || (int)t->tail.size() > kTailMaxLayers) return false;
    int off = 0;
    t->tail_nbuf = (int)buf_px.size();
    for (size_t b = 0; b < buf_px.size(); ++b) {
      const int ks = odd_quads(buf_c[b]);
      t->tail_buf_off[b] = off; t->tail_buf_ks[b] = ks; t->tail_buf_px[b] = buf_px[b];
      off += buf_px[b] * ks;
    }
    if (t->tail_buf_ks[0] > 256 && !t->tail_wide) return false;                // TMA box limit (k_chain_wide copies pixel by pixel)
    t->tail_act_floats = off;
    t->tail_in_bytes = buf_px[0] * t->tail_buf_ks[0] * 4;
    size_t wbuf = 0, tbuf = 0;
    int last0 = 0;
    for (size_t li = 0; li < t->tail.size(); ++li) {
      const TailLayerD& L = t->tail[li];
      wbuf = std::max(wbuf, (size_t)L.rec_bytes);
      tbuf = std::max(tbuf, (size_t)L.tap_bytes);
      if (L.rec_bytes % 16 || L.tap_bytes % 16) return false;
      if (L.src == 0 || L.dst == 0 || (L.res && L.rbuf == 0)) last0 = (int)li;
    }
    t->tail_last_a = last0;
    t->tail_tbuf = (int)((tbuf + 127) / 128 * 128);
    if (t->tail_wide) {
      // k_chain_wide: ring of W blocks (as deep as fits, at least the blocks one K chunk needs at once), bias area by offset
      int need = 2;
      for (const TailLayerD& L : t->tail) need = std::max(need, L.ng >> 8);
      for (const TailBlk& b : t->tail_blks) wbuf = std::max(wbuf, (size_t)b.bytes);
      t->tail_wbuf = (int)((wbuf + 127) / 128 * 128);
      if (t->tail_blks.size() > (size_t)kTailMaxBlks) return false;
      auto wbytes = [&](int depth) {      // = wide_smem_bytes() in kernels_tail.cu
        return (size_t)kTailMaxLayers * sizeof(TailLayerD) + (size_t)kTailMaxBlks * sizeof(TailBlk) + 2 * (size_t)t->tail_bias_floats * 4 + 32 * 8 + 16 * 4 + 128 +
               ((size_t)off + 512) * 4 + 128 + 2 * (size_t)t->tail_tbuf + (size_t)depth * t->tail_wbuf;
      };
      int depth = 4;
      while (depth > need && wbytes(depth) > (size_t)225 * 1024) --depth;
      t->tail_wdepth = depth;
      t->smem = wbytes(depth);
      return t->smem <= (size_t)225 * 1024;
    }
    t->tail_wbuf = (int)((wbuf + 127) / 128 * 128);
    auto bytes = [&](int depth) {       // = tail_smem_bytes() in kernels_tail.cu
      return (size_t)kTailMaxLayers * sizeof(TailLayerD) + 2 * (size_t)kTailMaxLayers * 512 + 16 * 8 + 16 * 4 + 128 +
             ((size_t)off + 128) * 4 + 128 + 2 * (size_t)t->tail_tbuf + (size_t)depth * t->tail_wbuf;
    };
    t->tail_wdepth = bytes(2) <= (size_t)225 * 1024 ? 2 : 1;
    t->smem = bytes(t->tail_wdepth);
    return t->smem <= (size_t)225 * 1024;
  }

  // Replaces steps [f, end) by the tail step and redoes the liveness bookkeeping.
  void redo_liveness() {
    for (PTensor& x : P.tensors) { x.materialized = false; x.def_step = -1; x.last_use = -1; }
    for (size_t i = 0; i < P.steps.size(); ++i) {
      const PStep& s = P.steps[i];
      P.tensors[s.out].materialized = true;
      if (s.out2 >= 0) P.tensors[s.out2].materialized = true;
      for (int e : s.extra_out) { P.tensors[e].materialized = true; use((int)i, e); }
      for (int e : s.tail_rsrc) use((int)i, e);
      if (!s.in_u8) use((int)i, s.in);
      use((int)i, s.in2);
      use((int)i, s.out);
      use((int)i, s.out2);
    }
  }

  void fuse_tail() {
    const int S = (int)P.steps.size();
    auto block_ok = [&](const PStep& s) {
      if ((s.kind != kStepBlockWs && s.kind != kStepBlockTs) || s.w_parts != 1 || s.in < 0) return false;
      const PTensor& in = P.tensors[s.in];
      const PTensor& out = P.tensors[s.out];
      if (in.H * in.W > 256 || in.root >= 0 || in.C > 128 || s.Cout > 128) return false;
      if (s.has_dw) {
        if (s.c2 > 0 || out.root >= 0 || (s.act != kActRelu && s.act != kActNone)) return false;
        if (s.in2 >= 0 && (s.res_mode != 1 || s.in2 != s.in)) return false;
        if (s.dws != 1 && s.dws != 2) return false;
        if (s.dws == 2 && (in.H % 2 || in.W % 2 || s.dpt != 0 || s.dpl != 0)) return false;
        if (s.dws == 1 && (s.dpt != 1 || s.dpl != 1)) return false;
        return true;
      }
      // head (pair): pointwise only, dense graph-output views
      if (s.in2 >= 0 || s.act != kActNone || out.root < 0) return false;
      if (s.c2 > 0 ? (s.c1 % 4 != 0 || s.out2 < 0 || P.tensors[s.out2].root < 0) : (s.Cout % 4 != 0)) return false;
      return true;
    };
    int f = S;
    while (f > 0 && block_ok(P.steps[f - 1])) --f;
    // the chain must start at a block and be closed: inputs come from inside it (or are the first block's input)
    for (; f < S; ++f) {
      if (!P.steps[f].has_dw) continue;
      const int X = P.steps[f].in;
      std::vector<int> produced{X};
      bool ok = true;
      for (int i = f; i < S && ok; ++i) {
        ok = std::find(produced.begin(), produced.end(), P.steps[i].in) != produced.end();
        produced.push_back(P.steps[i].out);
      }
      // nothing before the chain may be read after it started, and no tensor of the chain feeds two blocks
      for (int i = f; i < S && ok; ++i)
        for (int j = i + 1; j < S && ok; ++j)
          if (P.steps[i].has_dw && P.steps[j].has_dw && P.steps[i].in == P.steps[j].in) ok = false;
      if (ok) break;
    }
    if (S - f < 3) return;
    const int X = P.steps[f].in;
    // layer order: blocks as planned, every head right after the block that produces its input
    std::vector<int> order;
    for (int i = f; i < S; ++i) {
      if (!P.steps[i].has_dw) continue;
      order.push_back(i);
      for (int j = f; j < S; ++j)
        if (!P.steps[j].has_dw && P.steps[j].in == P.steps[i].out) order.push_back(j);
    }
    if ((int)order.size() != S - f || (int)order.size() > kTailMaxLayers) return;
    // buffers: 0 until the stride-2 block, 1 after it
    std::map<int, int> buf_of;
    buf_of[X] = 0;
    int maxA = P.tensors[X].Cs, maxB = 4, PA = P.tensors[X].H * P.tensors[X].W, PB = 1, n_s2 = 0;
    PStep t;
    t.kind = kStepTailWs;
    for (size_t li = 0; li < order.size(); ++li) {
      const PStep& s = P.steps[order[li]];
      const PTensor& in = P.tensors[s.in];
      const PTensor& out = P.tensors[s.out];
      auto bi = buf_of.find(s.in);
      if (bi == buf_of.end()) return;
      TailLayerD L = {};
      L.kind = s.has_dw ? 0 : 1;
      L.src = bi->second;
      L.rbuf = L.src;
      L.IH = in.H; L.IW = in.W; L.OH = out.H; L.OW = out.W;
      L.Cin = in.C; L.Cout = s.Cout; L.K16 = ru(in.C, 16); L.Npad = ru(s.Cout, 16);
      L.o1 = L.o2 = -1;
      L.w_parts = 1; L.wscale = 1.f; L.alpha_off = 0;
      if (s.has_dw) {
        L.stride = s.dws; L.pad = s.dpt;
        L.res = s.in2 >= 0 ? (s.res_pool ? 2 : 1) : 0;
        L.act = s.act == kActRelu ? 1 : 0;
        if (s.dws == 2) {
          if (L.src != 0 || ++n_s2 > 1) return;
          L.dst = 1;
          PB = out.H * out.W;
        } else {
          L.dst = L.src;
          if (L.res == 2) return;
        }
        buf_of[s.out] = L.dst;
        (L.dst ? maxB : maxA) = std::max(L.dst ? maxB : maxA, ru(s.Cout, 4));
      } else {
        L.dst = -1;
        L.c1 = s.c2 > 0 ? s.c1 : s.Cout; L.c2 = s.c2;
        L.o1 = (int)t.tail_outs.size(); t.tail_outs.push_back(s.out);
        if (s.c2 > 0) { L.o2 = (int)t.tail_outs.size(); t.tail_outs.push_back(s.out2); }
        if (t.tail_outs.size() > 4) return;
      }
      // the kernel's M-tile maps: at most 128 pixels (one tile) or an even-height map of at most 256 (two tiles: even / odd rows)
      if (!(L.OH * L.OW <= 128 || (L.OH * L.OW <= 256 && L.OH % 2 == 0 && (!s.has_dw || s.dws == 1)))) return;
      // the step's tensor-core operands back to plain arrays: W [Cout][Cin] (exact values), taps [9][K8], biases
      const size_t SBO8 = (size_t)(s.K8 / 4) * 128;
      const int K16 = L.K16;
      std::vector<float> wplain((size_t)s.Cout * in.C);
      for (int n = 0; n < s.Cout; ++n)
        for (int k = 0; k < in.C; ++k)
          wplain[(size_t)n * in.C + k] = P.blob[(size_t)s.w + ((size_t)(n >> 3) * SBO8 + (size_t)(k >> 2) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 3) * 4) / 4];
      if (!f16_exact(wplain)) return;                           // not an fp16-origin model: keep the TF32 hi/lo kernels
      std::vector<float> rec = pack_w_f16(wplain, s.Cout, in.C, L.Npad, K16, 1, &L.wscale);
      L.rec_bytes = (int)(rec.size() * 4);
      L.rec_off = (int)push(rec, rec.size());
      if (s.has_dw) {
        std::vector<float> dw((size_t)10 * K16, 0.f);
        for (int k = 0; k < 9; ++k)
          for (int c = 0; c < in.C; ++c) dw[(size_t)k * K16 + c] = P.blob[(size_t)s.dww + (size_t)k * s.K8 + c];
        for (int c = 0; c < in.C; ++c) dw[(size_t)9 * K16 + c] = P.blob[(size_t)s.dwb + c];
        L.tap_bytes = (int)(dw.size() * 4);
        L.tap_off = (int)push(dw, dw.size());
      }
      // pointwise bias [Npad]: the step's own array is zero-padded past Cout already (Npad16 >= Npad here)
      std::vector<float> bias((size_t)L.Npad, 0.f);
      for (int n = 0; n < s.Cout; ++n) bias[n] = P.blob[(size_t)s.bias + n];
      L.bias_off = (int)push(bias, bias.size());
      t.macs += s.macs;
      t.tail.push_back(L);
    }
    if (t.tail_outs.empty() || !finish_tail(&t, {PA, PB}, {maxA, maxB})) return;
    t.in = X;
    t.out = t.tail_outs[0];
    for (size_t k = 1; k < t.tail_outs.size(); ++k) t.extra_out.push_back(t.tail_outs[k]);
    t.name = "tail:" + P.steps[order.front()].name + ".." + P.steps[order.back()].name;
    // replace the suffix and redo the liveness bookkeeping
    P.steps.resize(f);
    P.steps.push_back(t);
    redo_liveness();
  }

  // ---- image-resident chain found on the GRAPH (before the per-conv planning) --------------------------------------
  // The face-landmark net below 12x12: seventeen small layers (two branches, PReLU, fp32 weights) that the per-layer
  // kernels run at 1-10 % of any roofline.  From the first block whose input map fits one CTA (<= 256 pixels, <= 128
  // channels) the matcher follows every [DW3x3 ->] PW1x1 [-> + residual] [-> ReLU / PReLU] pattern whose inputs live
  // inside the chain, plus whole-map convolutions with one filter (a dot product), assigns shared-memory buffers by
  // liveness (in place where the source dies) and emits ONE k_tail_ws step.  Tensors that are read outside the chain are
  // also written to HBM by the layer that produces them.  A chain that would end in 1x1 heads is left to fuse_tail().
  struct ChainOut { int emit_at; PStep step; };
  std::vector<ChainOut> chains;
  void fuse_chain() {
    static const int want = [] { const char* e = std::getenv("FDT_CHAIN"); return e ? std::atoi(e) : 1; }();   // FDT_CHAIN=0: A/B, one launch per layer
    if (!want) return;
    bool narrow_done = false;                // one narrow chain per graph (the face-landmark trunk), any number of wide ones
    for (size_t s0 = 0; s0 < m.ops.size(); ++s0)
      if (m.ops[s0].code == kOpConv2D && !done[s0]) try_chain((int)s0, &narrow_done);
  }

  struct ChainLayer { int conv; PwMatch M; int kind; int out_tf; };

  // one W block of k_chain_wide: rows [n0, n0 + ncols) x K range [k0, k0 + kw) of W [N][K] (zero outside), fp16 K-major core matrices,
  // parts == 2: the lo image follows the hi image; `scale` as in pack_w_f16
  std::vector<float> pack_w_block(const std::vector<float>& w, int N, int K, int n0, int ncols, int k0, int kw, int parts, float scale) {
    const size_t SBO16 = (size_t)(kw / 8) * 128;
    std::vector<uint16_t> img((size_t)parts * ncols * kw, 0);
    for (int n = 0; n < ncols && n0 + n < N; ++n)
      for (int k = 0; k < kw && k0 + k < K; ++k) {
        const float v = w[(size_t)(n0 + n) * K + k0 + k] * scale;
        const size_t o = ((size_t)(n >> 3) * SBO16 + (size_t)(k >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2) / 2;
        const uint16_t hb = f32_to_f16(v);
        img[o] = hb;
        if (parts == 2) img[(size_t)ncols * kw + o] = f32_to_f16(v - f16_to_f32(hb));
      }
    std::vector<float> rec((img.size() + 1) / 2, 0.f);
    std::memcpy(rec.data(), img.data(), img.size() * 2);
    return rec;
  }

  void try_chain(int s0, bool* narrow_done) {
    PwMatch M0;
    if (!match_pointwise(s0, &M0, true)) return;
    const int T0 = M0.src;
    // a layer no per-layer tensor-core kernel takes (more than 128 channels on either side) opens a WIDE chain (k_chain_wide):
    // up to 384 channels, residuals from HBM tensors, any number of layers down to one
    const bool wide = m.tensors[T0].shape.size() == 4 && (m.tensors[T0].dim(3) > 128 || m.tensors[m.ops[s0].in[1]].shape[0] > 128);
    if (!wide && (!M0.has_dw || *narrow_done)) return;
    const int cmax = wide ? 384 : 128;
    {
      const TfTensor& t0 = m.tensors[T0];
      if (is_view(T0) || t0.shape.size() != 4 || t0.dim(1) * t0.dim(2) > 256 || t0.dim(3) > cmax || m.producer(T0) < 0) return;
    }
    const int first_op = M0.has_dw ? M0.absorbed[0] : s0;
    std::vector<int> chainT{T0};
    auto in_chain = [&](int t) { return std::find(chainT.begin(), chainT.end(), t) != chainT.end(); };
    // a residual read from HBM: a plain NHWC tensor produced before the chain starts
    auto hbm_res_ok = [&](int t) {
      const int pr = m.producer(t);
      return wide && !is_view(t) && m.tensors[t].shape.size() == 4 && pr >= 0 && pr < first_op && m.tensors[t].dim(3) <= cmax;
    };
    std::vector<ChainLayer> layers;
    // pass A1: the convolutions, in graph order
    for (size_t j = s0; j < m.ops.size(); ++j) {
      const TfOp& op = m.ops[j];
      if (op.code != kOpConv2D || done[j]) continue;
      PwMatch M;
      if (match_pointwise((int)j, &M, true) && in_chain(M.src) && (M.res < 0 || in_chain(M.res) || hbm_res_ok(M.res))) {
        const TfTensor& it = m.tensors[M.src];
        const TfTensor& ot = m.tensors[M.cur];
        const int Cout = m.tensors[op.in[1]].shape[0];
        bool ok = it.shape.size() == 4 && ot.shape.size() == 4 && !is_view(M.cur) && it.dim(3) <= cmax && Cout <= cmax;
        const int npix = ok ? ot.dim(1) * ot.dim(2) : 0;
        ok = ok && (npix <= 128 || (npix <= 256 && ot.dim(1) % 2 == 0 && (!M.has_dw || M.dws == 1)));
        if (ok && M.has_dw) {
          if (M.dws == 1) ok = M.dpt == 1 && M.dpl == 1;
          else ok = it.dim(1) % 2 == 0 && it.dim(2) % 2 == 0 && M.dpt == 0 && M.dpl == 0;
        }
        if (ok && M.res >= 0) {
          const TfTensor& rt = m.tensors[M.res];
          if (M.pool) ok = rt.dim(1) == 2 * ot.dim(1) && rt.dim(2) == 2 * ot.dim(2);
        }
        if (ok) {
          layers.push_back({(int)j, M, M.has_dw ? 0 : 2, M.cur});
          chainT.push_back(M.cur);
          continue;
        }
      }
      // whole-map convolution with one filter: a dot product over the resident map
      if (in_chain(op.in[0]) && op.in.size() >= 2) {
        const TfTensor& wt = m.tensors[op.in[1]];
        const TfTensor& it = m.tensors[op.in[0]];
        if (wt.shape.size() == 4 && wt.shape[0] == 1 && wt.shape[1] == it.dim(1) && wt.shape[2] == it.dim(2) && op.padding == 1 &&
            op.act == 0 && op.dil_h == 1 && op.dil_w == 1 && numel(op.out[0]) == 1) {
          PwMatch M;
          M.src = op.in[0]; M.cur = op.out[0];
          layers.push_back({(int)j, M, 3, op.out[0]});
        }
      }
    }
    if (layers.size() > (size_t)kTailMaxLayers) {
      if (!wide) return;
      layers.resize(kTailMaxLayers);
    }
    if (!wide) {
      if (layers.size() >= 3 && build_chain(T0, layers, false)) *narrow_done = true;
      return;
    }
    // wide: the longest prefix (in graph order) whose buffers and weight ring fit one CTA
    for (size_t n = layers.size(); n >= 1; --n) {
      std::vector<ChainLayer> pre(layers.begin(), layers.begin() + n);
      // every layer of the prefix must read tensors of the prefix
      std::vector<int> have{T0};
      bool closed = true;
      for (const ChainLayer& cl : pre) {
        closed = closed && std::find(have.begin(), have.end(), cl.M.src) != have.end();
        if (cl.M.res >= 0 && std::find(have.begin(), have.end(), cl.M.res) == have.end()) closed = closed && hbm_res_ok(cl.M.res) && !in_chain(cl.M.res);
        have.push_back(cl.out_tf);
      }
      if (closed && build_chain(T0, pre, true)) return;
    }
  }

  bool build_chain(int T0, const std::vector<ChainLayer>& layers, bool wide) {
    const size_t blob_mark = P.blob.size();
    if (build_chain_inner(T0, layers, wide)) return true;
    P.blob.resize(blob_mark);                         // a failed attempt leaves no records behind
    return false;
  }

  bool build_chain_inner(int T0, const std::vector<ChainLayer>& layers, bool wide) {
    std::vector<char> absorbed(m.ops.size(), 0);
    std::vector<int> chainT{T0};
    for (const ChainLayer& cl : layers) {
      absorbed[cl.conv] = 1;
      for (int a : cl.M.absorbed) absorbed[a] = 1;
      if (cl.kind != 3) chainT.push_back(cl.out_tf);
    }
    auto in_chain = [&](int t) { return std::find(chainT.begin(), chainT.end(), t) != chainT.end(); };
    // pass A2: chain tensors read outside the chain (or graph outputs) are exits
    std::vector<int> exits;
    for (size_t j = 0; j < m.ops.size(); ++j) {
      if (absorbed[j]) continue;
      for (int i : m.ops[j].in) {
        if (i < 0 || i == T0 || !in_chain(i)) continue;
        const TfOp& op = m.ops[j];
        if (op.code == kOpConv2D && !wide) {               // a 1x1 head on a chain tensor: that tail belongs to fuse_tail()
          const TfTensor& wt = m.tensors[op.in[1]];
          if (wt.shape.size() == 4 && wt.shape[1] == 1 && wt.shape[2] == 1) return false;
        }
        if (std::find(exits.begin(), exits.end(), i) == exits.end()) exits.push_back(i);
      }
    }
    for (int o : m.outputs) if (in_chain(o) && o != T0) return false;
    // pass B: buffers by liveness
    const int NL = (int)layers.size();
    std::map<int, int> last_read;                     // tensor -> last layer that reads it
    for (int i = 0; i < NL; ++i) {
      last_read[layers[i].M.src] = i;
      if (layers[i].M.res >= 0) last_read[layers[i].M.res] = i;
    }
    std::map<int, int> buf_of;
    std::vector<int> buf_px, buf_c, buf_free_after;   // a buffer is free after layer buf_free_after (its tenant's last read)
    buf_of[T0] = 0;
    buf_px.push_back(m.tensors[T0].dim(1) * m.tensors[T0].dim(2));
    buf_c.push_back(wide ? P.tensors[pt(T0)].Cs : ru(m.tensors[T0].dim(3), 4));     // (the wide loader copies whole Cs-float records)
    buf_free_after.push_back(last_read[T0]);
    PStep t;
    t.kind = kStepTailWs;
    t.tail_wide = wide ? 1 : 0;
    int bias_s = 0;
    for (int i = 0; i < NL; ++i) {
      const ChainLayer& cl = layers[i];
      const TfOp& conv = m.ops[cl.conv];
      const TfTensor& it = m.tensors[cl.M.src];
      const TfTensor& ot = m.tensors[cl.out_tf];
      TailLayerD L = {};
      L.kind = cl.kind;
      L.src = buf_of[cl.M.src];
      const bool res_hbm = cl.M.res >= 0 && !in_chain(cl.M.res);
      L.rbuf = cl.M.res >= 0 && !res_hbm ? buf_of[cl.M.res] : L.src;
      L.o1 = L.o2 = -1;
      L.wscale = 1.f; L.w_parts = 1; L.dst = -1;
      L.nk = 1; L.ng = 1 | (1 << 8);
      L.IH = it.dim(1); L.IW = it.dim(2);
      L.Cin = it.dim(3);
      std::vector<float> w, b;
      if (!m.const_f32(conv.in[1], &w)) return false;
      if (conv.in.size() > 2 && conv.in[2] >= 0 && !m.const_f32(conv.in[2], &b)) return false;
      if (cl.kind == 3) {
        L.OH = L.IH; L.OW = L.IW; L.Cout = 1; L.K16 = 16; L.Npad = 4;
        b.resize(4, 0.f);
        L.bias_off = (int)push(b, 4);
        L.bias_s = wide ? bias_s : i * 128;
        bias_s += 4;
        const size_t n = ru((int)w.size(), 4);
        L.tap_bytes = (int)(n * 4);
        L.tap_off = (int)push(w, n);
        L.o1 = (int)t.tail_outs.size(); t.tail_outs.push_back(pt(cl.out_tf));
        t.macs += (double)w.size();
        t.tail.push_back(L);
        continue;
      }
      L.OH = ot.dim(1); L.OW = ot.dim(2);
      L.Cout = m.tensors[conv.in[1]].shape[0];
      L.K16 = ru(L.Cin, 16); L.Npad = ru(L.Cout, 16);
      L.stride = cl.M.dws; L.pad = cl.M.dpt;
      L.res = cl.M.res >= 0 ? (cl.M.pool ? 2 : 1) : 0;
      if (cl.M.res >= 0) L.c1 = ru(m.tensors[cl.M.res].dim(3), 4);          // kinds 0 / 2: channels of the residual tensor (k_chain_wide)
      if (res_hbm) {
        const int rp = pt(cl.M.res);
        if (P.tensors[rp].Cs % 4 != 0 || P.tensors[rp].root >= 0) return false;
        size_t k = std::find(t.tail_rsrc.begin(), t.tail_rsrc.end(), rp) - t.tail_rsrc.begin();
        if (k == t.tail_rsrc.size()) t.tail_rsrc.push_back(rp);
        if (t.tail_rsrc.size() > 2) return false;
        L.res += 2;                                   // 3: same pixel, 4: 2x2 max-pool, of the HBM tensor rsrc[rbuf]
        L.rbuf = (int)k;
      }
      L.act = cl.M.act;
      const int npix = L.OH * L.OW;
      const bool paired = npix > 128;
      const int ngroups = paired ? (L.Npad + 127) / 128 : 1;
      if (!wide && (L.K16 > 128 || L.Npad > 128)) return false;
      // destination: none when nothing in the chain reads the result (it only leaves to HBM); in place on the residual or the
      // source when that dies here and keeps its pixel count (never the source when the layer runs as several column groups:
      // the later groups still read it; never buffer 0 of a wide chain: its loader rewrites only the input's channels), else a
      // free buffer of that size, else a new one
      const int sb = L.src;
      const bool is_exit = std::find(exits.begin(), exits.end(), cl.out_tf) != exits.end();
      const bool read_inside = last_read.count(cl.out_tf) > 0;
      if (!read_inside && !is_exit) return false;
      if (read_inside || !wide) {
        const int rb = (cl.M.res >= 0 && !res_hbm && L.res == 1) ? L.rbuf : -1;
        if (wide && rb > 0 && buf_px[rb] == npix && last_read[cl.M.res] == i) {
          L.dst = rb;
        } else if (buf_px[sb] == npix && last_read[cl.M.src] == i && !(wide && sb == 0) && ngroups == 1) {
          L.dst = sb;
        } else {
          for (size_t k = 1; k < buf_px.size() && L.dst < 0; ++k)
            if (buf_px[k] == npix && buf_free_after[k] < i && (int)k != L.rbuf && (int)k != sb) L.dst = (int)k;
          if (L.dst < 0) { L.dst = (int)buf_px.size(); buf_px.push_back(npix); buf_c.push_back(4); buf_free_after.push_back(-1); }
        }
        if (L.dst >= kTailMaxBufs) return false;
        buf_of[cl.out_tf] = L.dst;
        buf_c[L.dst] = std::max(buf_c[L.dst], ru(L.Cout, 4));
        buf_free_after[L.dst] = last_read.count(cl.out_tf) ? last_read[cl.out_tf] : i;
      }
      if (is_exit) {
        L.o1 = (int)t.tail_outs.size(); t.tail_outs.push_back(pt(cl.out_tf));
      }
      if (t.tail_outs.size() > 4) return false;
      // weights
      b.resize(L.Cout, 0.f);
      L.w_parts = f16_exact(w) ? 1 : 2;
      if (!wide) {
        std::vector<float> rec = pack_w_f16(w, L.Cout, L.Cin, L.Npad, L.K16, L.w_parts, &L.wscale);
        L.rec_bytes = (int)(rec.size() * 4);
        L.rec_off = (int)push(rec, rec.size());
        L.bias_s = i * 128;
      } else {
        float scale = 1.f;
        if (L.w_parts == 2) {
          float mx = 0.f;
          for (float v : w) mx = std::max(mx, std::fabs(v));
          if (mx > 0.f) { int e; std::frexp(mx, &e); scale = std::ldexp(1.f, 14 - e); }
        }
        L.wscale = 1.f / scale;
        L.nk = (L.K16 + 127) / 128;
        const int nblocks_n = (L.Npad + 127) / 128;               // column blocks of <= 128, balanced, multiples of 16
        const int bw = ru((L.Npad / 16 + nblocks_n - 1) / nblocks_n * 16, 16);
        const int nbg = paired ? 1 : nblocks_n;
        if (!paired && L.Npad > 384) return false;
        L.ng = (paired ? nblocks_n : 1) | (nbg << 8);
        L.blk0 = (int)t.tail_blks.size();
        for (int grp = 0; grp < (paired ? nblocks_n : 1); ++grp)
          for (int kc = 0; kc < L.nk; ++kc)
            for (int nb = 0; nb < nbg; ++nb) {
              const int bi = paired ? grp : nb;
              const int n0 = bi * bw, ncols = std::min(bw, L.Npad - n0);
              const int k0 = kc * 128, kw = std::min(128, L.K16 - k0);
              std::vector<float> rec = pack_w_block(w, L.Cout, L.Cin, n0, ncols, k0, kw, L.w_parts, scale);
              TailBlk bk;
              bk.bytes = (int)(rec.size() * 4);
              bk.off = (int)push(rec, rec.size());
              bk.n0 = n0; bk.ncols = ncols;
              if (bk.bytes % 16 || ncols < 16 || ncols % 16) return false;
              t.tail_blks.push_back(bk);
            }
        L.bias_s = bias_s;
        bias_s += L.Npad;
      }
      std::vector<float> bias((size_t)L.Npad, 0.f);
      for (int n = 0; n < L.Cout; ++n) bias[n] = b[n];
      L.bias_off = (int)push(bias, bias.size());
      if (cl.M.alpha_tf >= 0) {
        std::vector<float> al;
        if (!m.const_f32(cl.M.alpha_tf, &al) || (int)al.size() != L.Cout) return false;
        L.alpha_off = (int)push(al, (size_t)L.Npad);
      }
      if (cl.kind == 0) {
        const TfOp& dw = m.ops[cl.M.absorbed[0]];
        std::vector<float> dwv, dbv;
        if (!m.const_f32(dw.in[1], &dwv) || !m.const_f32(dw.in[2], &dbv)) return false;
        std::vector<float> rec2((size_t)10 * L.K16, 0.f);
        for (int k = 0; k < 9; ++k)
          for (int c = 0; c < L.Cin; ++c) rec2[(size_t)k * L.K16 + c] = dwv[(size_t)k * L.Cin + c];
        for (int c = 0; c < L.Cin; ++c) rec2[(size_t)9 * L.K16 + c] = dbv[c];
        L.tap_bytes = (int)(rec2.size() * 4);
        L.tap_off = (int)push(rec2, rec2.size());
      }
      t.macs += (double)npix * ((double)L.Cin * L.Cout + (cl.kind == 0 ? 9.0 * L.Cin : 0.0));
      t.tail.push_back(L);
    }
    t.tail_bias_floats = wide ? ru(bias_s, 4) + 4 : kTailMaxLayers * 128;
    if (t.tail_outs.empty() || !finish_tail(&t, buf_px, buf_c)) return false;
    t.in = pt(T0);
    t.out = t.tail_outs[0];
    for (size_t k = 1; k < t.tail_outs.size(); ++k) t.extra_out.push_back(t.tail_outs[k]);
    t.name = "chain:" + m.tensors[layers.front().out_tf].name + ".." + m.tensors[layers.back().out_tf].name;
    // commit
    int first = (int)m.ops.size();
    for (size_t j = 0; j < m.ops.size(); ++j)
      if (absorbed[j]) { done[j] = 1; first = std::min(first, (int)j); }
    for (int e : exits) fused_outputs.push_back(e);
    chains.push_back({first, t});
    return true;
  }

  // ---- emission ---------------------------------------------------------------------------------
  void use(int step, int t) {
    if (t < 0) return;
    PTensor& x = P.tensors[t];
    if (x.def_step < 0) x.def_step = step;
    x.last_use = std::max(x.last_use, step);
  }

  bool emit(PStep s) {
    int idx = (int)P.steps.size();
    if (s.in >= 0 && !P.tensors[s.in].materialized && !(s.in_u8))
      return fail("internal: step '" + s.name + "' reads a tensor that is not materialised");
    if (s.in2 >= 0 && !P.tensors[s.in2].materialized)
      return fail("internal: step '" + s.name + "' reads a residual that is not materialised");
    P.tensors[s.out].materialized = true;
    if (s.out2 >= 0) P.tensors[s.out2].materialized = true;
    if (!s.in_u8) use(idx, s.in);
    use(idx, s.in2);
    use(idx, s.out);
    use(idx, s.out2);
    P.steps.push_back(std::move(s));
    return true;
  }

  bool run() {
    size_t nt = m.tensors.size(), no = m.ops.size();
    done.assign(no, 0);
    ncons.assign(nt, 0);
    views.assign(nt, View());
    for (const TfOp& op : m.ops)
      for (int i : op.in) if (i >= 0) ncons[i]++;
    for (int o : m.outputs) ncons[o]++;
    P.out_elems.clear();
    for (int o : m.outputs) P.out_elems.push_back(numel(o));
    for (size_t oi = 0; oi < m.outputs.size(); ++oi)
      if (!assign_view(m.outputs[oi], (int)oi, 0)) return false;

    const TfTensor& in = m.tensors[m.inputs[0]];
    if (in.shape.size() != 4 || in.shape[3] != 3) return fail("model input must be [1,H,W,3]");
    P.in_h = in.shape[1];
    P.in_w = in.shape[2];

    // pass 1: fusion analysis
    if (fuse == 1 && use_tc) fuse_chain();
    if (fuse >= 1) {
      for (size_t i = 0; i < no; ++i) {
        if (done[i] || m.ops[i].code != kOpConv2D) continue;
        Fused F;
        if (!analyse_pointwise((int)i, &F) && err.empty()) analyse_dense((int)i, &F);
        if (!err.empty()) return false;
        if (F.ok) fused[(int)i] = F;
      }
    }
    // graph input
    P.input = pt(m.inputs[0]);
    P.tensors[P.input].is_input = true;
    bool need_f32_input = fuse == 0;
    for (int c : m.consumers(m.inputs[0])) {
      auto it = fused.find(c);
      if (it == fused.end() || !it->second.st.in_u8) need_f32_input = true;
    }
    if (need_f32_input) {
      PStep s;
      s.kind = kStepNormalize;
      s.name = "normalize";
      s.out = P.input;
      s.in_u8 = true;
      if (!emit(s)) return false;
    }
    // pass 2: emission in graph order
    for (size_t i = 0; i < no; ++i) {
      const TfOp& op = m.ops[i];
      for (const ChainOut& co : chains) {
        if ((int)i != co.emit_at) continue;
        PStep s = co.step;
        if (!P.tensors[s.in].materialized) return fail("internal: the chain's input is not materialised");
        int idx = (int)P.steps.size();
        for (int e : s.extra_out) { P.tensors[e].materialized = true; use(idx, e); }
        for (int e : s.tail_rsrc) {
          if (!P.tensors[e].materialized) return fail("internal: a chain residual is not materialised");
          use(idx, e);
        }
        if (!emit(s)) return false;
      }
      auto fit = fused.find((int)i);
      if (fit != fused.end()) {
        Fused& F = fit->second;
        PStep s = F.st;
        s.in = pt(F.src_tf);
        s.in2 = F.res_tf >= 0 ? pt(F.res_tf) : -1;
        s.out = pt(F.out_tf);
        s.out2 = F.out2_tf >= 0 ? pt(F.out2_tf) : -1;
        if (s.in_u8) s.in = P.input;
        if (!emit(s)) return false;
        continue;
      }
      if (done[i] || op.code == kOpDequantize) continue;
      PStep s;
      s.name = op.out.empty() ? "" : m.tensors[op.out[0]].name;
      switch (op.code) {
        case kOpConv2D: case kOpDwConv2D: {
          const TfTensor& wt = m.tensors[op.in[1]];
          const TfTensor& it = m.tensors[op.in[0]];
          if (wt.shape.size() != 4 || it.shape.size() != 4) return fail("conv rank");
          if (op.dil_h != 1 || op.dil_w != 1 || (op.code == kOpDwConv2D && op.depth_mult != 1))
            return fail("dilated / depth-multiplied convolutions are not supported");
          if (op.act != 0 && op.act != 1) return fail("unsupported fused activation");
          s.kind = kStepNaiveConv;
          s.depthwise = op.code == kOpDwConv2D;
          s.kh = wt.shape[1]; s.kw = wt.shape[2]; s.sh = op.stride_h; s.sw = op.stride_w;
          if (op.padding == 0) { same_pad(it.dim(1), s.kh, s.sh, &s.pt); same_pad(it.dim(2), s.kw, s.sw, &s.pl); }
          s.act = op.act == 1 ? kActRelu : kActNone;
          std::vector<float> w, b;
          if (!m.const_f32(op.in[1], &w)) return fail("conv weights are not constant");
          if (op.in.size() > 2 && op.in[2] >= 0 && !m.const_f32(op.in[2], &b)) return fail("conv bias is not constant");
          s.Cout = s.depthwise ? wt.shape[3] : wt.shape[0];
          b.resize(s.Cout, 0.f);
          s.w = push(w, w.size());
          s.bias = push(b, b.size());
          s.in = pt(op.in[0]);
          s.out = pt(op.out[0]);
          const TfTensor& ot = m.tensors[op.out[0]];
          s.macs = (double)ot.dim(1) * ot.dim(2) * s.kh * s.kw * (s.depthwise ? s.Cout : (double)wt.shape[3] * s.Cout);
          break;
        }
        case kOpAdd:
          if (op.in.size() != 2 || numel(op.in[0]) != numel(op.in[1])) return fail("broadcasting ADD is not supported");
          if (op.act != 0 && op.act != 1) return fail("unsupported fused activation");
          s.kind = kStepAdd;
          s.act = op.act == 1 ? kActRelu : kActNone;
          s.in = pt(op.in[0]); s.in2 = pt(op.in[1]); s.out = pt(op.out[0]);
          break;
        case kOpRelu:
          s.kind = kStepAct; s.act = kActRelu; s.in = pt(op.in[0]); s.out = pt(op.out[0]);
          break;
        case kOpPrelu: {
          s.kind = kStepAct; s.act = kActPrelu; s.in = pt(op.in[0]); s.out = pt(op.out[0]);
          std::vector<float> a;
          if (!m.const_f32(op.in[1], &a) || (int)a.size() != P.tensors[s.out].C) return fail("PRELU alpha must be per-channel");
          s.alpha = push(a, a.size() + 8);
          break;
        }
        case kOpPad: {
          std::vector<int> pads;
          if (!m.const_i32(op.in[1], &pads) || pads.size() != 8) return fail("PAD paddings");
          for (int k = 0; k < 7; ++k) if (pads[k] != 0) return fail("only trailing channel PAD is supported");
          s.kind = kStepPadC; s.in = pt(op.in[0]); s.out = pt(op.out[0]);
          break;
        }
        case kOpMaxPool: {
          const TfTensor& it = m.tensors[op.in[0]];
          s.kind = kStepMaxPool; s.fh = op.filter_h; s.fw = op.filter_w; s.sh = op.stride_h; s.sw = op.stride_w;
          if (op.padding == 0) { same_pad(it.dim(1), s.fh, s.sh, &s.pt); same_pad(it.dim(2), s.fw, s.sw, &s.pl); }
          s.in = pt(op.in[0]); s.out = pt(op.out[0]);
          break;
        }
        case kOpResizeBilinear: {
          s.kind = kStepResize; s.align = op.align_corners; s.half = op.half_pixel;
          s.in = pt(op.in[0]); s.out = pt(op.out[0]);
          // RESIZE_BILINEAR -> ADD(skip, resized) [ReLU] (the FPN up-path): one kernel, the upsampled tensor never reaches HBM
          const int c = fuse >= 1 && !is_view(op.out[0]) ? sole_consumer(op.out[0]) : -1;
          if (c >= 0 && !done[c] && m.ops[c].code == kOpAdd && m.ops[c].in.size() == 2 && (m.ops[c].act == 0 || m.ops[c].act == 1)) {
            const TfOp& add = m.ops[c];
            const int other = add.in[0] == op.out[0] ? add.in[1] : add.in[0];
            auto it = P.tf2pt.find(other);
            const bool vec = P.tensors[s.out].Cs % 4 == 0 && it != P.tf2pt.end() && P.tensors[it->second].materialized &&
                             P.tensors[it->second].Cs == P.tensors[s.out].Cs && other != op.out[0] && numel(other) == numel(op.out[0]) &&
                             !is_view(other) && !is_view(add.out[0]);
            if (vec) {
              s.in2 = it->second;
              s.out = pt(add.out[0]);
              s.act = add.act == 1 ? kActRelu : kActNone;
              s.name = m.tensors[add.out[0]].name;
              done[c] = 1;
            }
          }
          break;
        }
        default: {
          char buf[96];
          snprintf(buf, sizeof buf, "unsupported TFLite op %d in graph", op.code);
          return fail(buf);
        }
      }
      if (!emit(s)) return false;
    }
    // outputs
    for (int o : m.outputs) {
      int t = pt(o);
      P.outputs.push_back(t);
    }
    static const int want_tail = [] { const char* e = std::getenv("FDT_TAIL"); return e ? std::atoi(e) : 1; }();   // FDT_TAIL=0: A/B, one launch per layer
    if (fuse == 1 && use_tc && want_tail) fuse_tail();
    // arena allocation
    struct Block { long long off, size; int free_after; };
    std::vector<Block> blocks;
    long long top = 0;
    std::vector<int> order;
    for (size_t t = 0; t < P.tensors.size(); ++t)
      if (P.tensors[t].materialized && P.tensors[t].root < 0) order.push_back((int)t);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return P.tensors[a].def_step < P.tensors[b].def_step; });
    for (int t : order) {
      PTensor& x = P.tensors[t];
      long long need = (x.istride + 7) / 8 * 8;      // 32-byte granules: the epilogues may use 256-bit stores
      bool placed = false;
      if (fuse == 1) {
        for (Block& b : blocks) {
          if (b.free_after < x.def_step && b.size >= need) {
            x.arena_off = b.off;
            b.free_after = x.last_use;
            placed = true;
            break;
          }
        }
      }
      if (!placed) {
        x.arena_off = top;
        blocks.push_back({top, need, x.last_use});
        top += need;
      }
    }
    P.arena_per_image = top;
    return true;
  }
};

}  // namespace

bool Plan::build(const TfModel& m, int fuse, std::string* err, bool use_tc) {
  *this = Plan();
  fuse_level = fuse;
  Builder b(m, *this, fuse);
  b.use_tc = use_tc;
  bool ok = b.run();
  if (!ok && err) *err = b.err;
  return ok;
}

std::string Plan::describe() const {
  static const char* kn[] = {"normalize", "naive_conv", "gemm_conv", "dwpw", "add", "act", "padc", "maxpool", "resize", "stem", "block_ws", "stem_ws", "tail_ws", "fc_tc", "block_ts"};
  std::string s;
  char buf[512];
  double macs = 0;
  for (size_t i = 0; i < steps.size(); ++i) {
    const PStep& st = steps[i];
    const PTensor& o = tensors[st.out];
    macs += st.macs;
    snprintf(buf, sizeof buf,
             "%3zu %-10s in=%d res=%d%s/m%d out=%d[%dx%dx%d Cs%d %s] act=%d dw=%d/s%d K=%d NC=%dx%d TM=%d NPG=%d tile=%dx%dx%d RS=%d nd=%d ns=%d na=%d no=%d smem=%zu  %s\n",
             i, kn[st.kind], st.in >= 0 ? tensors[st.in].tf : -1, st.in2 >= 0 ? tensors[st.in2].tf : -1,
             st.res_pool ? "(pool)" : "", st.res_mode, o.tf, o.H, o.W, o.C, o.Cs, o.root >= 0 ? "view" : "arena", st.act,
             (int)st.has_dw, st.dws, st.K, st.NC, st.nchunks, st.TM, st.NPG, st.TH, st.TW, st.G, st.RS, st.nd, st.ns, st.na, st.no, st.smem, st.name.c_str());
    s += buf;
    if (st.kind == kStepBlockWs && (st.nr > 0 || st.deint)) {
      snprintf(buf, sizeof buf, "      k_block_ws: residual ring %d x %d B (record %d floats)%s\n", st.nr, st.res_stage_floats * 4, st.KSr,
               st.deint ? ", stride 2 on even / odd column planes" : "");
      s += buf;
    }
    if (st.tail_wide) {
      snprintf(buf, sizeof buf, "      k_chain_wide: %zu W blocks (ring %d x %d B), act %d floats, %zu HBM residual tensor(s)\n", st.tail_blks.size(), st.tail_wdepth, st.tail_wbuf,
               st.tail_act_floats, st.tail_rsrc.size());
      s += buf;
    }
    for (size_t l = 0; l < st.tail.size(); ++l) {
      const TailLayerD& L = st.tail[l];
      static const char* lk[] = {"block", "heads", "pw", "dot"};
      snprintf(buf, sizeof buf, "      tail %2zu %s buf %d->%d %dx%d->%dx%d s%d C %d->%d K16=%d Npad=%d res=%d(buf %d) act=%d wparts=%d out=%d rec=%dB taps=%dB\n", l, lk[L.kind & 3],
               L.src, L.dst, L.IH, L.IW, L.OH, L.OW, L.stride, L.Cin, L.Cout, L.K16, L.Npad, L.res, L.rbuf, L.act, L.w_parts, L.o1, L.rec_bytes, L.tap_bytes);
      s += buf;
      if (st.tail_wide && L.kind != 3) {
        s.pop_back();
        snprintf(buf, sizeof buf, " wide: groups=%d blocks/group=%d Kchunks=%d blk0=%d\n", L.ng & 0xff, L.ng >> 8, L.nk, L.blk0);
        s += buf;
      }
    }
  }
  snprintf(buf, sizeof buf, "steps=%zu  MACs/image=%.3fM  arena/image=%.1f KB  weights=%.1f KB\n", steps.size(),
           macs / 1e6, arena_per_image * 4.0 / 1024, blob.size() * 4.0 / 1024);
  s += buf;
  return s;
}

}  // namespace fdt
