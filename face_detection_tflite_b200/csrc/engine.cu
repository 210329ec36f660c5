#include "engine.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace fdt {

Engine::~Engine() {
  if (d_blob_) cudaFree(d_blob_);
  for (TailLayerD* p : d_tail_) if (p) cudaFree(p);
  for (TailBlk* p : d_blks_) if (p) cudaFree(p);
}

bool Engine::init(const uint8_t* tflite, size_t len, int fuse_level, std::string* err, bool use_tc) {
  if (!model_.parse(tflite, len, err)) return false;
  if (!plan_.build(model_, fuse_level, err, use_tc)) return false;
  size_t bytes = plan_.blob.size() * sizeof(float) + 64;
  if (cudaMalloc(&d_blob_, bytes) != cudaSuccess) { *err = "cudaMalloc(weights) failed"; return false; }
  cudaMemset(d_blob_, 0, bytes);
  if (cudaMemcpy(d_blob_, plan_.blob.data(), plan_.blob.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess) {
    *err = "weight upload failed";
    return false;
  }
  d_tail_.assign(plan_.steps.size(), nullptr);
  d_blks_.assign(plan_.steps.size(), nullptr);
  for (size_t i = 0; i < plan_.steps.size(); ++i) {
    const PStep& st = plan_.steps[i];
    if (st.kind != kStepTailWs) continue;
    if (cudaMalloc(&d_tail_[i], st.tail.size() * sizeof(TailLayerD)) != cudaSuccess ||
        cudaMemcpy(d_tail_[i], st.tail.data(), st.tail.size() * sizeof(TailLayerD), cudaMemcpyHostToDevice) != cudaSuccess) {
      *err = "tail program upload failed";
      return false;
    }
    if (st.tail_blks.empty()) continue;
    if (cudaMalloc(&d_blks_[i], st.tail_blks.size() * sizeof(TailBlk)) != cudaSuccess ||
        cudaMemcpy(d_blks_[i], st.tail_blks.data(), st.tail_blks.size() * sizeof(TailBlk), cudaMemcpyHostToDevice) != cudaSuccess) {
      *err = "chain block table upload failed";
      return false;
    }
  }
  return true;
}

bool Engine::make_ctx(int cap, EngineCtx* ctx, std::string* err) const {
  ctx->cap = cap;
  size_t bytes = (size_t)cap * plan_.arena_per_image * sizeof(float) + 256;
  if (cudaMalloc(&ctx->arena, bytes) != cudaSuccess) { *err = "cudaMalloc(arena) failed"; return false; }
  cudaMemset(ctx->arena, 0, bytes);
  ctx->outputs.clear();
  for (long long e : plan_.out_elems) {
    float* p = nullptr;
    size_t ob = (size_t)cap * e * sizeof(float) + 256;
    if (cudaMalloc(&p, ob) != cudaSuccess) { *err = "cudaMalloc(outputs) failed"; return false; }
    cudaMemset(p, 0, ob);
    ctx->outputs.push_back(p);
  }
  return true;
}

void Engine::free_ctx(EngineCtx* ctx) const {
  if (ctx->arena) cudaFree(ctx->arena);
  for (float* p : ctx->outputs) cudaFree(p);
  ctx->arena = nullptr;
  ctx->outputs.clear();
}

TV Engine::view(const EngineCtx& ctx, int t) const {
  const PTensor& x = plan_.tensors[t];
  TV v;
  v.H = x.H; v.W = x.W; v.C = x.C; v.Cs = x.Cs;
  v.istride = x.istride;
  if (x.root >= 0) v.p = ctx.outputs[x.root] + x.view_off;
  else v.p = ctx.arena + (size_t)x.arena_off * ctx.cap;
  return v;
}

static int cta_cap(size_t smem, int threads) {
  int by_smem = (int)((227u * 1024u) / (smem + 1024u));
  int by_thr = 2048 / std::max(threads, 32);
  int per_sm = std::max(1, std::min(std::min(by_smem, by_thr), 16));
  return 148 * per_sm;
}

int Engine::run(const EngineCtx& ctx, const uint8_t* in_u8, int B, cudaStream_t s, cudaEvent_t* evs) const {
  int launches = 0;
  const float* blob = d_blob_;
  for (const PStep& st : plan_.steps) {
    if (evs) cudaEventRecord(evs[launches], s);
    TV out = view(ctx, st.out);
    switch (st.kind) {
      case kStepNormalize:
        launch_normalize(in_u8, out, B, s);
        break;
      case kStepNaiveConv: {
        NaiveConvP p;
        p.in = view(ctx, st.in); p.out = out;
        p.w = blob + st.w; p.bias = blob + st.bias;
        p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.kh = st.kh; p.kw = st.kw; p.sh = st.sh; p.sw = st.sw; p.pt = st.pt; p.pl = st.pl;
        p.act = st.act; p.depthwise = st.depthwise;
        launch_naive_conv(p, B, s);
        break;
      }
      case kStepStem: {
        StemP p;
        const PTensor& it = plan_.tensors[st.in];
        p.in8 = in_u8; p.H = it.H; p.W = it.W; p.OH = out.H; p.OW = out.W;
        p.kw = st.kw; p.pt = st.pt; p.pl = st.pl; p.K = st.K; p.KP = st.KP;
        p.out = out.p; p.out_istride = out.istride; p.Cout = st.Cout; p.CoutS = out.Cs;
        p.vec_store = (out.Cs % 4 == 0 && out.istride % 4 == 0 && ((size_t)out.p % 16 == 0)) ? ((out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 2 : 1) : 0;   // 2: 32-byte stores allowed
        p.w = blob + st.w; p.bias = blob + st.bias; p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act; p.CoutP = st.CoutP; p.NC = st.NC; p.smem_bytes = st.smem;
        launch_stem(p, B, s, cta_cap(st.smem, 32 * (st.NC / 4)));
        break;
      }
      case kStepStemWs: {
        StemWsP p;
        const PTensor& it = plan_.tensors[st.in];
        p.in8 = in_u8; p.H = it.H; p.W = it.W; p.OH = out.H; p.OW = out.W;
        p.kw = st.kw; p.pt = st.pt; p.pl = st.pl;
        p.out = out.p; p.out_istride = out.istride; p.Cout = st.Cout; p.CoutS = out.Cs;
        p.vec_store = (out.Cs % 4 == 0 && out.istride % 4 == 0 && ((size_t)out.p % 16 == 0)) ? ((out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 2 : 1) : 0;   // 2: 32-byte stores allowed
        p.wB = blob + st.w; p.bias = blob + st.bias; p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act; p.Npad = st.Npad; p.K8 = st.K8; p.tmem_cols = st.tmem_cols; p.w_parts = st.w_parts;
        p.out_scale = (float)st.out_scale; p.smem_bytes = st.smem;
        if (!launch_stem_ws(p, B, ctx.cap, s)) { failed_ = true; fprintf(stderr, "fdt: cuTensorMapEncodeTiled failed for step '%s'\n", st.name.c_str()); }
        break;
      }
      case kStepGemmConv: {
        GemmConvP p;
        const PTensor& it = plan_.tensors[st.in];
        if (st.in_u8) {
          p.in = nullptr; p.in8 = in_u8; p.in_istride = (long long)it.H * it.W * 4;
          p.Cin = 3; p.CinS = 3;
        } else {
          TV iv = view(ctx, st.in);
          p.in = iv.p; p.in8 = nullptr; p.in_istride = iv.istride; p.Cin = it.C; p.CinS = it.Cs;
        }
        p.H = it.H; p.W = it.W;
        p.kh = st.kh; p.kw = st.kw; p.sh = st.sh; p.sw = st.sw; p.pt = st.pt; p.pl = st.pl;
        p.OH = out.H; p.OW = out.W;
        p.K = st.K; p.KP = st.KP; p.KS = st.KS;
        p.out = out.p; p.out_istride = out.istride; p.Cout = st.Cout; p.CoutS = out.Cs;
        p.vec_store = (out.Cs % 4 == 0 && out.istride % 4 == 0 && ((size_t)out.p % 16 == 0)) ? ((out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 2 : 1) : 0;   // 2: 32-byte stores allowed
        p.w = blob + st.w; p.bias = blob + st.bias; p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act; p.CoutP = st.CoutP; p.NC = st.NC; p.nchunks = st.nchunks;
        p.NPG = st.NPG; p.TM = st.TM; p.smem_bytes = st.smem;
        p.flat = (!st.in_u8 && out.H == 1 && out.W == 1 && st.pt == 0 && st.pl == 0 && st.kh == it.H && st.kw == it.W &&
                  it.Cs == it.C && (st.K % 4) == 0 && it.istride % 4 == 0) ? 1 : 0;
        p.fd_KP = FastDiv(st.KP); p.fd_kwc = FastDiv(st.kw * p.Cin); p.fd_Cin = FastDiv(p.Cin);
        p.fd_OW = FastDiv(out.W); p.fd_OH = FastDiv(out.H); p.fd_NQ = FastDiv(st.NC / 4);
        launch_gemm_conv(p, B, s, cta_cap(st.smem, st.NPG * (st.NC / 4)));
        break;
      }
      case kStepDwPw: {
        DwPwP p;
        TV iv = view(ctx, st.in);
        p.in = iv.p; p.in_istride = iv.istride; p.H = iv.H; p.W = iv.W; p.Cin = iv.C; p.CinS = iv.Cs;
        p.has_dw = st.has_dw ? 1 : 0; p.s = st.dws; p.dpt = st.dpt; p.dpl = st.dpl;
        p.OH = out.H; p.OW = out.W;
        p.dww = st.dww >= 0 ? blob + st.dww : nullptr; p.dwb = st.dwb >= 0 ? blob + st.dwb : nullptr;
        p.KP = st.KP; p.KS = st.KS;
        p.out = out.p; p.out_istride = out.istride; p.Cout = st.Cout; p.CoutS = out.Cs;
        p.vec_store = (out.Cs % 4 == 0 && out.istride % 4 == 0 && ((size_t)out.p % 16 == 0)) ? ((out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 2 : 1) : 0;   // 2: 32-byte stores allowed
        p.w = blob + st.w; p.bias = blob + st.bias; p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act; p.CoutP = st.CoutP; p.NC = st.NC; p.nchunks = st.nchunks;
        p.NPG = st.NPG; p.TM = st.TM;
        p.res_mode = st.in2 >= 0 ? st.res_mode : 0;
        if (st.in2 >= 0) {
          TV rv = view(ctx, st.in2);
          p.res = rv.p; p.res_istride = rv.istride; p.res_H = rv.H; p.res_W = rv.W; p.res_C = rv.C; p.res_Cs = rv.Cs;
        } else {
          p.res = nullptr; p.res_istride = 0; p.res_H = p.res_W = p.res_C = p.res_Cs = 0;
        }
        p.res_pool = st.res_pool;
        p.TH = st.TH; p.TW = st.TW; p.G = st.G; p.IH = st.IH; p.IW = st.IW; p.tilesX = st.tilesX; p.tilesY = st.tilesY;
        p.fd_Q = FastDiv(st.KP / 4); p.fd_IW = FastDiv(st.IW); p.fd_IH = FastDiv(st.IH); p.fd_TW = FastDiv(st.TW);
        p.fd_thw = FastDiv(st.TH * st.TW); p.fd_NQ = FastDiv(st.NC / 4); p.fd_tpg = FastDiv(st.tilesX * st.tilesY);
        p.fd_tilesX = FastDiv(st.tilesX);
        p.smem_bytes = st.smem;
        launch_dwpw(p, B, s, cta_cap(st.smem, st.NPG * (st.NC / 4)));
        break;
      }
      case kStepBlockWs: {
        DwPwTcP p;
        TV iv = view(ctx, st.in);
        p.in = iv.p; p.in_istride = iv.istride; p.H = iv.H; p.W = iv.W; p.Cin = iv.C; p.CinS = iv.Cs;
        p.has_dw = st.has_dw ? 1 : 0; p.s = st.dws; p.dpt = st.dpt; p.dpl = st.dpl;
        p.OH = out.H; p.OW = out.W;
        p.dww = st.dww >= 0 ? blob + st.dww : nullptr; p.dwb = st.dwb >= 0 ? blob + st.dwb : nullptr;
        p.K8 = st.K8; p.KS = st.KS;
        p.out = out.p; p.out_istride = out.istride; p.Cout = st.Cout; p.CoutS = out.Cs;
        p.vec_store = (out.Cs % 4 == 0 && out.istride % 4 == 0 && ((size_t)out.p % 16 == 0)) ? ((out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 2 : 1) : 0;   // 2: 32-byte stores allowed
        p.wB = blob + st.w; p.bias = blob + st.bias; p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act; p.Npad = st.Npad; p.tmem_cols = st.tmem_cols; p.a_rows = st.a_rows; p.RS = st.RS; p.w_parts = st.w_parts; p.nbuf = st.nbuf;
        p.res_mode = st.in2 >= 0 ? st.res_mode : 0;
        if (st.in2 >= 0) {
          TV rv = view(ctx, st.in2);
          p.res = rv.p; p.res_istride = rv.istride; p.res_H = rv.H; p.res_W = rv.W; p.res_C = rv.C; p.res_Cs = rv.Cs;
        } else {
          p.res = nullptr; p.res_istride = 0; p.res_H = p.res_W = p.res_C = p.res_Cs = 0;
        }
        p.res_pool = st.res_pool; p.res_lim = iv.Cs;
        p.TH = st.TH; p.TW = st.TW; p.G = st.G; p.IH = st.IH; p.IW = st.IW; p.tilesX = st.tilesX; p.tilesY = st.tilesY;
        p.fd_Q8 = FastDiv(st.K8 / 4); p.fd_IW = FastDiv(st.IW); p.fd_IH = FastDiv(st.IH); p.fd_TW = FastDiv(st.TW);
        p.fd_thw = FastDiv(st.TH * st.TW); p.fd_tpg = FastDiv(st.tilesX * st.tilesY); p.fd_tilesX = FastDiv(st.tilesX);
        p.fd_nstrips = FastDiv(st.TH / st.RS); p.fd_nslots = FastDiv(st.G * st.TH * st.TW);
        p.n_chunks = st.G * st.IH * st.IW * (st.K8 / 4);
        p.n_items = st.has_dw ? st.G * (st.K8 / 4) * (st.TH / st.RS) * st.TW : 0;
        {
          long long in = (long long)st.G * st.IH * st.IW * st.KS, tail = (long long)(128 - st.a_rows) * st.K8;
          p.in_floats = (int)((std::max(in, tail) + 3) / 4 * 4);
        }
        p.smem_bytes = st.smem;
        p.ns = st.ns; p.na = st.na; p.nd = st.nd; p.nt = st.nt; p.trace = nullptr;
        p.deint = st.deint; p.plane_floats = st.plane_floats; p.PW = st.PW;
        p.nr = st.in2 >= 0 ? st.nr : 0; p.KSr = st.KSr; p.res_stage_floats = st.res_stage_floats;
        p.no = st.no; p.KSo = st.KSo; p.out_stage_floats = st.out_stage_floats;
        p.out2 = nullptr; p.out2_istride = 0; p.Cs2 = 0; p.c1 = 0; p.c2 = 0;
        if (st.out2 >= 0) {
          TV o2 = view(ctx, st.out2);
          p.out2 = o2.p; p.out2_istride = o2.istride; p.Cs2 = o2.Cs; p.c1 = st.c1; p.c2 = st.c2;
          p.Cout = st.c1;                       // columns of the first output
        }
        p.in_floats = st.in_stage_floats;
        if (!launch_block_ws(p, B, ctx.cap, s)) { failed_ = true; fprintf(stderr, "fdt: k_block_ws could not be launched for step '%s'\n", st.name.c_str()); }
        break;
      }
      case kStepTailWs: {
        TailP p;
        TV iv = view(ctx, st.in);
        p.in = iv.p; p.in_istride = iv.istride; p.H = iv.H; p.W = iv.W; p.CinS = iv.Cs;
        p.blob = blob; p.layers = d_tail_[&st - plan_.steps.data()]; p.nlayers = (int)st.tail.size();
        p.nbuf = st.tail_nbuf;
        for (int k = 0; k < kTailMaxBufs; ++k) { p.buf_off[k] = st.tail_buf_off[k]; p.buf_ks[k] = st.tail_buf_ks[k]; }
        p.act_floats = st.tail_act_floats; p.in_bytes = st.tail_in_bytes;
        p.generic = 0;
        for (const TailLayerD& L : st.tail)
          if (L.kind >= 2 || L.act == 2 || L.w_parts != 1 || L.wscale != 1.f || L.rbuf != L.src || (L.kind != 1 && L.o1 >= 0)) p.generic = 1;
        p.blks = nullptr; p.nblks = 0; p.bias_floats = 0; p.in_px = iv.H * iv.W;
        for (int k = 0; k < 2; ++k) { p.rsrc[k] = nullptr; p.rs_istride[k] = 0; p.rs_cs[k] = p.rs_w[k] = p.rs_c[k] = 0; }
        if (st.tail_wide) {
          p.generic = 2;
          p.blks = d_blks_[&st - plan_.steps.data()]; p.nblks = (int)st.tail_blks.size(); p.bias_floats = st.tail_bias_floats;
          for (size_t k = 0; k < st.tail_rsrc.size() && k < 2; ++k) {
            TV rv = view(ctx, st.tail_rsrc[k]);
            p.rsrc[k] = rv.p; p.rs_istride[k] = rv.istride; p.rs_cs[k] = rv.Cs; p.rs_w[k] = rv.W; p.rs_c[k] = rv.C;
          }
        }
        p.last_a_layer = st.tail_last_a; p.wbuf_bytes = st.tail_wbuf; p.wdepth = st.tail_wdepth; p.tbuf_bytes = st.tail_tbuf; p.smem_bytes = st.smem;
        for (int k = 0; k < 4; ++k) { p.outs[k] = nullptr; p.out_istride[k] = 0; p.out_pix[k] = 0; }
        for (size_t k = 0; k < st.tail_outs.size(); ++k) {
          TV ov = view(ctx, st.tail_outs[k]);
          p.outs[k] = ov.p; p.out_istride[k] = ov.istride; p.out_pix[k] = ov.Cs;
        }
        if (!launch_tail_ws(p, B, ctx.cap, s)) { failed_ = true; fprintf(stderr, "fdt: k_tail_ws could not be launched for step '%s'\n", st.name.c_str()); }
        break;
      }
      case kStepFcTc: {
        FcP p;
        TV iv = view(ctx, st.in);
        p.in = iv.p; p.in_istride = iv.istride;
        p.out = out.p; p.out_istride = out.istride;
        p.w = blob + st.w; p.bias = blob + st.bias;
        p.K = st.K; p.K16 = st.K8; p.N = st.Cout; p.w_parts = st.w_parts; p.tile_bytes = st.ts_rec_bytes; p.wscale = (float)st.out_scale;
        if (!launch_fc_tc(p, B, s)) { failed_ = true; fprintf(stderr, "fdt: k_fc_tc could not be launched for step '%s'\n", st.name.c_str()); }
        break;
      }
      case kStepBlockTs: {
        BlockTsP p;
        TV iv = view(ctx, st.in);
        p.in = iv.p; p.in_istride = iv.istride; p.H = iv.H; p.W = iv.W; p.CinS = iv.Cs;
        p.out = out.p; p.out_istride = out.istride; p.OH = out.H; p.OW = out.W; p.CoutS = out.Cs;
        p.rec = blob + st.ts_rec; p.bias = blob + st.ts_bias; p.rec_bytes = st.ts_rec_bytes;
        p.Cin = iv.C; p.K16 = st.ts_k16; p.Npad = st.ts_npad; p.KS = st.ts_ks;
        p.stride = st.dws; p.res = st.ts_res; p.relu = st.ts_relu;
        p.wide = (out.Cs % 8 == 0 && out.istride % 8 == 0 && ((size_t)out.p % 32 == 0)) ? 1 : 0;
        p.ns = st.ts_ns; p.stage_bytes = st.ts_stage_bytes; p.smem_bytes = st.smem;
        {
          const float* hdw = plan_.blob.data() + st.ts_rec + (size_t)st.ts_npad * st.ts_k16 / 2;    // [9][K16] taps, [K16] bias (host copy)
          for (int q = 0; q < 16; ++q)                               // quad-major: 9 taps + bias of channels 4 q .. 4 q + 3 in one 160-byte run
            for (int k = 0; k < 10; ++k)                             // (immediate offsets in the kernel, two or three constant-cache lines per quad)
              for (int c = 0; c < 4; ++c) p.dw[(q * 10 + k) * 4 + c] = 4 * q + c < st.ts_k16 ? hdw[k * st.ts_k16 + 4 * q + c] : 0.f;
        }
        if (!launch_block_ts(p, B, ctx.cap, s)) { failed_ = true; fprintf(stderr, "fdt: k_block_ts could not be launched for step '%s'\n", st.name.c_str()); }
        break;
      }
      case kStepAdd: case kStepAct: case kStepPadC: {
        EltP p;
        p.a = view(ctx, st.in);
        p.b = st.in2 >= 0 ? view(ctx, st.in2) : p.a;
        p.out = out;
        p.alpha = st.alpha >= 0 ? blob + st.alpha : nullptr;
        p.act = st.act;
        if (st.kind == kStepAdd) launch_add(p, B, s);
        else if (st.kind == kStepAct) launch_act(p, B, s);
        else launch_padc(p, B, s);
        break;
      }
      case kStepMaxPool: {
        PoolP p;
        p.in = view(ctx, st.in); p.out = out;
        p.fh = st.fh; p.fw = st.fw; p.sh = st.sh; p.sw = st.sw; p.pt = st.pt; p.pl = st.pl;
        launch_maxpool(p, B, s);
        break;
      }
      case kStepResize: {
        ResizeP p;
        p.in = view(ctx, st.in); p.out = out; p.align_corners = st.align; p.half_pixel = st.half;
        p.has_add = st.in2 >= 0 ? 1 : 0; p.act = st.act;
        p.add = st.in2 >= 0 ? view(ctx, st.in2) : p.in;
        launch_resize_bilinear(p, B, s);
        break;
      }
    }
    ++launches;
  }
  if (evs) cudaEventRecord(evs[launches], s);
  return launches;
}

}  // namespace fdt
