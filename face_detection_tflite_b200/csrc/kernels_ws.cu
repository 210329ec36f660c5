// k_block_ws — warp-specialised, TMA-fed BlazeBlock kernel (sm_100a).
//
// Same arithmetic as k_dwpw_tc (depthwise 3x3 on CUDA cores -> TF32 hi/lo split -> tcgen05.mma into
// TMEM -> bias + residual + activation epilogue) but organised as an asynchronous pipeline so that a
// CTA always has several tiles in flight instead of walking load -> depthwise -> MMA -> epilogue
// for one tile at a time:
//
//   warp  4+ND   (1 lane)  producer : cp.async.bulk.tensor.4d (TMA) of the input tile + halo into a ring of
//                                     NS shared-memory stages; the tensor map's out-of-bounds zero fill is
//                                     TFLite's SAME padding, the image tail and the channel pad at once
//   warps 4..3+ND          depthwise: 3x3 depthwise (rolling 3-row window) of stage s -> hi/lo split ->
//                                     UMMA K-major core-matrix layout, ring of NA operand buffers
//   warp  5+ND   (1 lane)  MMA      : tcgen05.mma.kind::tf32 (A_hi*W + A_lo*W [+ A_hi*W_lo]) into one of two
//                                     TMEM accumulators; tcgen05.commit releases the operand buffer and
//                                     publishes the accumulator
//   warps 0..3             epilogue : tcgen05.ld (32 lanes x 16 columns), + bias + residual (from the still
//                                     resident input stage, optional 2x2 max-pool, or from HBM) + ReLU/PReLU,
//                                     float4 stores; releases the accumulator and the input stage
//
// All hand-offs are mbarriers (full/empty per ring slot); there is no __syncthreads in the steady state.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.h"
#include "tile_walk.h"

namespace fdt {
namespace {

constexpr uint32_t kLBO = 128;   // bytes between the two 16-byte K chunks of one MMA (adjacent core matrices)
constexpr int kEpiWarps = 4;
#ifndef FDT_MMA_WARPS
#define FDT_MMA_WARPS 2
#endif
constexpr int kMmaWarps = FDT_MMA_WARPS;   // issuing one tcgen05.mma costs its thread ~150 cycles: two issuers alternate tiles

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16_u32(uint32_t smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(smem_dst), "l"(gmem_src));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// arrive without release ordering: used where the signal only says "I have finished READING" (TMEM accumulator / input
// stage consumed into registers).  A releasing arrive would also wait for the epilogue's global stores to be performed.
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  }
}

// SMEM matrix descriptor, K-major, no swizzle: start >> 4 in [0,14), LBO >> 4 in [16,30), SBO >> 4 in [32,46),
// version 1 in [46,48), layout type 0 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kLBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

// The MMA issuer is ONE thread: every instruction it spends per tcgen05.mma is serial latency for the whole CTA.
// Descriptors are therefore kept as two 32-bit words: the low word (start address >> 4 | LBO) advances by a plain
// 32-bit add per K step (16 = 2 core matrices of 128 B, >> 4), the high word (SBO | version) is loop-invariant.
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16); }
__device__ __forceinline__ uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3FFFu) | (1u << 14); }
template <int F16>
__device__ __forceinline__ void mma_ss(uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  if (F16)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], da, db, %5, p;\n\t}\n" ::"r"(tmem_d), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
}
// A operand from tensor memory (kind::f16): used by k_stem_ws, whose im2col rows are written with tcgen05.st
__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void split_store(float* sAhi, float* sAlo, uint32_t off_floats, float4 a) {
  float4 hi, lo;
  hi.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); lo.x = a.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u); lo.y = a.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); lo.z = a.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u); lo.w = a.w - hi.w;
  *reinterpret_cast<float4*>(sAhi + off_floats) = hi;
  *reinterpret_cast<float4*>(sAlo + off_floats) = lo;
}

#ifndef FDT_NO_FFMA2
// two packed fp32 FMAs (sm_100 FFMA2): same rounding as four scalar fmaf, half the issue slots
__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
  asm("{\n\t.reg .b64 ra, rv, rw;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rv, {%4, %5};\n\tmov.b64 rw, {%8, %9};\n\t"
      "fma.rn.f32x2 ra, rv, rw, ra;\n\tmov.b64 {%0, %1}, ra;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rv, {%6, %7};\n\tmov.b64 rw, {%10, %11};\n\t"
      "fma.rn.f32x2 ra, rv, rw, ra;\n\tmov.b64 {%2, %3}, ra;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w));
}
#else
__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
  a.x = fmaf(v.x, w.x, a.x); a.y = fmaf(v.y, w.y, a.y); a.z = fmaf(v.z, w.z, a.z); a.w = fmaf(v.w, w.w, a.w);
}
#endif
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
// a >= 0 ? a : alpha * a   with alpha = 0 (ReLU), 1 (none) or the PReLU slope: max(a,0) + alpha * min(a,0), exact in all three cases
__device__ __forceinline__ float4 leaky4(const float4& a, const float4& al) {
  return make_float4(fmaf(al.x, fminf(a.x, 0.f), fmaxf(a.x, 0.f)), fmaf(al.y, fminf(a.y, 0.f), fmaxf(a.y, 0.f)),
                     fmaf(al.z, fminf(a.z, 0.f), fmaxf(a.z, 0.f)), fmaf(al.w, fminf(a.w, 0.f), fmaxf(a.w, 0.f)));
}

// ---- depthwise 3x3 of one work item: RS output rows of one channel quad at one tile column -------------------
// Row-streaming form: every staged input row is loaded once (3 LDS.128) and scattered into the accumulators of the
// up to three output rows it feeds, so only the accumulators stay live (no 3-row window in registers).  Each output
// still sums bias, then taps (ky, kx) in row-major order - the same order as the per-output form.
template <int S, int RS>
__device__ __forceinline__ void dw_item(uint32_t base, uint32_t row_b, uint32_t c1_b, uint32_t c2_b, const float4 (&w)[9], const float4& bias,
                                        float* sAhi, float* sAlo, uint32_t sl, uint32_t aq, uint32_t SBO, uint32_t TW) {
  // c1_b / c2_b: byte offsets of window columns 1 and 2 (one and two records, or - stride 2 with the even / odd input columns
  // staged as two planes - the odd plane and one record)
  constexpr int NR = S == 1 ? RS + 2 : 2 * RS + 1;     // staged rows this item reads
  float4 acc[3];                                        // acc[t % 3] = output row t
#pragma unroll
  for (int r = 0; r < NR; ++r) {
    float4 v[3];
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) v[kx] = lds4(base + (uint32_t)r * row_b + (kx == 0 ? 0u : (kx == 1 ? c1_b : c2_b)));
    if (S == 1) {
      if (r >= 2) {                                     // ky = 2 of output r-2: complete
        float4& a = acc[(r - 2) % 3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[6 + kx]);
        split_store(sAhi, sAlo, ((sl >> 3) * SBO + aq + (sl & 7u) * 16u) >> 2, a);
        sl += TW;
      }
      if (r >= 1 && r - 1 < RS) {                       // ky = 1 of output r-1
        float4& a = acc[(r - 1) % 3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[3 + kx]);
      }
      if (r < RS) {                                     // ky = 0 of output r
        float4& a = acc[r % 3];
        a = bias;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[kx]);
      }
    } else {
      const int t = r >> 1;
      if ((r & 1) == 0) {
        if (t >= 1) {                                   // ky = 2 of output t-1: complete
          float4& a = acc[(t - 1) % 3];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[6 + kx]);
          split_store(sAhi, sAlo, ((sl >> 3) * SBO + aq + (sl & 7u) * 16u) >> 2, a);
          sl += TW;
        }
        if (t < RS) {                                   // ky = 0 of output t
          float4& a = acc[t % 3];
          a = bias;
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[kx]);
        }
      } else {                                          // ky = 1 of output t
        float4& a = acc[t % 3];
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) fma4(a, v[kx], w[3 + kx]);
      }
    }
  }
}

// ---- fast epilogue (float4 stores, residual none / staged / staged + 2x2 max-pool, staged pixel stride KS >= CoutS) ----
// Straight-line per 8-column group: TMEM load, bias and residual LDS issued before tcgen05.wait::ld, add, activation,
// two float4 stores.  Channels >= Cout inside CoutS come out as exact zeros (zero weights, bias and TMA zero fill).
// 256-bit store (sm_100 STG.256): halves the number of store instructions, each of which costs the LSU one slot per
// 128-byte line it touches (24-32 lines at a CoutS-float pixel stride)
__device__ __forceinline__ void stg8(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// RES 3 / 4: residual from another HBM tensor (same size / 2x2 max-pooled), float4 loads; `gres` points at this thread's
// residual pixel (top-left of the pool window), gks / grow are its pixel / row strides in floats, gres_c its channels.
// SM = 1: the results go to the shared-memory output tile at `out_s` (then one TMA store per tile) instead of to HBM.
template <int RES, int LEAKY, int SM, int DUAL = 0>
__device__ __forceinline__ void epi_fast(uint32_t tcol0, uint32_t res_a, uint32_t bias_a, uint32_t alpha_a, float* orow, uint32_t out_s,
                                         bool valid, bool store_ok, int cout_s, uint32_t ks_b, uint32_t row_b,
                                         const float* gres = nullptr, int gks = 0, int grow = 0, int gres_c = 0, bool wide = false,
                                         float* orow2 = nullptr, int c1 = 1 << 30, int c2 = 0) {
  // 16 columns per TMEM round trip (8 for the pooled residuals, whose four taps per quad need the registers); the
  // accumulator is Npad = 16k columns wide, quads >= cout_s are computed on whatever lies there and dropped at the store
#ifdef FDT_EPI_X8
  constexpr int NQ = 2;
#else
  constexpr int NQ = (RES == 2 || RES == 4) ? 2 : 4;
#endif
#pragma unroll 1
  for (int c0 = 0; c0 < cout_s; c0 += 4 * NQ) {
    uint32_t u[16];
    if (NQ == 4)
      asm volatile(
          "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
          : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
            "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
          : "r"(tcol0 + (uint32_t)c0));
    else
      asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                   : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                   : "r"(tcol0 + (uint32_t)c0));
    const uint32_t cb = 4u * (uint32_t)c0;
    float4 bv[NQ], rv[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      const uint32_t o = cb + 16u * q;
      bv[q] = lds4(bias_a + o);
      rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (RES == 1) {
        rv[q] = lds4(res_a + o);
      } else if (RES == 2) {
        // gres_c != 0: the staged record holds only the residual's own gres_c channels (pooled tiles of the residual ring)
        if (gres_c == 0 || c0 + 4 * q < gres_c)
          rv[q] = max4(max4(lds4(res_a + o), lds4(res_a + ks_b + o)), max4(lds4(res_a + row_b + o), lds4(res_a + row_b + ks_b + o)));
      } else if (RES == 3) {
        if (valid && c0 + 4 * q < gres_c) rv[q] = ldg4(gres + c0 + 4 * q);
      } else if (RES == 4) {
        if (valid && c0 + 4 * q < gres_c)
          rv[q] = max4(max4(ldg4(gres + c0 + 4 * q), ldg4(gres + gks + c0 + 4 * q)), max4(ldg4(gres + grow + c0 + 4 * q), ldg4(gres + grow + gks + c0 + 4 * q)));
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float4 vq[NQ];
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float4 v = make_float4(__uint_as_float(u[4 * q]) + bv[q].x, __uint_as_float(u[4 * q + 1]) + bv[q].y,
                             __uint_as_float(u[4 * q + 2]) + bv[q].z, __uint_as_float(u[4 * q + 3]) + bv[q].w);
      if (RES) add4(v, rv[q]);
      if (LEAKY) v = leaky4(v, lds4(alpha_a + cb + 16u * q));
      else v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
      vq[q] = v;
    }
    if (DUAL) {
      // merged head pair: columns [0, c1) -> orow (float4 / 256-bit stores); columns [c1, c1 + c2) -> orow2 (dense
      // [.., c2] view of the other head: scalar stores)
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        const int c = c0 + 4 * q;
        if (!store_ok || c >= cout_s) continue;
        if (c < c1) {
          if (wide) { if ((q & 1) == 0) stg8(orow + c, vq[q], vq[q + 1 < NQ ? q + 1 : q]); }   // c1 % 8 == 0 when wide
          else *reinterpret_cast<float4*>(orow + c) = vq[q];
        } else {
          const float vv[4] = {vq[q].x, vq[q].y, vq[q].z, vq[q].w};
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (c - c1 + j < c2) orow2[c - c1 + j] = vv[j];
        }
      }
    } else if (!SM && wide) {
      // CoutS % 8 == 0 and a 32-byte aligned pixel record: one 256-bit store per pair of quads
#pragma unroll
      for (int q = 0; q < NQ; q += 2)
        if (store_ok && c0 + 4 * q < cout_s) stg8(orow + c0 + 4 * q, vq[q], vq[q + 1]);
    } else {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
        if (store_ok && c0 + 4 * q < cout_s) {
          if (SM) asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(out_s + 4u * (uint32_t)(c0 + 4 * q)), "f"(vq[q].x), "f"(vq[q].y), "f"(vq[q].z), "f"(vq[q].w) : "memory");
          else *reinterpret_cast<float4*>(orow + c0 + 4 * q) = vq[q];
        }
    }
  }
}

// ---- epilogue of one tile for the thread's pixel: TMEM -> + bias + residual -> activation -> HBM --------------
// RES: 0 none, 1 from the staged input tile, 2 from the staged tile with 2x2 max-pool, 3 from HBM (generic path).
// LEAKY: 0 = ReLU (max only), 1 = slope from sAlpha (1 = identity, PReLU slopes otherwise).
// Loads of an 8-column group (bias, slopes, residual) are issued before tcgen05.wait::ld so they overlap the TMEM read.
template <int RES, int LEAKY>
__device__ __forceinline__ void epi_tile(const DwPwTcP& p, uint32_t tcol0, uint32_t res_a, uint32_t bias_a, uint32_t alpha_a,
                                         float* orow, bool valid, const float* rbase, int oy, int ox, uint32_t ks_b, uint32_t row_b) {
  for (int c0 = 0; c0 < p.Npad; c0 += 8) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                 : "r"(tcol0 + (uint32_t)c0));
    float4 bv[2], al[2], rv[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = c0 + 4 * q;
      bv[q] = lds4(bias_a + 4u * c);
      if (LEAKY) al[q] = lds4(alpha_a + 4u * c);
      rv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (RES == 1) {
        if (c < p.res_lim) rv[q] = lds4(res_a + 4u * c);
      } else if (RES == 2) {
        if (c < p.res_lim)
          rv[q] = max4(max4(lds4(res_a + 4u * c), lds4(res_a + ks_b + 4u * c)), max4(lds4(res_a + row_b + 4u * c), lds4(res_a + row_b + ks_b + 4u * c)));
      } else if (RES == 3) {
        if (valid) {
          float r4[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int ch = c + j;
            r4[j] = 0.f;
            if (ch >= p.res_C) continue;
            if (p.res_pool) {
              float m = -INFINITY;
#pragma unroll
              for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                  int ry = 2 * oy + dy, rx = 2 * ox + dx;
                  if (ry < p.res_H && rx < p.res_W) m = fmaxf(m, rbase[((size_t)ry * p.res_W + rx) * p.res_Cs + ch]);
                }
              r4[j] = m;
            } else {
              r4[j] = rbase[((size_t)oy * p.res_W + ox) * p.res_Cs + ch];
            }
          }
          rv[q] = make_float4(r4[0], r4[1], r4[2], r4[3]);
        }
      }
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (!valid) continue;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = c0 + 4 * q;
      if (c >= p.CoutS) break;
      float4 v = make_float4(__uint_as_float(u[4 * q]) + bv[q].x, __uint_as_float(u[4 * q + 1]) + bv[q].y,
                             __uint_as_float(u[4 * q + 2]) + bv[q].z, __uint_as_float(u[4 * q + 3]) + bv[q].w);
      add4(v, rv[q]);
      if (LEAKY) v = leaky4(v, al[q]);
      else v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
      if (p.vec_store) {
        // lanes >= Cout need no masking: their weights and bias are zero and the residual is the zero
        // channel pad there, so they come out as exact zeros
        *reinterpret_cast<float4*>(orow + c) = v;
      } else {
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < p.Cout) orow[c + j] = vv[j];
      }
    }
  }
}

// debug trace: role r, CTA-local tile i, stamp j (0 = before waits, 1 = after waits, 2 = work done)
// (compiled in only with -DFDT_TRACE_BUILD: the stamps cost ~3 % even when the trace pointer is null)
#ifndef FDT_TRACE_BUILD
#define WS_TRACE(r, i, j) do { (void)tr; } while (0)
#else
#define WS_TRACE(r, i, j) do { if (tr && (i) < 64) tr[((r) * 64 + (i)) * 3 + (j)] = clock64(); } while (0)
#endif

// Shared-memory carve-up (must match plan_ws in plan.cpp):
//   [W (w_parts x Npad x K8)] [bias Npad] [alpha Npad] [dw taps+bias 10 x K8] [dtab n_items x 4 B] [barriers 256 B]
//   | 128-byte aligned: [A ring: NA x (hi, lo) x a_rows x K8] [input ring: NS x in_stage_bytes] [output tiles]
// S: depthwise stride (0 = no depthwise, pointwise only); RS: output rows per depthwise work item.
template <int ND, int S, int RS>
__global__ void __launch_bounds__((ND + kEpiWarps + 2 + kMmaWarps) * 32, 1)
k_block_ws(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_res,
           DwPwTcP p, int B, int ntiles) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int kThreads = (ND + kEpiWarps + 2 + kMmaWarps) * 32;
  constexpr int kDwThreads = ND * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel may begin its own prologue
  const int Q8 = p.K8 >> 2;                       // 16-byte K chunks per row
  const uint32_t SBO = (uint32_t)Q8 * 128u;       // bytes between 8-row groups
  const int NS = p.ns, NA = p.na;
  float* sB = smem;
  float* sBias = sB + (size_t)p.w_parts * p.Npad * p.K8;
  float* sAlpha = sBias + p.Npad;
  float* sDw = sAlpha + p.Npad;
  uint32_t* dtab = reinterpret_cast<uint32_t*>(sDw + (S ? 10 * p.K8 : 0));   // [n_items]: in_off/16 | slot0 << 14 | qq << 22
  uint64_t* bars = reinterpret_cast<uint64_t*>(dtab + ((p.n_items + 1) & ~1));
  // operand buffer: hi + lo parts of a_rows (<= 128) rows; the MMA always reads 128 rows, the surplus rows come from
  // whatever follows in shared memory (finite or not, they only feed accumulator rows that are never stored)
  const uint32_t a_part_floats = (uint32_t)p.a_rows * (uint32_t)p.K8, a_stage_floats = 2u * a_part_floats;
  float* sA = reinterpret_cast<float*>(((uintptr_t)(bars + 48) + 127) & ~(uintptr_t)127);
  float* sIn0 = sA + (size_t)NA * a_stage_floats;
  const uint32_t in_stage_floats = (uint32_t)p.in_floats;     // multiple of 32 floats (128 B)
  // barrier slots: full_in[NS] | empty_in[NS] | a_full[NA] | a_empty[NA] | d_full[NT] | d_empty[NT]   (NS <= 6, NA, NT <= 4)
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_in = bar0, empty_in = bar0 + 8u * 6, a_full = bar0 + 8u * 12, a_empty = bar0 + 8u * 16,
                 d_full = bar0 + 8u * 20, d_empty = bar0 + 8u * 24;
  const int NT = p.nt;
  // stride 2 with the even / odd input columns staged as two planes [G][IH][PW][KS] (plan_ws: deint): neighbouring output
  // columns read neighbouring records (conflict-free LDS.128); one TMA per plane, the map walks x with an element stride of 2
  const bool deint = S == 2 && p.deint != 0;
  const uint32_t plane_b = (uint32_t)p.plane_floats * 4u;
  const int RW = deint ? p.PW : p.IW;                 // staged row length in records
  const uint32_t tma_bytes = deint ? 2u * (uint32_t)(p.G * p.IH * p.PW * p.KS) * 4u : (uint32_t)(p.G * p.IH * p.IW * p.KS) * 4u;
  const int thw = p.TH * p.TW;
  const int nslots = p.G * thw;
  const bool epi_reads_stage = p.res_mode == 1;
  // residual in another HBM tensor, staged by TMA into its own ring (plan_ws: nr stages of [G][m TH][m TW][KSr], m = 2 when pooled):
  // the epilogue then reads it like a staged residual (kinds 1 / 2), several tiles' worth of it are in flight
  const int NR = p.res_mode == 2 ? p.nr : 0;
  const bool res_ring = NR > 0;
  const int rm = p.res_pool ? 2 : 1;
  const uint32_t ksr_b = (uint32_t)p.KSr * 4u, rowr_b = (uint32_t)(rm * p.TW) * ksr_b, res_stage_b = (uint32_t)p.res_stage_floats * 4u;
  const uint32_t res_tma_bytes = (uint32_t)(p.G * rm * p.TH * rm * p.TW * p.KSr) * 4u;
  // epilogue mode (uniform for the CTA; the epilogue and the store warp must agree on it)
  const int res_kind = p.res_mode == 1 ? (p.res_pool ? 2 : 1) : (p.res_mode == 2 ? (res_ring ? (p.res_pool ? 2 : 1) : 3) : 0);
  // global residual fast path: float4-aligned residual tensor; pooled windows must lie inside it (even dimensions)
  const bool gfast = res_kind == 3 && p.res_Cs % 4 == 0 && p.res_C % 4 == 0 && p.res_istride % 4 == 0 && ((size_t)p.res % 16 == 0) &&
                     (!p.res_pool || (p.res_H == 2 * p.OH && p.res_W == 2 * p.OW));
  const bool fast = p.vec_store && (res_kind == 3 ? gfast : (res_kind == 0 || res_ring || p.KS >= p.CoutS));
  const int gmul = p.res_pool ? 2 : 1, gks = p.res_Cs, grow = p.res_W * p.res_Cs;
  const bool tma_out = fast && p.no > 0;
  const bool wide = p.vec_store == 2;
  const uint32_t sOut_a = smem_u32(sIn0 + (size_t)NS * in_stage_floats), out_stage_b = (uint32_t)p.out_stage_floats * 4u, kso_b = (uint32_t)p.KSo * 4u;
  const uint32_t out_full = bar0 + 8u * 28, out_empty = bar0 + 8u * 30, full_res = bar0 + 8u * 32, empty_res = bar0 + 8u * 36;
  const uint32_t sRes_a = sOut_a + (uint32_t)p.no * out_stage_b;

  // ---- prologue ---------------------------------------------------------------------------------------
  // The producer lane arms the input ring and fires the CTA's first loads before anything else, so their latency
  // overlaps the weight / table set-up of the other threads.
  int early = 0;                                  // tiles already requested (producer lane only)
  if (warp == kEpiWarps + ND && lane == 0) {
    for (int i = 0; i < NS; ++i) {
      mbar_init(full_in + 8u * i, 1);
      mbar_init(empty_in + 8u * i, ND + (epi_reads_stage ? kEpiWarps : 0));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    if (p.no > 0) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_out) : "memory");
    if (res_ring) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_res) : "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");          // (PDL) the input is the previous kernel's output
    for (int tile = blockIdx.x; tile < ntiles && early < NS; tile += gridDim.x, ++early) {
      int grp, trem, tyi, txi;
      p.fd_tpg.divmod(tile, grp, trem);
      p.fd_tilesX.divmod(trem, tyi, txi);
      const uint32_t bar = full_in + 8u * early;
      mbar_expect_tx(bar, tma_bytes);
      asm volatile(
          "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
          ::"r"(smem_u32(sIn0 + (size_t)early * in_stage_floats)), "l"(&tmap), "r"(0), "r"(txi * p.TW * p.s - p.dpl),
            "r"(tyi * p.TH * p.s - p.dpt), "r"(grp * p.G), "r"(bar) : "memory");
      if (deint)
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(smem_u32(sIn0 + (size_t)early * in_stage_floats) + plane_b), "l"(&tmap), "r"(0), "r"(txi * p.TW * p.s + 1),
              "r"(tyi * p.TH * p.s), "r"(grp * p.G), "r"(bar) : "memory");
    }
  }
  // all threads: weights, bias, slopes, depthwise taps (cp.async), depthwise table
  const uint32_t sB_u32 = smem_u32(sB);
  for (int i = tid; i < p.w_parts * p.Npad * Q8; i += kThreads) cp_async16_u32(sB_u32 + 16u * i, p.wB + 4 * (size_t)i);
  for (int i = tid; i < (p.Npad >> 2); i += kThreads) cp_async16_u32(smem_u32(sBias) + 16u * i, p.bias + 4 * (size_t)i);
  if (p.act == kActPrelu) {
    for (int i = tid; i < (p.Npad >> 2); i += kThreads) cp_async16_u32(smem_u32(sAlpha) + 16u * i, p.alpha + 4 * (size_t)i);
  } else {
    for (int i = tid; i < p.Npad; i += kThreads) sAlpha[i] = p.act == kActRelu ? 0.f : 1.f;
  }
  if (S) {
    for (int i = tid; i < 9 * Q8; i += kThreads) cp_async16_u32(smem_u32(sDw) + 16u * i, p.dww + 4 * (size_t)i);
    for (int i = tid; i < Q8; i += kThreads) cp_async16_u32(smem_u32(sDw + 9 * p.K8) + 16u * i, p.dwb + 4 * (size_t)i);
    // depthwise table: item = (((g*Q8 + qq)*nstrips + st)*TW + tx)  (tx fastest => conflict-free LDS)
    for (int it = tid; it < p.n_items; it += kThreads) {
      int tx, r, st, r2, qq, g;
      p.fd_TW.divmod(it, r, tx);
      p.fd_nstrips.divmod(r, r2, st);
      p.fd_Q8.divmod(r2, g, qq);
      const int tyb = st * RS;
      uint32_t in_off = deint ? (uint32_t)(((g * p.IH + tyb * 2) * p.PW + tx) * p.KS + 4 * qq)
                              : (uint32_t)(((g * p.IH + tyb * S) * p.IW + tx * S) * p.KS + 4 * qq);
      uint32_t slot0 = (uint32_t)(g * thw + tyb * p.TW + tx);
      dtab[it] = (in_off >> 2) | (slot0 << 14) | ((uint32_t)qq << 22);
    }
  }
  if (tid == 0) {
    for (int i = 0; i < NA; ++i) {
      mbar_init(a_full + 8u * i, ND);
      mbar_init(a_empty + 8u * i, 1);
    }
    for (int i = 0; i < NT; ++i) {
      mbar_init(d_full + 8u * i, 1);
      mbar_init(d_empty + 8u * i, kEpiWarps);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(out_full + 8u * i, kEpiWarps);
      mbar_init(out_empty + 8u * i, 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(full_res + 8u * i, 1);
      mbar_init(empty_res + 8u * i, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + ND + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  cp_async_wait_all();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  // Programmatic dependent launch: everything above touched only constants (weights, tables, barriers, TMEM);
  // the activations written by the previous kernel in the stream are read (and ours written) only after this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");

  if (warp < kEpiWarps) {
    // =============================== epilogue warps =================================================
    const int slot = warp * 32 + lane;            // TMEM lane == pixel slot
    int e_g, e_r, e_ty, e_tx;
    p.fd_thw.divmod(slot, e_g, e_r);
    p.fd_TW.divmod(e_r, e_ty, e_tx);
    const bool slot_ok = slot < nslots;
    const long long o_rel = slot_ok ? (long long)e_g * p.out_istride + ((long long)e_ty * p.OW + e_tx) * p.CoutS : 0;
    const long long o_rel2 = slot_ok ? (long long)e_g * p.out2_istride + ((long long)e_ty * p.OW + e_tx) * p.Cs2 : 0;
    const int cout_loop = p.c2 > 0 ? p.c1 + ((p.c2 + 3) & ~3) : p.CoutS, c1 = p.c2 > 0 ? p.c1 : (1 << 30);
    const int rs = p.res_pool ? 2 : 1;
    const uint32_t res_off = !slot_ok ? 0u : deint ? (uint32_t)((((size_t)e_g * p.IH + e_ty * 2) * p.PW + e_tx) * p.KS)
                                                   : (uint32_t)((((size_t)e_g * p.IH + e_ty * rs + p.dpt) * p.IW + e_tx * rs + p.dpl) * p.KS);
    const uint32_t sIn0_a = smem_u32(sIn0), bias_a = smem_u32(sBias), alpha_a = smem_u32(sAlpha);
    int ob = 0, oph = 0;
    // byte distance of the residual's right / lower neighbour in the stage (pooled residuals only read them)
    const uint32_t ks_b = res_ring ? ksr_b : (deint ? plane_b : (uint32_t)p.KS * 4u), row_b = res_ring ? rowr_b : (uint32_t)RW * (uint32_t)p.KS * 4u;
    // this thread's residual record inside a stage of the residual ring
    const uint32_t rr_off = slot_ok ? (uint32_t)((((size_t)e_g * rm * p.TH + e_ty * rm) * rm * p.TW + e_tx * rm) * p.KSr) * 4u : 0u;
    const int ring_c = (res_ring && p.KSr < p.CoutS) ? p.KSr : 0;     // channel limit of a staged residual record narrower than the output
    int ri = 0, rph = 0;
    int si = 0, sph = 0, di = 0, dph = 0;
    long long* tr = (p.trace && blockIdx.x == 0 && tid == 0) ? p.trace : nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      int grp, trem, tyi, txi;
      p.fd_tpg.divmod(tile, grp, trem);
      p.fd_tilesX.divmod(trem, tyi, txi);
      const int ty0 = tyi * p.TH, tx0 = txi * p.TW;
      const int b0 = grp * p.G;
      WS_TRACE(0, it, 0);
      const uint32_t res_a = res_ring ? sRes_a + (uint32_t)ri * res_stage_b + rr_off : sIn0_a + ((uint32_t)si * in_stage_floats + res_off) * 4u;
      const int oy = ty0 + e_ty, ox = tx0 + e_tx, b = b0 + e_g;
      const bool valid = slot_ok && b < B && oy < p.OH && ox < p.OW;
      if (epi_reads_stage) mbar_wait(full_in + 8u * si, (uint32_t)sph);   // visibility of the TMA writes to this thread
      if (res_ring) mbar_wait(full_res + 8u * ri, (uint32_t)rph);
      mbar_wait(d_full + 8u * di, (uint32_t)dph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      WS_TRACE(0, it, 1);
      float* orow = p.out + (long long)b0 * p.out_istride + ((long long)ty0 * p.OW + tx0) * p.CoutS + o_rel;
      float* orow2 = p.c2 > 0 ? p.out2 + (long long)b0 * p.out2_istride + ((long long)ty0 * p.OW + tx0) * p.Cs2 + o_rel2 : nullptr;
      const float* rbase = p.res_mode == 2 ? p.res + (size_t)(valid ? b : 0) * p.res_istride : nullptr;
      const uint32_t tcol0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(di * p.Npad);
      if (fast) {
        const float* gres = res_kind == 3 ? rbase + ((size_t)(valid ? oy : 0) * gmul * p.res_W + (size_t)(valid ? ox : 0) * gmul) * p.res_Cs : nullptr;
        if (tma_out) {
          // the output tile is assembled in shared memory and leaves with one TMA store (full-line writes; direct
          // float4 stores at a CoutS-float pixel stride touch 24..32 lines per instruction and saturate the LSU)
          const uint32_t out_s = sOut_a + (uint32_t)ob * out_stage_b + (uint32_t)slot * kso_b;
          WS_TRACE(5, it, 0);
          mbar_wait(out_empty + 8u * ob, (uint32_t)(oph ^ 1));      // the store warp has finished reading buffer `ob`
          WS_TRACE(5, it, 1);
          if (res_kind == 3) {
            if (p.act == kActRelu) {
              if (p.res_pool) epi_fast<4, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b, gres, gks, grow, p.res_C);
              else epi_fast<3, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b, gres, gks, grow, p.res_C);
            } else {
              if (p.res_pool) epi_fast<4, 1, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b, gres, gks, grow, p.res_C);
              else epi_fast<3, 1, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b, gres, gks, grow, p.res_C);
            }
          } else if (p.act == kActRelu) {
            if (res_kind == 1) epi_fast<1, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
            else if (res_kind == 2) epi_fast<2, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
            else epi_fast<0, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
          } else {
            if (res_kind == 1) epi_fast<1, 1, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
            else if (res_kind == 2) epi_fast<2, 1, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
            else epi_fast<0, 1, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, slot_ok, p.CoutS, ks_b, row_b);
          }
          WS_TRACE(5, it, 2);
          WS_TRACE(6, it, 0);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile's STS -> visible to the TMA store
          WS_TRACE(6, it, 1);
          __syncwarp();
          if (lane == 0) mbar_arrive(out_full + 8u * ob);
          WS_TRACE(6, it, 2);
          if (++ob == p.no) { ob = 0; oph ^= 1; }
        } else {
          const uint32_t out_s = 0u;
          if (res_kind == 3) {
            if (p.act == kActRelu) {
              if (p.res_pool) epi_fast<4, 0, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, gres, gks, grow, p.res_C, wide);
              else epi_fast<3, 0, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, gres, gks, grow, p.res_C, wide);
            } else {
              if (p.res_pool) epi_fast<4, 1, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, gres, gks, grow, p.res_C, wide);
              else epi_fast<3, 1, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, gres, gks, grow, p.res_C, wide);
            }
          } else if (p.act == kActRelu) {
            if (res_kind == 1) epi_fast<1, 0, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, 0, wide);
            else if (res_kind == 2) epi_fast<2, 0, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, ring_c, wide);
            else epi_fast<0, 0, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, 0, wide);
          } else {
            if (res_kind == 1) epi_fast<1, 1, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, 0, wide);
            else if (res_kind == 2) epi_fast<2, 1, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, ring_c, wide);
            else if (p.c2 > 0) epi_fast<0, 1, 0, 1>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, 0, wide, orow2, c1, p.c2);
            else epi_fast<0, 1, 0>(tcol0, res_a, bias_a, alpha_a, orow, out_s, valid, valid, cout_loop, ks_b, row_b, nullptr, 0, 0, 0, wide);
          }
        }
      } else {
        if (res_kind == 1) epi_tile<1, 1>(p, tcol0, res_a, bias_a, alpha_a, orow, valid, rbase, oy, ox, ks_b, row_b);
        else if (res_kind == 2) epi_tile<2, 1>(p, tcol0, res_a, bias_a, alpha_a, orow, valid, rbase, oy, ox, ks_b, row_b);
        else if (res_kind == 3) epi_tile<3, 1>(p, tcol0, res_a, bias_a, alpha_a, orow, valid, rbase, oy, ox, ks_b, row_b);
        else epi_tile<0, 1>(p, tcol0, res_a, bias_a, alpha_a, orow, valid, rbase, oy, ox, ks_b, row_b);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive_relaxed(d_empty + 8u * di);
        if (epi_reads_stage) mbar_arrive(empty_in + 8u * si);     // release: the warp's residual loads from the stage precede the producer's refill
        if (res_ring) mbar_arrive(empty_res + 8u * ri);
      }
      if (res_ring && ++ri == NR) { ri = 0; rph ^= 1; }
      WS_TRACE(0, it, 2);
      if (++si == NS) { si = 0; sph ^= 1; }
      if (++di == NT) { di = 0; dph ^= 1; }
    }
  } else if (warp < kEpiWarps + ND) {
    // =============================== depthwise / A-operand warps ====================================
    const int dtid = tid - kEpiWarps * 32;
    const uint32_t ks_b = (uint32_t)p.KS * 4u, row_b = (uint32_t)RW * ks_b;
    const uint32_t c1_b = deint ? plane_b : ks_b, c2_b = deint ? ks_b : 2u * ks_b;
    const uint32_t sIn0_a = smem_u32(sIn0), sDw_a = smem_u32(sDw);
    // one work item per thread: its taps stay in registers for the whole kernel
    const bool hoist = S != 0 && p.n_items <= kDwThreads;
    const bool mine = hoist && dtid < p.n_items;
    uint32_t e0 = 0u;
    float4 w0[9], bias0 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int t = 0; t < 9; ++t) w0[t] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (S != 0 && mine) {
      e0 = dtab[dtid];
      const uint32_t wq = sDw_a + 16u * (e0 >> 22);
#pragma unroll
      for (int t = 0; t < 9; ++t) w0[t] = lds4(wq + (uint32_t)(t * p.K8) * 4u);
      bias0 = lds4(wq + (uint32_t)(9 * p.K8) * 4u);
    }
    int si = 0, sph = 0, ai = 0, aph = 0;
    long long* tr = (p.trace && blockIdx.x == 0 && dtid == 0) ? p.trace : nullptr;
    int it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
      WS_TRACE(1, it, 0);
      const uint32_t sIn_a = sIn0_a + (uint32_t)si * in_stage_floats * 4u;
      float* sAhi = sA + (size_t)ai * a_stage_floats;
      float* sAlo = sAhi + a_part_floats;
      mbar_wait(full_in + 8u * si, (uint32_t)sph);
      mbar_wait(a_empty + 8u * ai, (uint32_t)(aph ^ 1));
      WS_TRACE(1, it, 1);
      if (S != 0) {
        if (hoist) {
          if (mine) dw_item<S ? S : 1, RS>(sIn_a + ((e0 & 0x3FFFu) << 4), row_b, c1_b, c2_b, w0, bias0, sAhi, sAlo, (e0 >> 14) & 0xFFu, (e0 >> 22) * kLBO, SBO, (uint32_t)p.TW);
        } else {
          for (int it = dtid; it < p.n_items; it += kDwThreads) {
            const uint32_t e = dtab[it];
            const uint32_t wq = sDw_a + 16u * (e >> 22);
            float4 w[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) w[t] = lds4(wq + (uint32_t)(t * p.K8) * 4u);
            const float4 bias = lds4(wq + (uint32_t)(9 * p.K8) * 4u);
            dw_item<S ? S : 1, RS>(sIn_a + ((e & 0x3FFFu) << 4), row_b, c1_b, c2_b, w, bias, sAhi, sAlo, (e >> 14) & 0xFFu, (e >> 22) * kLBO, SBO, (uint32_t)p.TW);
          }
        }
      } else {
        // pointwise only: slot s <-> staged pixel s (IH = TH, IW = TW); slot fastest
        for (int it = dtid; it < nslots * Q8; it += kDwThreads) {
          int qq, sl;
          p.fd_nslots.divmod(it, qq, sl);
          const float4 a = lds4(sIn_a + ((uint32_t)sl * (uint32_t)p.KS + 4u * qq) * 4u);
          split_store(sAhi, sAlo, (((uint32_t)sl >> 3) * SBO + (uint32_t)qq * kLBO + ((uint32_t)sl & 7u) * 16u) >> 2, a);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full + 8u * ai);
        mbar_arrive(empty_in + 8u * si);
      }
      WS_TRACE(1, it, 2);
      if (++si == NS) { si = 0; sph ^= 1; }
      if (++ai == NA) { ai = 0; aph ^= 1; }
    }
  } else if (warp == kEpiWarps + ND) {
    // =============================== TMA producer ===================================================
    if (lane == 0) {
      int si = early == NS ? 0 : early, sph = early == NS ? 1 : 0;     // the first `early` tiles were requested in the prologue
      long long* tr = (p.trace && blockIdx.x == 0) ? p.trace : nullptr;
      int it = early;
      for (int tile = blockIdx.x + early * gridDim.x; tile < ntiles; tile += gridDim.x, ++it) {
        int grp, trem, tyi, txi;
        p.fd_tpg.divmod(tile, grp, trem);
        p.fd_tilesX.divmod(trem, tyi, txi);
        const int b0 = grp * p.G;
        const int iy0 = tyi * p.TH * p.s - p.dpt, ix0 = txi * p.TW * p.s - p.dpl;
        WS_TRACE(2, it, 0);
        mbar_wait(empty_in + 8u * si, (uint32_t)(sph ^ 1));
        WS_TRACE(2, it, 1);
        const uint32_t bar = full_in + 8u * si;
        mbar_expect_tx(bar, tma_bytes);
        const uint32_t dst = smem_u32(sIn0 + (size_t)si * in_stage_floats);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(dst), "l"(&tmap), "r"(0), "r"(ix0), "r"(iy0), "r"(b0), "r"(bar) : "memory");
        if (deint)
          asm volatile(
              "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
              ::"r"(dst + plane_b), "l"(&tmap), "r"(0), "r"(ix0 + 1), "r"(iy0), "r"(b0), "r"(bar) : "memory");
        WS_TRACE(2, it, 2);
        if (++si == NS) { si = 0; sph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == kEpiWarps + ND + 1 + kMmaWarps) {
    // =============================== output store warp ==============================================
    // One TMA store per tile from the shared-memory output tile the epilogue warps have filled; keeps the bulk-group
    // bookkeeping (commit / wait_group.read) off the epilogue's critical path.
    if (lane == 0 && res_ring) {
      // residual producer: one TMA per tile into the residual ring
      int ri = 0, rph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int grp, trem, tyi, txi;
        p.fd_tpg.divmod(tile, grp, trem);
        p.fd_tilesX.divmod(trem, tyi, txi);
        mbar_wait(empty_res + 8u * ri, (uint32_t)(rph ^ 1));
        const uint32_t bar = full_res + 8u * ri;
        mbar_expect_tx(bar, res_tma_bytes);
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(sRes_a + (uint32_t)ri * res_stage_b), "l"(&tmap_res), "r"(0), "r"(txi * p.TW * rm), "r"(tyi * p.TH * rm), "r"(grp * p.G), "r"(bar) : "memory");
        if (++ri == NR) { ri = 0; rph ^= 1; }
      }
    }
    if (lane == 0 && tma_out) {
      int ob = 0, oph = 0, prev = -1;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int grp, trem, tyi, txi;
        p.fd_tpg.divmod(tile, grp, trem);
        p.fd_tilesX.divmod(trem, tyi, txi);
        mbar_wait(out_full + 8u * ob, (uint32_t)oph);
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];"
                     ::"l"(&tmap_out), "r"(0), "r"(txi * p.TW), "r"(tyi * p.TH), "r"(grp * p.G), "r"(sOut_a + (uint32_t)ob * out_stage_b) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        if (p.no == 2) {
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // every store but the newest has read its tile
          if (prev >= 0) mbar_arrive(out_empty + 8u * prev);
          prev = ob;
        } else {
          asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          mbar_arrive(out_empty + 8u * ob);
        }
        if (++ob == p.no) { ob = 0; oph ^= 1; }
      }
      asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");            // all output tiles written before the CTA exits
    }
    __syncwarp();
  } else {
    // =============================== MMA issuer =====================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = p.K8 >> 3;
      const uint32_t hi = desc_hi(SBO);
      const uint32_t w_hi0 = desc_lo(sB_u32), w_lo0 = desc_lo(sB_u32 + (uint32_t)p.Npad * p.K8 * 4u);
      const uint32_t a0 = desc_lo(smem_u32(sA)), a_buf = (a_stage_floats * 4u) >> 4, a_half = (a_part_floats * 4u) >> 4;
      const bool w_split = p.w_parts > 1;
      // Issuer mi takes the CTA's tiles it = mi, mi + nmi, ...  A parity wait must see every phase of its barrier, so a
      // slot may only ever be visited by one issuer: two issuers need even ring sizes, otherwise issuer 0 does all tiles.
      const int nmi = (NA % kMmaWarps == 0 && NT % kMmaWarps == 0) ? kMmaWarps : 1;
      const int mi = warp - (kEpiWarps + ND + 1);
      int ai = mi, aph = 0, di = mi, dph = 0;
      while (ai >= NA) { ai -= NA; aph ^= 1; }
      while (di >= NT) { di -= NT; dph ^= 1; }
      long long* tr = (p.trace && blockIdx.x == 0 && mi == 0) ? p.trace : nullptr;
      int it = mi;
      for (int tile = mi < nmi ? blockIdx.x + mi * gridDim.x : ntiles; tile < ntiles; tile += nmi * gridDim.x, it += nmi) {
        WS_TRACE(3, it, 0);
        mbar_wait(a_full + 8u * ai, (uint32_t)aph);
        mbar_wait(d_empty + 8u * di, (uint32_t)(dph ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        WS_TRACE(3, it, 1);
        const uint32_t dcol = tmem_base + (uint32_t)(di * p.Npad);
        uint32_t ah = a0 + (uint32_t)ai * a_buf, al = ah + a_half, wh = w_hi0, wl = w_lo0, acc = 0u;
#ifdef FDT_OLD_MMA
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t db = ((uint64_t)hi << 32) | (wh + 16u * ks), dah = ((uint64_t)hi << 32) | (ah + 16u * ks), dal = ((uint64_t)hi << 32) | (al + 16u * ks);
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol), "l"(dah), "l"(db), "r"(idesc), "r"(ks > 0 ? 1u : 0u));
          asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol), "l"(dal), "l"(db), "r"(idesc), "r"(1u));
          if (w_split) { const uint64_t dbl = ((uint64_t)hi << 32) | (wl + 16u * ks);
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(dcol), "l"(dah), "l"(dbl), "r"(idesc), "r"(1u)); }
        }
        (void)acc;
#else
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ++ks) {
          mma_ss<0>(dcol, ah, hi, wh, hi, idesc, acc);
          mma_ss<0>(dcol, al, hi, wh, hi, idesc, 1u);
          // fp32 weights (face_landmark): W = W_hi + W_lo, third product A_hi * W_lo (A_lo * W_lo ~ 2^-22, dropped)
          if (w_split) mma_ss<0>(dcol, ah, hi, wl, hi, idesc, 1u);
          ah += 16u; al += 16u; wh += 16u; wl += 16u; acc = 1u;
        }
#endif
        WS_TRACE(4, it, 0);
        mma_commit(a_empty + 8u * ai);   // operand buffer reusable once these MMAs have read it
        WS_TRACE(4, it, 1);
        mma_commit(d_full + 8u * di);    // accumulator complete
        WS_TRACE(4, it, 2);
        WS_TRACE(3, it, 2);
        ai += nmi; while (ai >= NA) { ai -= NA; aph ^= 1; }
        di += nmi; while (di >= NT) { di -= NT; dph ^= 1; }
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kEpiWarps + ND + 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

// ------------------------------------------------------------------------------------------------
// k_stem_ws — the KW x KW / stride-2 input convolution as an exact fp16 im2col GEMM.
//
// u8 pixel values minus 127.5 are half-integers in [-127.5, 127.5]: exact in fp16, as are the fp16-origin weights,
// so one tcgen05.mma.kind::f16 per K step reproduces the fp32 convolution of the normalised image up to the
// accumulation order (fp32 weights: three fp16 terms).  The im2col matrix costs no arithmetic at all: the converted
// patch keeps 4 halves {B,G,R,0} per pixel, so the KW taps of one kernel row are SEGP consecutive pixels = CPK
// 16-byte chunks that are copied verbatim (LDS.128 -> STS.128) into the UMMA K-major core-matrix layout.
//
//   warp 12 (1 lane) producer: TMA of the raw u8x4 patch [PH][40] (zero fill outside the image) into a 6-stage ring
//   warps 4..11     builders : two groups of four warps that take alternate tiles; per tile: raw -> fp16 patch (PRMT + 2 HSUB2 per
//                              pixel pair, 0 outside the image), then thread = TMEM lane = output pixel copies its im2col row
//                              (KW x CPK 16-byte chunks, LDS.128) straight into TENSOR MEMORY (tcgen05.st): the A operand never
//                              touches shared memory (round 1 wrote it there: 240 STS + 256 tensor-core read wavefronts per tile of
//                              the ~1,430 that bound the kernel on the LSU pipe)
//   warp 13 (1 lane) MMA     : K8/16 x w_parts tcgen05.mma (A from TMEM, W from shared memory) into one of four TMEM accumulators
//   warps 0..3      epilogue : D * out_scale + bias, ReLU/PReLU, float4 stores
// Two such CTAs share an SM (registers are allocated per four warps: 14 warps x 56 registers fit twice, 18 would not).
// Warps: 4*kStemEG epilogue, 8 builders, 1 producer, 1 MMA.  Registers are allocated per 4 warps, so 14 warps x 56 registers
// let two CTAs share an SM (16 builder + 8 epilogue warps per SM); 18 warps would not.
constexpr int kStemEG = 1;                       // epilogue groups of four warps (alternate tiles when 2)
constexpr int kStemThreads = (4 * kStemEG + 10) * 32;
// (launch bound declared for 576 threads so that ptxas caps the kernel at 56 registers: 2 x 16 allocated warps x 56 fit the SM)
template <int KW>
__global__ void __launch_bounds__(576, 2) k_stem_ws(const __grid_constant__ CUtensorMap tmap, StemWsP p, int B, int ntiles) {
  constexpr int TH = 8, TW = 16, PWP = 36;
  constexpr int PH = (TH - 1) * 2 + KW;
  constexpr int SEGP = KW == 5 ? 6 : 4, CPK = SEGP / 2;
  constexpr int NCH = (KW * CPK + 1) / 2 * 2;          // 16-byte K chunks per row (even)
  constexpr int K8 = NCH * 8;                          // halves
  constexpr uint32_t SBO = (uint32_t)NCH * 128u;
#ifndef FDT_STEM_NA
#define FDT_STEM_NA 2
#endif
  constexpr int NS = 6, NA = FDT_STEM_NA;           // (NA only sizes the plan's shared-memory estimate now: A lives in TMEM)
  // TMEM: accumulators [0, 128) (four of <= 32 columns, or two of <= 64), operand slot of builder group g at 128 + 64 g:
  // 256 columns per CTA, two CTAs per SM
  const int NT = p.Npad <= 32 ? 4 : 2, nt_sh = p.Npad <= 32 ? 2 : 1;
  constexpr uint32_t kStemTmemCols = 256, kStemACol = 128;
  constexpr int RAWW = 40;                             // raw patch row: the TMA box must start 16-byte aligned in x, so it
                                                       // begins up to 3 pixels left of the patch (offset rx_off) and is 40 wide
  constexpr uint32_t RAW_STAGE = (PH * RAWW * 4 + 127) / 128 * 128;
  constexpr uint32_t HP_BYTES = PH * PWP * 8;
  constexpr uint32_t A_BYTES = 128u * K8 * 2u;
  extern __shared__ __align__(128) float smem[];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel may begin its own prologue
  uint8_t* sW = reinterpret_cast<uint8_t*>(smem);
  const uint32_t w_bytes = (uint32_t)p.w_parts * p.Npad * K8 * 2u;
  float* sBias = reinterpret_cast<float*>(sW + w_bytes);
  float* sAlpha = sBias + p.Npad;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAlpha + p.Npad);
  uint8_t* sHp = reinterpret_cast<uint8_t*>(((uintptr_t)(bars + 32) + 127) & ~(uintptr_t)127);   // 2 groups x 2 fp16 patches
  uint8_t* sRaw = sHp + 4 * ((HP_BYTES + 127) / 128 * 128);
  (void)A_BYTES; (void)NA;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_raw = bar0, empty_raw = bar0 + 8u * 6, a_full = bar0 + 8u * 12, a_empty = bar0 + 8u * 16,
                 d_full = bar0 + 8u * 20, d_empty = bar0 + 8u * 24;

  // ---- prologue ----
  const uint32_t sW_u32 = smem_u32(sW);
  for (int i = tid; i < (int)(w_bytes >> 4); i += kStemThreads) cp_async16_u32(sW_u32 + 16u * i, reinterpret_cast<const uint8_t*>(p.wB) + 16 * (size_t)i);
  for (int i = tid; i < p.Npad; i += kStemThreads) {
    sBias[i] = p.bias[i];
    sAlpha[i] = p.act == kActPrelu ? p.alpha[i] : (p.act == kActRelu ? 0.f : 1.f);
  }
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(full_raw + 8u * i, 1); mbar_init(empty_raw + 8u * i, 4); }
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + 8u * i, 4); mbar_init(a_empty + 8u * i, 1); }
    for (int i = 0; i < NT; ++i) { mbar_init(d_full + 8u * i, 1); mbar_init(d_empty + 8u * i, 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  constexpr int kW0 = 4 * kStemEG;                  // first builder warp
  if (warp == kW0 + 9) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(kStemTmemCols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  cp_async_wait_all();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  // Programmatic dependent launch: everything above touched only constants (weights, tables, barriers, TMEM);
  // the activations written by the previous kernel in the stream are read (and ours written) only after this point.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int tilesX = (p.OW + TW - 1) / TW, tilesY = (p.OH + TH - 1) / TH;
  const int tiles_per_img = tilesX * tilesY;
  const int rx_off = (4 - (p.pl & 3)) & 3;             // ix0 = 32*txi - pl  ->  aligned start ix0 - rx_off

  if (warp < kW0) {
    // =============================== epilogue: group g takes every kStemEG-th tile of this CTA ==========
    const int g = warp >> 2, quarter = warp & 3;
    const int slot = quarter * 32 + lane;
    const int e_ty = slot >> 4, e_tx = slot & 15;
    const uint32_t bias_a = smem_u32(sBias), alpha_a = smem_u32(sAlpha);
    const float scale = p.out_scale;
    const bool relu = p.act == kActRelu;
    int k = 0;
    const TileStep step = tile_step(kStemEG * (int)gridDim.x, tiles_per_img, tilesX);        // (tile_walk.h: no division per tile)
    TileAt at = tile_at((int)(blockIdx.x + g * gridDim.x), tiles_per_img, tilesX);
    for (int tile = blockIdx.x + g * gridDim.x; tile < ntiles; tile += kStemEG * gridDim.x, ++k, tile_advance(at, step, tilesX, tilesY)) {
      const int b = at.b;
      const int ty0 = at.ty * TH, tx0 = at.tx * TW;
      const int oy = ty0 + e_ty, ox = tx0 + e_tx;
      const bool valid = oy < p.OH && ox < p.OW;
      float* orow = p.out + (size_t)b * p.out_istride + ((size_t)(valid ? oy : 0) * p.OW + (valid ? ox : 0)) * p.CoutS;
      const int kl = kStemEG * k + g, di = kl & (NT - 1);  // CTA-local tile index -> accumulator buffer
      mbar_wait(d_full + 8u * di, (uint32_t)((kl >> nt_sh) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t tcol0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(di * p.Npad);
#pragma unroll 1
      for (int c0 = 0; c0 < p.CoutS; c0 += 8) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(tcol0 + (uint32_t)c0));
        const uint32_t cb = 4u * (uint32_t)c0;
        const float4 b0 = lds4(bias_a + cb), b1 = lds4(bias_a + cb + 16u);
        const float4 a0 = lds4(alpha_a + cb), a1 = lds4(alpha_a + cb + 16u);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        float4 v0 = make_float4(fmaf(__uint_as_float(u[0]), scale, b0.x), fmaf(__uint_as_float(u[1]), scale, b0.y),
                                fmaf(__uint_as_float(u[2]), scale, b0.z), fmaf(__uint_as_float(u[3]), scale, b0.w));
        float4 v1 = make_float4(fmaf(__uint_as_float(u[4]), scale, b1.x), fmaf(__uint_as_float(u[5]), scale, b1.y),
                                fmaf(__uint_as_float(u[6]), scale, b1.z), fmaf(__uint_as_float(u[7]), scale, b1.w));
        if (relu) { v0 = max4(v0, make_float4(0.f, 0.f, 0.f, 0.f)); v1 = max4(v1, make_float4(0.f, 0.f, 0.f, 0.f)); }
        else { v0 = leaky4(v0, a0); v1 = leaky4(v1, a1); }
        if (valid) {
          if (p.vec_store == 2) {
            stg8(orow + c0, v0, v1);                 // CoutS % 8 == 0: one 256-bit store per 8 channels
          } else if (p.vec_store) {
            *reinterpret_cast<float4*>(orow + c0) = v0;
            if (c0 + 4 < p.CoutS) *reinterpret_cast<float4*>(orow + c0 + 4) = v1;
          } else {
            const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (c0 + j < p.Cout) orow[c0 + j] = vv[j];
          }
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive_relaxed(d_empty + 8u * di);
    }
  } else if (warp < kW0 + 8) {
    // =============================== builders: raw patch -> fp16 patch -> im2col row -> TMEM ==============
    const int grp = (warp - kW0) >> 2, quarter = warp & 3;     // group g takes the CTA's tiles k = g, g + 2, ...; hardware lane quarter = warp % 4
    const int bt = (quarter << 5) | lane;                      // thread of the group = row of the M-tile = TMEM lane
    const int r_ty = bt >> 4, r_tx = bt & 15;
    const uint32_t src_px = (uint32_t)((2 * r_ty) * PWP + 2 * r_tx) * 8u;
    const uint32_t sHp_a = smem_u32(sHp), sRaw_a = smem_u32(sRaw);
    constexpr uint32_t HP_STRIDE = (HP_BYTES + 127) / 128 * 128;
    const uint32_t tm_a = tmem_base + ((uint32_t)(quarter * 32) << 16) + kStemACol + 64u * (uint32_t)grp;
    // the K pad (chunks KW * CPK .. NCH - 1) of this group's operand slot stays zero for the whole kernel
#pragma unroll
    for (int j = KW * CPK; j < NCH; ++j)
      asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %1, %1, %1};" ::"r"(tm_a + 4u * (uint32_t)j), "r"(0u) : "memory");
    int hb = 0;
    const TileStep step = tile_step(2 * (int)gridDim.x, tiles_per_img, tilesX);              // (tile_walk.h: no division per tile)
    TileAt at = tile_at((int)(blockIdx.x + grp * gridDim.x), tiles_per_img, tilesX);
    int si = grp % NS;                                                                       // ring stage k % NS and its phase (k / NS) & 1
    uint32_t raw_phase = (uint32_t)((grp / NS) & 1);
    for (int k = grp; blockIdx.x + (long long)k * gridDim.x < ntiles; k += 2) {
      const int iy0 = at.ty * TH * 2 - p.pt, ix0 = at.tx * TW * 2 - p.pl;
      mbar_wait(full_raw + 8u * si, raw_phase);
      // ---- convert: u8x4 BGRX -> {B,G,R,0} - 127.5 as halves; pixels outside the image (SAME padding) -> 0
      const uint32_t raw_a = sRaw_a + (uint32_t)si * RAW_STAGE, hp_a = sHp_a + (uint32_t)(2 * grp + hb) * HP_STRIDE;
      for (int i = bt; i < PH * PWP; i += 128) {
        const int ly = i / PWP, lx = i - ly * PWP;
        uint32_t raw;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(raw) : "r"(raw_a + 4u * (uint32_t)(ly * RAWW + lx + rx_off)));
        const bool ok = (unsigned)(iy0 + ly) < (unsigned)p.H && (unsigned)(ix0 + lx) < (unsigned)p.W;
        uint32_t lo = __byte_perm(raw, 0x64646464u, 0x4140), hi = __byte_perm(raw, 0x64646464u, 0x4342);   // halves 1024 + byte
        __half2 l = *reinterpret_cast<__half2*>(&lo), h = *reinterpret_cast<__half2*>(&hi);
        const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f), k127 = __floats2half2_rn(127.5f, 127.5f);
        l = __hsub2(__hsub2(l, k1024), k127);
        h = __hsub2(__hsub2(h, k1024), k127);
        uint32_t lo2 = *reinterpret_cast<uint32_t*>(&l), hi2 = *reinterpret_cast<uint32_t*>(&h) & 0x0000FFFFu;
        if (!ok) { lo2 = 0u; hi2 = 0u; }
        asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(hp_a + 8u * i), "r"(lo2), "r"(hi2) : "memory");
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_raw + 8u * si);   // raw stage consumed
      // patch complete; also: every thread of the group has finished gathering from the patch of two tiles ago (same buffer)
      if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory"); else asm volatile("bar.sync 2, 128;" ::: "memory");
      mbar_wait(a_empty + 8u * (uint32_t)grp, (uint32_t)((((k >> 1) & 1)) ^ 1));   // the MMAs of this group's previous tile have read the slot
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      // ---- im2col row of this thread's output pixel: CPK 16-byte chunks per kernel row, LDS.128 -> four TMEM columns each
      const uint32_t src = hp_a + src_px;
#pragma unroll
      for (int ky = 0; ky < KW; ++ky) {
        uint4 v[CPK];
#pragma unroll
        for (int c = 0; c < CPK; ++c)
          asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v[c].x), "=r"(v[c].y), "=r"(v[c].z), "=r"(v[c].w)
                       : "r"(src + (uint32_t)(ky * PWP * 8 + c * 16)));
#pragma unroll
        for (int c = 0; c < CPK; ++c)
          asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                       ::"r"(tm_a + 4u * (uint32_t)(ky * CPK + c)), "r"(v[c].x), "r"(v[c].y), "r"(v[c].z), "r"(v[c].w) : "memory");
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8u * (uint32_t)grp);
      hb ^= 1;
      tile_advance(at, step, tilesX, tilesY);
      for (int j = 0; j < 2; ++j) if (++si == NS) { si = 0; raw_phase ^= 1u; }
    }
  } else if (warp == kW0 + 8) {
    // =============================== TMA producer ===================================================
    if (lane == 0) {
      int si = 0, sph = 0;
      const TileStep step = tile_step((int)gridDim.x, tiles_per_img, tilesX);
      TileAt at = tile_at((int)blockIdx.x, tiles_per_img, tilesX);
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, tile_advance(at, step, tilesX, tilesY)) {
        const int b = at.b;
        const int iy0 = at.ty * TH * 2 - p.pt, ix0 = at.tx * TW * 2 - p.pl;
        mbar_wait(empty_raw + 8u * si, (uint32_t)(sph ^ 1));
        const uint32_t bar = full_raw + 8u * si;
        {
        mbar_expect_tx(bar, (uint32_t)(PH * RAWW * 4));
        const uint32_t dst = smem_u32(sRaw) + (uint32_t)si * RAW_STAGE;
        asm volatile(
            "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
            ::"r"(dst), "l"(&tmap), "r"(ix0 - rx_off), "r"(iy0), "r"(b), "r"(bar) : "memory");
        }
        if (++si == NS) { si = 0; sph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =============================== MMA issuer =====================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);   // D f32, A/B f16, K-major
      const uint32_t part_q = ((uint32_t)p.Npad * K8 * 2u) >> 4;   // one weight part, in descriptor address units
      const uint32_t hi = desc_hi(SBO), w_desc0 = desc_lo(sW_u32);
      const int parts = p.w_parts;
      int k = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++k) {
        const int di = k & (NT - 1), slot = k & 1;
        mbar_wait(a_full + 8u * (uint32_t)slot, (uint32_t)((k >> 1) & 1));
        mbar_wait(d_empty + 8u * di, (uint32_t)(((k >> nt_sh) & 1) ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcol = tmem_base + (uint32_t)(di * p.Npad);
        const uint32_t acol = tmem_base + kStemACol + 64u * (uint32_t)slot;
        uint32_t acc = 0u;
#pragma unroll
        for (int ks = 0; ks < K8 / 16; ++ks) {
          mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, w_desc0 + 16u * ks, hi, idesc, acc);
          acc = 1u;
          if (parts > 1) mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, w_desc0 + part_q + 16u * ks, hi, idesc, 1u);
          if (parts > 2) mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, w_desc0 + 2u * part_q + 16u * ks, hi, idesc, 1u);
        }
        mma_commit(a_empty + 8u * (uint32_t)slot);
        mma_commit(d_full + 8u * di);
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kW0 + 9) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kStemTmemCols));
  }
}

template <typename Kern, typename... Args>
void launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t s, Args... args) {
  // (programmatic dependent launch was measured slower with two streams - the other stream already fills the gaps - so the
  //  kernels are launched plainly; their griddepcontrol instructions are no-ops then)
  kern<<<grid, block, smem, s>>>(args...);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// Tensor map of the block's input activation: f32 [cap][H][W][CinS], box {KS, IW, IH, G} (box dims beyond the
// tensor are zero-filled: channel pad, SAME padding, batch tail).
bool input_tensor_map(const DwPwTcP& p, int cap, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int, long long> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  Key key(p.in, cap, p.H, p.W, p.CinS, p.KS, p.deint ? -p.IW : p.IW, p.IH, p.G, p.in_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.CinS, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.CinS * 4, (cuuint64_t)p.W * p.CinS * 4, (cuuint64_t)p.in_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.KS, (cuuint32_t)p.IW, (cuuint32_t)p.IH, (cuuint32_t)p.G};
  cuuint32_t estr[4] = {1, p.deint ? 2u : 1u, 1, 1};          // deint: every other column (IW traversed -> PW = (IW + 1) / 2 delivered)
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  cache[key] = tm;
  *out = tm;
  return true;
}

// Tensor map of the block's residual tensor: f32 [cap][res_H][res_W][res_Cs], box {KSr, m TW, m TH, G} (m = 2 when the residual is
// 2x2 max-pooled); channels >= res_Cs are zero-filled: the residual's zero channel pad.
bool res_tensor_map(const DwPwTcP& p, int cap, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int, long long> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  const int m = p.res_pool ? 2 : 1;
  Key key(p.res, cap, p.res_H, p.res_W, p.res_Cs, p.KSr, m * p.TW, m * p.TH, p.G, p.res_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.res_Cs, (cuuint64_t)p.res_W, (cuuint64_t)p.res_H, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.res_Cs * 4, (cuuint64_t)p.res_W * p.res_Cs * 4, (cuuint64_t)p.res_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.KSr, (cuuint32_t)(m * p.TW), (cuuint32_t)(m * p.TH), (cuuint32_t)p.G};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.res), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 256) cache.clear();
  cache[key] = tm;
  *out = tm;
  return true;
}

// Tensor map of the block's output activation: f32 [cap][OH][OW][CoutS], box {KSo, TW, TH, G}; channels >= CoutS of the
// shared-memory tile, pixels outside the image and images >= cap are dropped by the store.
bool output_tensor_map(const DwPwTcP& p, int cap, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, int, int, long long> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  Key key(p.out, cap, p.OH, p.OW, p.CoutS, p.KSo, p.TW, p.TH, p.G, p.out_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.CoutS, (cuuint64_t)p.OW, (cuuint64_t)p.OH, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.CoutS * 4, (cuuint64_t)p.OW * p.CoutS * 4, (cuuint64_t)p.out_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.KSo, (cuuint32_t)p.TW, (cuuint32_t)p.TH, (cuuint32_t)p.G};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, p.out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  cache[key] = tm;
  *out = tm;
  return true;
}

template <int ND, int S, int RS>
void launch_ws_k(const CUtensorMap& tm, const CUtensorMap& tmo, const CUtensorMap& tmr, const DwPwTcP& p, int B, int ntiles, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      cudaFuncSetAttribute(k_block_ws<ND, S, RS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      c = p.smem_bytes;
    }
  }
  int grid = std::min(ntiles, 148);
  if (grid < 1) grid = 1;
  launch_pdl(k_block_ws<ND, S, RS>, grid, (ND + kEpiWarps + 2 + kMmaWarps) * 32, p.smem_bytes, s, tm, tmo, tmr, p, B, ntiles);
}

template <int ND>
bool launch_ws_nd(const CUtensorMap& tm, const CUtensorMap& tmo, const CUtensorMap& tmr, const DwPwTcP& p, int B, int ntiles, cudaStream_t s) {
  const int S = p.has_dw ? p.s : 0;
  switch (S * 16 + (S ? p.RS : 1)) {
    case 1: launch_ws_k<ND, 0, 1>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 16 + 1: launch_ws_k<ND, 1, 1>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 16 + 2: launch_ws_k<ND, 1, 2>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 16 + 4: launch_ws_k<ND, 1, 4>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 32 + 1: launch_ws_k<ND, 2, 1>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 32 + 2: launch_ws_k<ND, 2, 2>(tm, tmo, tmr, p, B, ntiles, s); break;
    case 32 + 4: launch_ws_k<ND, 2, 4>(tm, tmo, tmr, p, B, ntiles, s); break;
    default: return false;      // no instantiation for this (stride, rows per item): the caller marks the engine failed
  }
  return true;
}

// Tensor map of the letterboxed u8x4 image: u32 [cap][H][W], box {40, PH, 1} (x start 16-byte aligned)
bool stem_tensor_map(const StemWsP& p, int cap, int PH, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  Key key(p.in8, cap, p.H, p.W, PH);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)cap};
  cuuint64_t gstr[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.W * p.H * 4};
  cuuint32_t box[3] = {40u, (cuuint32_t)PH, 1u};
  cuuint32_t estr[3] = {1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<uint8_t*>(p.in8), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  cache[key] = tm;
  *out = tm;
  return true;
}

template <int KW>
void launch_stem_ws_kw(const CUtensorMap& tm, const StemWsP& p, int B, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, size_t> cur;                       // device -> opted-in dynamic shared memory
  static std::map<std::pair<int, size_t>, int> occ;       // (device, smem) -> resident CTAs per SM
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 1;
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      cudaFuncSetAttribute(k_stem_ws<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      cudaFuncSetAttribute(k_stem_ws<KW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      c = p.smem_bytes;
    }
    auto key = std::make_pair(dev, p.smem_bytes);
    auto it = occ.find(key);
    if (it == occ.end()) {
      // Resident CTAs per SM from the resources themselves (the occupancy API answers 1 for this kernel, yet two CTAs
      // do run concurrently: 101 vs 124 us measured): shared memory, registers (allocated per 4 warps), TMEM columns.
      const int by_smem = (int)((227u * 1024u) / (p.smem_bytes + 1024u));
      const int by_regs = 65536 / (((kStemThreads / 32 + 3) / 4 * 4) * 32 * 56);
      const int by_tmem = 512 / 256;                 // kStemTmemCols
      int nb = std::max(1, std::min(std::min(by_smem, by_regs), by_tmem));
      it = occ.emplace(key, nb).first;
    }
    per_sm = it->second;
  }
  const int ntiles = ((p.OW + 15) / 16) * ((p.OH + 7) / 8) * B;
  int grid = std::min(ntiles, 148 * per_sm);
  if (grid < 1) grid = 1;
  launch_pdl(k_stem_ws<KW>, grid, kStemThreads, p.smem_bytes, s, tm, p, B, ntiles);
}

}  // namespace

bool launch_block_ws(const DwPwTcP& p0, int B, int cap, cudaStream_t s) {
  DwPwTcP p = p0;
  CUtensorMap tm, tmo;
  if (!input_tensor_map(p, cap, &tm)) return false;
  // the TMA-store epilogue needs a float4-aligned output tensor (the heads' dense views keep direct stores)
  if (p.no > 0 && !(p.vec_store && p.CoutS % 4 == 0 && p.out_istride % 4 == 0)) p.no = 0;
  if (p.c2 > 0 && (!p.vec_store || p.res_mode != 0 || p.no > 0 || p.act != kActNone)) return false;   // merged heads: float4-aligned first output, no residual
  if (p.no > 0) { if (!output_tensor_map(p, cap, &tmo)) return false; } else tmo = tm;
  // residual ring: float4-aligned residual tensor of the output's size (or exactly twice it when pooled); anything else loads directly
  CUtensorMap tmr = tm;
  if (p.nr > 0) {
    const int m = p.res_pool ? 2 : 1;
    const bool ok = p.res_mode == 2 && p.no == 0 && p.c2 == 0 && p.vec_store && p.res_Cs % 4 == 0 && p.res_istride % 4 == 0 && (size_t)p.res % 16 == 0 &&
                    p.res_H == m * p.OH && p.res_W == m * p.OW && p.res_C <= p.KSr;
    if (!ok || !res_tensor_map(p, cap, &tmr)) p.nr = 0;
  }
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  // FDT_WS_TRACE=1 (library built with FDT_NVCC_FLAGS=-DFDT_TRACE_BUILD): per-role timeline of CTA 0 (diagnostic;
  // synchronises after every launch)
#ifdef FDT_TRACE_BUILD
  static const bool trace = [] { const char* e = std::getenv("FDT_WS_TRACE"); return e && e[0] == '1'; }();
#else
  static const bool trace = false;
#endif
  static long long* d_trace = nullptr;
  if (trace) {
    if (!d_trace) cudaMalloc(&d_trace, 7 * 64 * 3 * sizeof(long long));
    cudaMemsetAsync(d_trace, 0, 7 * 64 * 3 * sizeof(long long), s);
    p.trace = d_trace;
  }
  if (!(p.nd == 12 ? launch_ws_nd<12>(tm, tmo, tmr, p, B, ntiles, s) : launch_ws_nd<8>(tm, tmo, tmr, p, B, ntiles, s))) return false;
  if (trace) {
    static long long h[7 * 64 * 3];
    cudaStreamSynchronize(s);
    cudaMemcpy(h, d_trace, sizeof h, cudaMemcpyDeviceToHost);
    const int n = std::min(64, (ntiles + 147) / 148);
    static const char* rn[7] = {"epi", "dw", "prod", "mma", "commit(a,d)", "epi(bufwait,math)", "epi(fence,arrive)"};
    fprintf(stderr, "WS_TRACE K8=%d Npad=%d S=%d RS=%d nd=%d ns=%d na=%d tiles/cta=%d:", p.K8, p.Npad, p.has_dw ? p.s : 0, p.RS, p.nd, p.ns, p.na, n);
    for (int r = 0; r < 7; ++r) {
      double wait = 0, work = 0; int cnt = 0;
      for (int i = 2; i < n - 1; ++i) {          // steady state: skip the first two and the last tile
        const long long* t = h + (r * 64 + i) * 3;
        if (!t[0] || !t[1] || !t[2]) continue;
        wait += (double)(t[1] - t[0]); work += (double)(t[2] - t[1]); ++cnt;
      }
      double period = 0;
      if (n > 4 && h[(r * 64 + n - 2) * 3] && h[(r * 64 + 2) * 3]) period = (double)(h[(r * 64 + n - 2) * 3] - h[(r * 64 + 2) * 3]) / (n - 4);
      if (cnt) fprintf(stderr, "  %s wait %.0f work %.0f period %.0f |", rn[r], wait / cnt, work / cnt, period);
    }
    fprintf(stderr, "\n");
  }
  return true;
}

}  // namespace fdt

namespace fdt {
bool launch_stem_ws(const StemWsP& p, int B, int cap, cudaStream_t s) {
  CUtensorMap tm;
  const int PH = 14 + p.kw;
  if (!stem_tensor_map(p, cap, PH, &tm)) return false;
  if (p.kw == 5) launch_stem_ws_kw<5>(tm, p, B, s);
  else launch_stem_ws_kw<3>(tm, p, B, s);
  return true;
}
}  // namespace fdt
