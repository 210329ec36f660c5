// k_dwpw_tc — BlazeBlock with the pointwise 1x1 convolution on the 5th-generation tensor cores.
//
//   stage input tile (+halo) with cp.async  ->  depthwise 3x3 on CUDA cores, result split into
//   TF32 hi + lo parts and written straight into the UMMA K-major core-matrix layout  ->
//   tcgen05.mma.kind::tf32 (M = 128 pixel slots, N = Cout padded to 16, K = Cin padded to 8; two MMAs
//   per K step: A_hi*W + A_lo*W, accumulator in TMEM)  ->  tcgen05.commit -> mbarrier  ->
//   epilogue: tcgen05.ld 32 lanes x 8 columns per warp, + bias + residual (from the staged tile,
//   optional 2x2 max-pool, zero channel pad) + ReLU/PReLU, float4 stores.
//
// The kernel is persistent and every tile of a layer has the same geometry, so all index
// arithmetic is done once per CTA: a staging table (one entry per 16-byte chunk of the input tile:
// global offset relative to the tile origin, shared-memory offset, tile-local coordinates for the
// border test) and a depthwise table (one entry per work item: window origin, first output slot,
// channel quad) live in shared memory, and each thread keeps its epilogue pixel in registers.  Per
// tile a chunk then costs ~12 instructions and the depthwise loop is pure LDS / FFMA / STS.
//
// Precision: the detector weights are fp16-origin, hence exact in TF32; the activation is split as
// a = hi + lo with hi = a & 0xFFFFE000 (exact) and lo = a - hi (exact in fp32, |lo| < 2^-10 |a|), so the
// only loss is the hardware's truncation of lo to TF32: <= 2^-21 relative per product, i.e. fp32-grade
// (measured 3e-7..9e-7 of max|ref| in tools/probe/tc_probe.cu vs 2e-7..3e-7 for an fp32 FMA chain).
#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <utility>

#include "kernels.h"

namespace fdt {
namespace {

constexpr int kTcThreads = 256;
constexpr uint32_t kLBO = 128;   // bytes between the two 16-byte K chunks of one MMA (adjacent core matrices)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16_u32(uint32_t smem_dst, const void* gmem_src, bool valid) {
  int sz = valid ? 16 : 0;   // src-size 0 => 16 bytes of zeros (TFLite SAME zero padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_dst), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

// SMEM matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor): start >> 4 in [0,14),
// LBO >> 4 in [16,30), SBO >> 4 in [32,46), version 1 in [46,48), layout type 0 in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((kLBO >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
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

__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
  a.x = fmaf(v.x, w.x, a.x); a.y = fmaf(v.y, w.y, a.y); a.z = fmaf(v.z, w.z, a.z); a.w = fmaf(v.w, w.w, a.w);
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

#ifndef FDT_TC_MINB
#define FDT_TC_MINB 1
#endif
__global__ void __launch_bounds__(kTcThreads, FDT_TC_MINB) k_dwpw_tc(DwPwTcP p, int B, int ntiles) {
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q8 = p.K8 >> 2;                       // 16-byte K chunks per row
  const uint32_t SBO = (uint32_t)Q8 * 128u;       // bytes between 8-row groups
  float* sB = smem;                               // [w_parts][Npad x K8] canonical (hi [, lo] parts of W)
  float* sBias = sB + (size_t)p.w_parts * p.Npad * p.K8;   // [Npad]
  float* sAlpha = sBias + p.Npad;                 // [Npad]
  float* sDw = sAlpha + p.Npad;                   // [10][K8]: 9 taps + bias
  float* sAhi = sDw + (p.has_dw ? 10 * p.K8 : 0); // [a_rows x K8] canonical
  float* sAlo = sAhi + (size_t)p.a_rows * p.K8;
  float* sIn = sAlo + (size_t)p.a_rows * p.K8;    // [G][IH][IW][KS]
  uint2* stab = reinterpret_cast<uint2*>(sIn + (size_t)p.nbuf * p.in_floats);   // staging table [n_chunks]
  uint2* dtab = stab + p.n_chunks;                              // depthwise table [n_items]
  const int thw = p.TH * p.TW;
  const int nslots = p.G * thw;

  // ---- prologue: weights, bias, tables, barrier, TMEM ------------------------------------------------
  const uint32_t sB_u32 = smem_u32(sB);
  for (int i = tid; i < p.w_parts * p.Npad * Q8; i += kTcThreads) cp_async16_u32(sB_u32 + 16u * i, p.wB + 4 * (size_t)i, true);
  for (int i = tid; i < p.Npad; i += kTcThreads) {
    sBias[i] = p.bias[i];
    sAlpha[i] = p.alpha ? p.alpha[i] : 0.f;
  }
  if (p.has_dw)
    for (int i = tid; i < 10 * p.K8; i += kTcThreads) sDw[i] = i < 9 * p.K8 ? p.dww[i] : p.dwb[i - 9 * p.K8];
  // staging table: chunk i = ((g*IH + ly)*IW + lx)*Q8 + qq
  for (int i = tid; i < p.n_chunks; i += kTcThreads) {
    int pix, qq, gy, lx, g, ly;
    p.fd_Q8.divmod(i, pix, qq);
    p.fd_IW.divmod(pix, gy, lx);
    p.fd_IH.divmod(gy, g, ly);
    int goff = 4 * qq < p.CinS ? (int)(g * p.in_istride + ((long long)ly * p.W + lx) * p.CinS + 4 * qq) : -1;
    uint32_t soff16 = (uint32_t)(((gy * p.IW + lx) * p.KS + 4 * qq) >> 2);
    stab[i] = make_uint2((uint32_t)goff, soff16 | ((uint32_t)ly << 14) | ((uint32_t)lx << 20) | ((uint32_t)g << 26));
  }
  // depthwise table: item = (((g*Q8 + qq)*nstrips + st)*TW + tx)  (tx fastest => conflict-free LDS)
  if (p.has_dw) {
    for (int it = tid; it < p.n_items; it += kTcThreads) {
      int tx, r, st, r2, qq, g;
      p.fd_TW.divmod(it, r, tx);
      p.fd_nstrips.divmod(r, r2, st);
      p.fd_Q8.divmod(r2, g, qq);
      const int tyb = st * p.RS;
      uint32_t in_off = (uint32_t)(((g * p.IH + tyb * p.s) * p.IW + tx * p.s) * p.KS + 4 * qq);
      uint32_t slot0 = (uint32_t)(g * thw + tyb * p.TW + tx);
      dtab[it] = make_uint2(in_off, slot0 | ((uint32_t)qq << 8));
    }
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  // epilogue pixel of this thread (fixed for the whole kernel)
  const int lq = warp & 3, half = warp >> 2;
  const int slot = lq * 32 + lane;
  int e_g, e_r, e_ty, e_tx;
  p.fd_thw.divmod(slot, e_g, e_r);
  p.fd_TW.divmod(e_r, e_ty, e_tx);
  const bool slot_ok = slot < nslots;
  const long long o_rel = slot_ok ? (long long)e_g * p.out_istride + ((long long)e_ty * p.OW + e_tx) * p.CoutS : 0;
  const int rs = p.res_pool ? 2 : 1;
  const uint32_t res_off = slot_ok ? (uint32_t)((((size_t)e_g * p.IH + e_ty * rs + p.dpt) * p.IW + e_tx * rs + p.dpl) * p.KS) : 0u;
  const uint32_t row_f = (uint32_t)p.IW * p.KS;          // floats per staged row
  float* const sInA = sIn;
  float* const sInB = p.nbuf > 1 ? sIn + p.in_floats : sIn;

  cp_async_wait_all();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0;

  // Stages the input tile (+halo) of `t` into `dst`: table-driven cp.async, zero fill outside the image.
  auto stage = [&](int t, float* dst) {
    int grp, trem, tyi, txi;
    p.fd_tpg.divmod(t, grp, trem);
    p.fd_tilesX.divmod(trem, tyi, txi);
    const int b0 = grp * p.G;
    const int iy0 = tyi * p.TH * p.s - p.dpt, ix0 = txi * p.TW * p.s - p.dpl;
    const float* tbase = p.in + (long long)b0 * p.in_istride + ((long long)iy0 * p.W + ix0) * p.CinS;
    const unsigned uH = (unsigned)p.H, uW = (unsigned)p.W;
    const int nb = B - b0;
    const uint32_t dst_u32 = smem_u32(dst);
    for (int i = tid; i < p.n_chunks; i += kTcThreads) {
      const uint2 e = stab[i];
      const int goff = (int)e.x;
      const unsigned y = (unsigned)(iy0 + (int)((e.y >> 14) & 63u));
      const unsigned x = (unsigned)(ix0 + (int)((e.y >> 20) & 63u));
      const bool ok = goff >= 0 && y < uH && x < uW && (int)(e.y >> 26) < nb;
      cp_async16_u32(dst_u32 + ((e.y & 0x3FFFu) << 4), ok ? tbase + goff : p.in, ok);
    }
  };
  int cur = 0;
  if ((int)blockIdx.x < ntiles) stage(blockIdx.x, sInA);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int grp, trem, tyi, txi;
    p.fd_tpg.divmod(tile, grp, trem);
    p.fd_tilesX.divmod(trem, tyi, txi);
    const int ty0 = tyi * p.TH, tx0 = txi * p.TW;
    const int b0 = grp * p.G;
    sIn = cur ? sInB : sInA;
    const float* res_s = sIn + res_off;
    cp_async_wait_all();          // this tile's input (issued one iteration ago, or in the prologue)
    __syncthreads();
    // software pipeline: the next tile's input streams into the other buffer while this tile is computed
    if (p.nbuf > 1 && tile + (int)gridDim.x < ntiles) stage(tile + gridDim.x, cur ? sInA : sInB);
    // ---- A operand: depthwise 3x3 (or plain copy) -> hi/lo split -> canonical layout
    if (p.has_dw) {
      for (int it = tid; it < p.n_items; it += kTcThreads) {
        const uint2 e = dtab[it];
        const uint32_t qq = e.y >> 8;
        uint32_t sl = e.y & 0xFFu;
        const float* wq = sDw + 4 * qq;
        float4 w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = ld4(wq + t * p.K8);
        const float4 bias = ld4(wq + 9 * p.K8);
        const float* base = sIn + e.x;
        const uint32_t aq = qq * kLBO;
        if (p.s == 1) {
          float4 r0[3], r1[3], rr[3];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            r0[kx] = ld4(base + kx * p.KS);
            r1[kx] = ld4(base + row_f + kx * p.KS);
          }
          const float* nrow = base + 2 * row_f;
          for (int t = 0; t < p.RS; ++t) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) rr[kx] = ld4(nrow + kx * p.KS);
            float4 a = bias;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) fma4(a, r0[kx], w[kx]);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) fma4(a, r1[kx], w[3 + kx]);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) fma4(a, rr[kx], w[6 + kx]);
            split_store(sAhi, sAlo, ((sl >> 3) * SBO + aq + (sl & 7u) * 16u) >> 2, a);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) { r0[kx] = r1[kx]; r1[kx] = rr[kx]; }
            nrow += row_f;
            sl += p.TW;
          }
        } else {
          const float* row = base;
          for (int t = 0; t < p.RS; ++t) {
            float4 a = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) fma4(a, ld4(row + ky * row_f + kx * p.KS), w[ky * 3 + kx]);
            }
            split_store(sAhi, sAlo, ((sl >> 3) * SBO + aq + (sl & 7u) * 16u) >> 2, a);
            row += 2 * row_f;
            sl += p.TW;
          }
        }
      }
    } else {
      // pointwise only: slot s <-> staged pixel s (IH = TH, IW = TW); slot fastest
      for (int it = tid; it < nslots * Q8; it += kTcThreads) {
        int qq, sl;
        p.fd_nslots.divmod(it, qq, sl);
        const float4 a = ld4(sIn + (size_t)sl * p.KS + 4 * qq);
        split_store(sAhi, sAlo, (((uint32_t)sl >> 3) * SBO + (uint32_t)qq * kLBO + ((uint32_t)sl & 7u) * 16u) >> 2, a);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- tensor-core GEMM: D[128 x Npad] (TMEM) = (A_hi + A_lo)[128 x K8] * W[Npad x K8]^T
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo);
      const int ksteps = p.K8 >> 3;
      const uint32_t b_lo = sB_u32 + (uint32_t)p.Npad * p.K8 * 4u;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t db = make_desc(sB_u32 + ks * 2 * kLBO, SBO);
        const uint64_t dah = make_desc(a_hi + ks * 2 * kLBO, SBO);
        mma_tf32(tmem_base, dah, db, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(tmem_base, make_desc(a_lo + ks * 2 * kLBO, SBO), db, idesc, 1u);
        // fp32 weights (face_landmark): W = W_hi + W_lo, third product A_hi * W_lo (A_lo * W_lo ~ 2^-22, dropped)
        if (p.w_parts > 1) mma_tf32(tmem_base, dah, make_desc(b_lo + ks * 2 * kLBO, SBO), idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
      }
      parity ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- epilogue: warp w owns TMEM lanes 32*(w%4).., and the column half w/4
    {
      const int ncol = p.Npad >> 1;
      const int oy = ty0 + e_ty, ox = tx0 + e_tx, b = b0 + e_g;
      const bool valid = slot_ok && b < B && oy < p.OH && ox < p.OW;
      float* orow = p.out + (long long)b0 * p.out_istride + ((long long)ty0 * p.OW + tx0) * p.CoutS + o_rel;
      const float* rbase = p.res_mode == 2 ? p.res + (size_t)(valid ? b : 0) * p.res_istride : nullptr;
      for (int cc = 0; cc < ncol; cc += 8) {
        const int c = half * ncol + cc;
        uint32_t u[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!valid || c >= p.CoutS) continue;
        const float4 b0v = ld4(sBias + c), b1v = ld4(sBias + c + 4);
        float4 v0 = make_float4(__uint_as_float(u[0]) + b0v.x, __uint_as_float(u[1]) + b0v.y, __uint_as_float(u[2]) + b0v.z, __uint_as_float(u[3]) + b0v.w);
        float4 v1 = make_float4(__uint_as_float(u[4]) + b1v.x, __uint_as_float(u[5]) + b1v.y, __uint_as_float(u[6]) + b1v.z, __uint_as_float(u[7]) + b1v.w);
        if (p.res_mode == 1) {
          // residual straight from the staged tile (channels >= Cin are the zero channel pad)
          if (p.res_pool) {
            const float* r1p = res_s + p.KS;
            const float* r2p = res_s + row_f;
            const float* r3p = r2p + p.KS;
            if (c < p.res_lim) add4(v0, max4(max4(ld4(res_s + c), ld4(r1p + c)), max4(ld4(r2p + c), ld4(r3p + c))));
            if (c + 4 < p.res_lim) add4(v1, max4(max4(ld4(res_s + c + 4), ld4(r1p + c + 4)), max4(ld4(r2p + c + 4), ld4(r3p + c + 4))));
          } else {
            if (c < p.res_lim) add4(v0, ld4(res_s + c));
            if (c + 4 < p.res_lim) add4(v1, ld4(res_s + c + 4));
          }
        } else if (p.res_mode == 2) {
          float rv[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ch = c + j;
            rv[j] = 0.f;
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
              rv[j] = m;
            } else {
              rv[j] = rbase[((size_t)oy * p.res_W + ox) * p.res_Cs + ch];
            }
          }
          add4(v0, make_float4(rv[0], rv[1], rv[2], rv[3]));
          add4(v1, make_float4(rv[4], rv[5], rv[6], rv[7]));
        }
        if (p.act == kActRelu) {
          v0 = max4(v0, make_float4(0.f, 0.f, 0.f, 0.f));
          v1 = max4(v1, make_float4(0.f, 0.f, 0.f, 0.f));
        } else if (p.act == kActPrelu) {
          const float4 a0 = ld4(sAlpha + c), a1 = ld4(sAlpha + c + 4);
          v0.x = v0.x >= 0.f ? v0.x : v0.x * a0.x; v0.y = v0.y >= 0.f ? v0.y : v0.y * a0.y;
          v0.z = v0.z >= 0.f ? v0.z : v0.z * a0.z; v0.w = v0.w >= 0.f ? v0.w : v0.w * a0.w;
          v1.x = v1.x >= 0.f ? v1.x : v1.x * a1.x; v1.y = v1.y >= 0.f ? v1.y : v1.y * a1.y;
          v1.z = v1.z >= 0.f ? v1.z : v1.z * a1.z; v1.w = v1.w >= 0.f ? v1.w : v1.w * a1.w;
        }
        if (p.vec_store) {
          // lanes >= Cout need no masking: their weights and bias are zero and the residual is the zero
          // channel pad there, so they come out as exact zeros
          *reinterpret_cast<float4*>(orow + c) = v0;
          if (c + 4 < p.CoutS) *reinterpret_cast<float4*>(orow + c + 4) = v1;
        } else {
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < p.Cout) orow[c + j] = vv[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // TMEM, sIn and the A tiles are free again
    if (p.nbuf > 1) cur ^= 1;
    else if (tile + (int)gridDim.x < ntiles) stage(tile + gridDim.x, sInA);
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}


// im2col gather of one half of the K chunks of pixel `prow` (patch origin of the pixel's receptive field)
template <int KW, int HALF>
__device__ __forceinline__ void stem_gather(const float* prow, float* sAhi, float* sAlo, uint32_t a_row) {
  constexpr int PW3 = ((16 - 1) * 2 + KW) * 3, SEG = KW * 3, K = KW * SEG, K8 = (K + 7) / 8 * 8, Q8 = K8 / 4;
#pragma unroll
  for (int kk = 0; kk < Q8 / 2; ++kk) {
    constexpr int dummy = 0; (void)dummy;
    const int kq = HALF * (Q8 / 2) + kk;
    float e[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int k = 4 * kq + j;                       // compile-time after unrolling
      e[j] = k < K ? prow[(k / SEG) * PW3 + (k % SEG)] : 0.f;
    }
    split_store(sAhi, sAlo, a_row + (uint32_t)kq * (kLBO >> 2), make_float4(e[0], e[1], e[2], e[3]));
  }
}

// ------------------------------------------------------------------------------------------------
// k_stem_tc — the KW x KW / stride-2 input convolution (3 channels) as an im2col GEMM on tcgen05:
// D[128 px x Npad] = A[128 x K8] * W[Npad x K8]^T with K = KW*KW*3 (75 -> 80 for the 5x5 stem).
// The u8x4 BGRX patch of a tile of 8 x 16 output pixels is normalised (fma(v, 1/127.5, -1), BGR->RGB) into
// shared memory as dense RGB float triplets; an im2col row is then KW contiguous segments of KW*3
// floats, gathered, split into TF32 hi + lo and written in the UMMA core-matrix layout.
template <int KW>
__global__ void __launch_bounds__(kTcThreads, 2) k_stem_tc(StemTcP p, int B, int ntiles) {
  constexpr int TH = 8, TW = 16;
  constexpr int PH = (TH - 1) * 2 + KW, PW = (TW - 1) * 2 + KW, PW3 = PW * 3;
  constexpr int SEG = KW * 3, K = KW * SEG, K8 = (K + 7) / 8 * 8, Q8 = K8 / 4;
  constexpr uint32_t SBO = (uint32_t)Q8 * 128u;
  extern __shared__ __align__(128) float smem[];
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sB = smem;                               // [w_parts][Npad x K8] canonical
  float* sBias = sB + (size_t)p.w_parts * p.Npad * K8;   // [Npad]
  float* sAlpha = sBias + p.Npad;
  float* sAhi = sAlpha + p.Npad;                  // [128 x K8] canonical
  float* sAlo = sAhi + 128 * K8;
  float* sP = sAlo + 128 * K8;                    // [PH][PW*3] normalised RGB patch
  uint32_t* sRaw = reinterpret_cast<uint32_t*>(sP + PH * PW3);   // [2][PH*PW] raw BGRX pixels (cp.async, double-buffered)

  const uint32_t sB_u32 = smem_u32(sB);
  for (int i = tid; i < p.w_parts * p.Npad * Q8; i += kTcThreads) cp_async16_u32(sB_u32 + 16u * i, p.wB + 4 * (size_t)i, true);
  for (int i = tid; i < p.Npad; i += kTcThreads) {
    sBias[i] = p.bias[i];
    sAlpha[i] = p.alpha ? p.alpha[i] : 0.f;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"((uint32_t)p.tmem_cols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  cp_async_wait_all();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
  uint32_t parity = 0;
  const int tilesX = (p.OW + TW - 1) / TW, tilesY = (p.OH + TH - 1) / TH;
  const int tiles_per_img = tilesX * tilesY;
  // this thread's pixel for the im2col gather (slot r) and for the epilogue (slot e)
  const int r = tid & 127, khalf = tid >> 7;
  const int r_ty = r >> 4, r_tx = r & 15;
  const float* prow = sP + (2 * r_ty) * PW3 + 6 * r_tx;
  const uint32_t a_row = (((uint32_t)r >> 3) * SBO + ((uint32_t)r & 7u) * 16u) >> 2;
  const int lq = warp & 3, half = warp >> 2;
  const int eslot = lq * 32 + lane;
  const int e_ty = eslot >> 4, e_tx = eslot & 15;

  // raw BGRX patch of tile `t` -> sRaw[buf] with 4-byte cp.async (zero fill outside the image = SAME padding;
  // a zero pixel must become 0.0, not -1.0, so validity is re-derived when converting)
  auto stage_raw = [&](int t, int buf) {
    const int b = t / tiles_per_img;
    const int trem = t - b * tiles_per_img;
    const int iy0 = (trem / tilesX) * TH * 2 - p.pt, ix0 = (trem % tilesX) * TW * 2 - p.pl;
    const uint32_t* img = reinterpret_cast<const uint32_t*>(p.in8) + (size_t)b * p.H * p.W;
    const uint32_t dst = smem_u32(sRaw + buf * (PH * PW));
    for (int i = tid; i < PH * PW; i += kTcThreads) {
      const int ly = i / PW, lx = i - ly * PW;
      const int y = iy0 + ly, x = ix0 + lx;
      const bool ok = (unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W;
      const int sz = ok ? 4 : 0;
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;\n" ::"r"(dst + 4u * i), "l"(ok ? img + (size_t)y * p.W + x : img), "r"(sz));
    }
  };
  int cur = 0;
  if ((int)blockIdx.x < ntiles) stage_raw(blockIdx.x, 0);

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int trem = tile - b * tiles_per_img;
    const int ty0 = (trem / tilesX) * TH, tx0 = (trem % tilesX) * TW;
    const int iy0 = ty0 * 2 - p.pt, ix0 = tx0 * 2 - p.pl;
    cp_async_wait_all();
    __syncthreads();
    if (tile + (int)gridDim.x < ntiles) stage_raw(tile + gridDim.x, cur ^ 1);   // prefetch the next patch
    // ---- patch: u8x4 BGRX -> normalised RGB floats (bgrMatToSignedFloat32, helpers.dart:401-406)
    {
      const uint32_t* raw = sRaw + cur * (PH * PW);
      for (int i = tid; i < PH * PW; i += kTcThreads) {
        const int ly = i / PW, lx = i - ly * PW;
        const int y = iy0 + ly, x = ix0 + lx;
        float3 v = make_float3(0.f, 0.f, 0.f);
        if ((unsigned)y < (unsigned)p.H && (unsigned)x < (unsigned)p.W) {
          const uint32_t u = raw[i];
          v.x = fmaf((float)((u >> 16) & 0xFFu), 1.0f / 127.5f, -1.0f);
          v.y = fmaf((float)((u >> 8) & 0xFFu), 1.0f / 127.5f, -1.0f);
          v.z = fmaf((float)(u & 0xFFu), 1.0f / 127.5f, -1.0f);
        }
        float* d = sP + ly * PW3 + lx * 3;
        d[0] = v.x; d[1] = v.y; d[2] = v.z;
      }
    }
    cur ^= 1;
    __syncthreads();
    // ---- im2col gather -> hi/lo split -> canonical A (thread = pixel r, one half of the K chunks)
    if (khalf == 0) stem_gather<KW, 0>(prow, sAhi, sAlo, a_row);     // khalf is warp-uniform
    else stem_gather<KW, 1>(prow, sAhi, sAlo, a_row);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo);
#pragma unroll 1
      const uint32_t b_lo = sB_u32 + (uint32_t)p.Npad * K8 * 4u;
      for (int ks = 0; ks < K8 / 8; ++ks) {
        const uint64_t db = make_desc(sB_u32 + ks * 2 * kLBO, SBO);
        const uint64_t dah = make_desc(a_hi + ks * 2 * kLBO, SBO);
        mma_tf32(tmem_base, dah, db, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(tmem_base, make_desc(a_lo + ks * 2 * kLBO, SBO), db, idesc, 1u);
        if (p.w_parts > 1) mma_tf32(tmem_base, dah, make_desc(b_lo + ks * 2 * kLBO, SBO), idesc, 1u);
      }
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {
      uint32_t ok = 0;
      while (!ok) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(parity) : "memory");
      }
      parity ^= 1u;
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    {
      const int ncol = p.Npad >> 1;
      const int oy = ty0 + e_ty, ox = tx0 + e_tx;
      const bool valid = oy < p.OH && ox < p.OW;
      float* orow = p.out + (size_t)b * p.out_istride + ((size_t)(valid ? oy : 0) * p.OW + (valid ? ox : 0)) * p.CoutS;
      for (int cc = 0; cc < ncol; cc += 8) {
        const int c = half * ncol + cc;
        uint32_t u[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!valid || c >= p.CoutS) continue;
        const float4 b0v = ld4(sBias + c), b1v = ld4(sBias + c + 4);
        float4 v0 = make_float4(__uint_as_float(u[0]) + b0v.x, __uint_as_float(u[1]) + b0v.y, __uint_as_float(u[2]) + b0v.z, __uint_as_float(u[3]) + b0v.w);
        float4 v1 = make_float4(__uint_as_float(u[4]) + b1v.x, __uint_as_float(u[5]) + b1v.y, __uint_as_float(u[6]) + b1v.z, __uint_as_float(u[7]) + b1v.w);
        if (p.act == kActRelu) {
          v0 = max4(v0, make_float4(0.f, 0.f, 0.f, 0.f));
          v1 = max4(v1, make_float4(0.f, 0.f, 0.f, 0.f));
        } else if (p.act == kActPrelu) {
          const float4 a0 = ld4(sAlpha + c), a1 = ld4(sAlpha + c + 4);
          v0.x = v0.x >= 0.f ? v0.x : v0.x * a0.x; v0.y = v0.y >= 0.f ? v0.y : v0.y * a0.y;
          v0.z = v0.z >= 0.f ? v0.z : v0.z * a0.z; v0.w = v0.w >= 0.f ? v0.w : v0.w * a0.w;
          v1.x = v1.x >= 0.f ? v1.x : v1.x * a1.x; v1.y = v1.y >= 0.f ? v1.y : v1.y * a1.y;
          v1.z = v1.z >= 0.f ? v1.z : v1.z * a1.z; v1.w = v1.w >= 0.f ? v1.w : v1.w * a1.w;
        }
        if (p.vec_store) {
          *reinterpret_cast<float4*>(orow + c) = v0;
          if (c + 4 < p.CoutS) *reinterpret_cast<float4*>(orow + c + 4) = v1;
        } else {
          const float vv[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < p.Cout) orow[c + j] = vv[j];
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

}  // namespace

void launch_dwpw_tc(const DwPwTcP& p, int B, cudaStream_t s, int max_ctas) {
  // grid = resident CTAs: 148 SMs x the occupancy the kernel really gets at this shared-memory size
  static std::mutex mu;
  static std::map<std::pair<int, size_t>, int> occ;
  static std::map<int, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 1;
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      cudaFuncSetAttribute(k_dwpw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      c = p.smem_bytes;
    }
    auto key = std::make_pair(dev, p.smem_bytes);
    auto it = occ.find(key);
    if (it == occ.end()) {
      int nb = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_dwpw_tc, kTcThreads, p.smem_bytes) != cudaSuccess || nb < 1) nb = 1;
      it = occ.emplace(key, nb).first;
    }
    per_sm = it->second;
  }
  (void)max_ctas;
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  // Grid = resident CTAs, doubled for layers with many tiles (measured: a second wave de-synchronises the
  // CTAs' load / compute phases; for small layers the smaller grid leaves room for the other stream's kernel).
  static const int force = [] { const char* e = std::getenv("FDT_TC_GRIDMULT"); return e ? std::atoi(e) : 0; }();
  const int resident = 148 * per_sm;
  const int mult = force > 0 ? force : (ntiles >= 6 * resident ? 2 : 1);
  int grid = std::min(ntiles, resident * mult);
  if (grid < 1) grid = 1;
  k_dwpw_tc<<<grid, kTcThreads, p.smem_bytes, s>>>(p, B, ntiles);
}

template <int KW>
static void launch_stem_tc_kw(const StemTcP& p, int B, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, int> occ;     // device -> resident CTAs per SM
  int dev = 0;
  cudaGetDevice(&dev);
  int per_sm = 1;
  {
    std::lock_guard<std::mutex> g(mu);
    auto it = occ.find(dev);
    if (it == occ.end()) {
      cudaFuncSetAttribute(k_stem_tc<KW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      int nb = 1;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_stem_tc<KW>, kTcThreads, p.smem_bytes) != cudaSuccess || nb < 1) nb = 1;
      it = occ.emplace(dev, nb).first;
    }
    per_sm = it->second;
  }
  const int tiles = ((p.OW + 15) / 16) * ((p.OH + 7) / 8);
  const int ntiles = tiles * B;
  int grid = std::min(ntiles, 148 * per_sm * 2);
  if (grid < 1) grid = 1;
  k_stem_tc<KW><<<grid, kTcThreads, p.smem_bytes, s>>>(p, B, ntiles);
}

void launch_stem_tc(const StemTcP& p, int B, cudaStream_t s) {
  if (p.kw == 5) launch_stem_tc_kw<5>(p, B, s);
  else launch_stem_tc_kw<3>(p, B, s);
}

}  // namespace fdt
