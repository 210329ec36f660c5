// k_dwpw_tc — BlazeBlock with the pointwise 1x1 convolution on the 5th-generation tensor cores.
//
//   stage input tile (+halo) with cp.async  ->  depthwise 3x3 on CUDA cores, result split into
//   TF32 hi + lo parts and written straight into the UMMA K-major core-matrix layout  ->
//   tcgen05.mma.kind::tf32 (M = 128 pixel slots, N = Cout padded to 16, K = Cin padded to 8; two MMAs
//   per K step: A_hi*W + A_lo*W, accumulator in TMEM)  ->  tcgen05.commit -> mbarrier  ->
//   epilogue: tcgen05.ld 32 lanes x 8 columns per warp, + bias + residual (from the staged tile,
//   optional 2x2 max-pool, zero channel pad) + ReLU/PReLU, float4 stores.
//
// Precision: the detector weights are fp16-origin, hence exact in TF32; the activation is split as
// a = hi + lo with hi = a & 0xFFFFE000 (exact) and lo = a - hi (exact in fp32, |lo| < 2^-10 |a|), so the
// only loss is the hardware's truncation of lo to TF32: <= 2^-21 relative per product, i.e. fp32-grade
// (measured 3e-7..9e-7 of max|ref| in tools/probe/tc_probe.cu vs 2e-7..3e-7 for an fp32 FMA chain).
#include <map>
#include <mutex>
#include <utility>

#include "kernels.h"

namespace fdt {
namespace {

constexpr int kTcThreads = 256;
constexpr uint32_t kLBO = 128;   // bytes between the two 16-byte K chunks of one MMA (adjacent core matrices)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float act1(float v, int act, float alpha) {
  if (act == kActRelu) return fmaxf(v, 0.f);
  if (act == kActPrelu) return v >= 0.f ? v : v * alpha;
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  int sz = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(smem_u32(smem_dst)), "l"(gmem_src), "r"(sz));
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

__device__ __forceinline__ void split_store(float* sAhi, float* sAlo, size_t off, float4 a) {
  float4 hi, lo;
  hi.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); lo.x = a.x - hi.x;
  hi.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u); lo.y = a.y - hi.y;
  hi.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); lo.z = a.z - hi.z;
  hi.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u); lo.w = a.w - hi.w;
  *reinterpret_cast<float4*>(sAhi + off) = hi;
  *reinterpret_cast<float4*>(sAlo + off) = lo;
}

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
  const int in_elems = p.G * p.IH * p.IW * p.KS;
  float* sB = smem;                               // [Npad x K8] canonical
  float* sBias = sB + (size_t)p.Npad * p.K8;      // [Npad]
  float* sAlpha = sBias + p.Npad;                 // [Npad]
  float* sAhi = sAlpha + p.Npad;                  // [a_rows x K8] canonical
  float* sAlo = sAhi + (size_t)p.a_rows * p.K8;
  float* sIn = sAlo + (size_t)p.a_rows * p.K8;    // [G][IH][IW][KS]
  const int thw = p.TH * p.TW;
  const int nslots = p.G * thw;

  // ---- prologue: weights (already in canonical layout in global memory), bias, barrier, TMEM
  for (int i = tid; i < p.Npad * Q8; i += kTcThreads) cp_async16(sB + 4 * (size_t)i, p.wB + 4 * (size_t)i, true);
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

  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int grp, trem, tyi, txi;
    p.fd_tpg.divmod(tile, grp, trem);
    p.fd_tilesX.divmod(trem, tyi, txi);
    const int ty0 = tyi * p.TH, tx0 = txi * p.TW;
    const int b0 = grp * p.G;
    const int iy0 = ty0 * p.s - p.dpt, ix0 = tx0 * p.s - p.dpl;
    // ---- stage the input tile (+halo); zero outside the image and in the K padding lanes.
    //      TPR threads share one tile row: the row's 64-bit base and the y / batch bounds are computed
    //      once, each 16-byte chunk then costs one FastDiv, an x bound check and two adds.
    {
      const int nrows = p.G * p.IH;
      const int lane_r = tid & (p.TPR - 1);
      const int chunks = p.IW * Q8;
      for (int r = tid >> p.TPR_log2; r < nrows; r += (kTcThreads >> p.TPR_log2)) {
        int g, ly;
        p.fd_IH.divmod(r, g, ly);
        const int b = b0 + g, y = iy0 + ly;
        const bool row_ok = b < B && y >= 0 && y < p.H;
        const float* grow = p.in + (size_t)(row_ok ? b : 0) * p.in_istride + ((long long)(row_ok ? y : 0) * p.W + ix0) * p.CinS;
        float* srow = sIn + (size_t)r * p.IW * p.KS;
        for (int c = lane_r; c < chunks; c += p.TPR) {
          int lx, qq;
          p.fd_Q8.divmod(c, lx, qq);
          const int x = ix0 + lx;
          const bool ok = row_ok && x >= 0 && x < p.W && 4 * qq < p.CinS;
          cp_async16(srow + lx * p.KS + 4 * qq, ok ? grow + lx * p.CinS + 4 * qq : p.in, ok);
        }
      }
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- A operand: depthwise 3x3 (or plain copy) -> hi/lo split -> canonical layout
    if (p.has_dw) {
      // item = (g, qq, row strip of RS outputs, tx); tx fastest => conflict-free smem reads
      const int nstrips = p.TH / p.RS;
      const int nitems = p.G * Q8 * nstrips * p.TW;
      for (int it = tid; it < nitems; it += kTcThreads) {
        int tx, r, st, r2, qq, g;
        p.fd_TW.divmod(it, r, tx);
        p.fd_nstrips.divmod(r, r2, st);
        p.fd_Q8.divmod(r2, g, qq);
        float4 w[9];
#pragma unroll
        for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const float4*>(p.dww + (size_t)t * p.K8 + 4 * qq);
        const float4 bias = *reinterpret_cast<const float4*>(p.dwb + 4 * qq);
        const size_t rstride = (size_t)p.IW * p.KS;
        const int tyb = st * p.RS;
        const float* base = sIn + ((size_t)g * p.IH * p.IW + (size_t)tyb * p.s * p.IW + (size_t)tx * p.s) * p.KS + 4 * qq;
        if (p.s == 1) {
          float4 r0[3], r1[3], rr[3];
#pragma unroll
          for (int kx = 0; kx < 3; ++kx) {
            r0[kx] = *reinterpret_cast<const float4*>(base + (size_t)kx * p.KS);
            r1[kx] = *reinterpret_cast<const float4*>(base + rstride + (size_t)kx * p.KS);
          }
          for (int t = 0; t < p.RS; ++t) {
#pragma unroll
            for (int kx = 0; kx < 3; ++kx)
              rr[kx] = *reinterpret_cast<const float4*>(base + (size_t)(t + 2) * rstride + (size_t)kx * p.KS);
            float4 a = bias;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              a.x = fmaf(r0[kx].x, w[kx].x, a.x); a.y = fmaf(r0[kx].y, w[kx].y, a.y);
              a.z = fmaf(r0[kx].z, w[kx].z, a.z); a.w = fmaf(r0[kx].w, w[kx].w, a.w);
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              a.x = fmaf(r1[kx].x, w[3 + kx].x, a.x); a.y = fmaf(r1[kx].y, w[3 + kx].y, a.y);
              a.z = fmaf(r1[kx].z, w[3 + kx].z, a.z); a.w = fmaf(r1[kx].w, w[3 + kx].w, a.w);
            }
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              a.x = fmaf(rr[kx].x, w[6 + kx].x, a.x); a.y = fmaf(rr[kx].y, w[6 + kx].y, a.y);
              a.z = fmaf(rr[kx].z, w[6 + kx].z, a.z); a.w = fmaf(rr[kx].w, w[6 + kx].w, a.w);
            }
            int slot = g * thw + (tyb + t) * p.TW + tx;
            split_store(sAhi, sAlo, ((size_t)(slot >> 3) * SBO + (size_t)qq * kLBO + (size_t)(slot & 7) * 16) >> 2, a);
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) { r0[kx] = r1[kx]; r1[kx] = rr[kx]; }
          }
        } else {
          for (int t = 0; t < p.RS; ++t) {
            float4 a = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const float* row = base + (size_t)(t * p.s + ky) * rstride;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float4 v = *reinterpret_cast<const float4*>(row + (size_t)kx * p.KS);
                const float4 ww = w[ky * 3 + kx];
                a.x = fmaf(v.x, ww.x, a.x); a.y = fmaf(v.y, ww.y, a.y);
                a.z = fmaf(v.z, ww.z, a.z); a.w = fmaf(v.w, ww.w, a.w);
              }
            }
            int slot = g * thw + (tyb + t) * p.TW + tx;
            split_store(sAhi, sAlo, ((size_t)(slot >> 3) * SBO + (size_t)qq * kLBO + (size_t)(slot & 7) * 16) >> 2, a);
          }
        }
      }
    } else {
      // pointwise only: slot s <-> staged pixel s (IH = TH, IW = TW)
      for (int it = tid; it < nslots * Q8; it += kTcThreads) {
        int qq, slot;
        p.fd_nslots.divmod(it, qq, slot);      // slot fastest
        const float4 a = *reinterpret_cast<const float4*>(sIn + (size_t)slot * p.KS + 4 * qq);
        split_store(sAhi, sAlo, ((size_t)(slot >> 3) * SBO + (size_t)qq * kLBO + (size_t)(slot & 7) * 16) >> 2, a);
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    // ---- tensor-core GEMM: D[128 x Npad] (TMEM) = (A_hi + A_lo)[128 x K8] * W[Npad x K8]^T
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), bb = smem_u32(sB);
      const int ksteps = p.K8 >> 3;
      for (int ks = 0; ks < ksteps; ++ks) {
        const uint64_t db = make_desc(bb + ks * 2 * kLBO, SBO);
        mma_tf32(tmem_base, make_desc(a_hi + ks * 2 * kLBO, SBO), db, idesc, ks > 0 ? 1u : 0u);
        mma_tf32(tmem_base, make_desc(a_lo + ks * 2 * kLBO, SBO), db, idesc, 1u);
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
      const int lq = warp & 3, half = warp >> 2;
      const int ncol = p.Npad >> 1;
      const int slot = lq * 32 + lane;
      int g, r, ty, tx;
      p.fd_thw.divmod(slot, g, r);
      p.fd_TW.divmod(r, ty, tx);
      const int oy = ty0 + ty, ox = tx0 + tx, b = b0 + g;
      const bool valid = slot < nslots && b < B && oy < p.OH && ox < p.OW;
      float* orow = p.out + (size_t)(valid ? b : 0) * p.out_istride + ((size_t)(valid ? oy : 0) * p.OW + (valid ? ox : 0)) * p.CoutS;
      const float* rbase = p.res_mode == 2 ? p.res + (size_t)(valid ? b : 0) * p.res_istride : nullptr;
      for (int cc = 0; cc < ncol; cc += 8) {
        const int c = half * ncol + cc;
        uint32_t u[8];
        const uint32_t taddr = tmem_base + ((uint32_t)(lq * 32) << 16) + (uint32_t)c;
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!valid || c >= p.CoutS) continue;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(u[j]) + sBias[c + j];
        if (p.res_mode == 1) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = c + 4 * h;
            if (ch >= p.res_lim) continue;      // channels >= Cin: zero channel pad of the residual
            float4 rv;
            if (p.res_pool) {
              const float* rb = sIn + (((size_t)g * p.IH + 2 * ty + p.dpt) * p.IW + 2 * tx + p.dpl) * p.KS + ch;
              const float4 m0 = *reinterpret_cast<const float4*>(rb);
              const float4 m1 = *reinterpret_cast<const float4*>(rb + p.KS);
              const float4 m2 = *reinterpret_cast<const float4*>(rb + (size_t)p.IW * p.KS);
              const float4 m3 = *reinterpret_cast<const float4*>(rb + (size_t)p.IW * p.KS + p.KS);
              rv.x = fmaxf(fmaxf(m0.x, m1.x), fmaxf(m2.x, m3.x)); rv.y = fmaxf(fmaxf(m0.y, m1.y), fmaxf(m2.y, m3.y));
              rv.z = fmaxf(fmaxf(m0.z, m1.z), fmaxf(m2.z, m3.z)); rv.w = fmaxf(fmaxf(m0.w, m1.w), fmaxf(m2.w, m3.w));
            } else {
              rv = *reinterpret_cast<const float4*>(sIn + (((size_t)g * p.IH + ty + p.dpt) * p.IW + tx + p.dpl) * p.KS + ch);
            }
            v[4 * h + 0] += rv.x; v[4 * h + 1] += rv.y; v[4 * h + 2] += rv.z; v[4 * h + 3] += rv.w;
          }
        } else if (p.res_mode == 2) {
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int ch = c + j;
            if (ch >= p.res_C) continue;
            float rv;
            if (p.res_pool) {
              rv = -INFINITY;
#pragma unroll
              for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                for (int dx = 0; dx < 2; ++dx) {
                  int ry = 2 * oy + dy, rx = 2 * ox + dx;
                  if (ry < p.res_H && rx < p.res_W) rv = fmaxf(rv, rbase[((size_t)ry * p.res_W + rx) * p.res_Cs + ch]);
                }
            } else {
              rv = rbase[((size_t)oy * p.res_W + ox) * p.res_Cs + ch];
            }
            v[j] += rv;
          }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = act1(v[j], p.act, sAlpha[c + j]);
        if (p.vec_store) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ch = c + 4 * h;
            if (ch >= p.CoutS) continue;
            float4 o = make_float4(ch + 0 < p.Cout ? v[4 * h] : 0.f, ch + 1 < p.Cout ? v[4 * h + 1] : 0.f,
                                   ch + 2 < p.Cout ? v[4 * h + 2] : 0.f, ch + 3 < p.Cout ? v[4 * h + 3] : 0.f);
            *reinterpret_cast<float4*>(orow + ch) = o;
          }
        } else {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c + j < p.CoutS) orow[c + j] = c + j < p.Cout ? v[j] : 0.f;
        }
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();   // TMEM, sIn and the A tiles are free again
  }
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)p.tmem_cols));
  }
}

}  // namespace

void launch_dwpw_tc(const DwPwTcP& p, int B, cudaStream_t s, int max_ctas) {
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      cudaFuncSetAttribute(k_dwpw_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      c = p.smem_bytes;
    }
  }
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  k_dwpw_tc<<<grid, kTcThreads, p.smem_bytes, s>>>(p, B, ntiles);
}

}  // namespace fdt
