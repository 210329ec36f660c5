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

#include <algorithm>
#include <cstdlib>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.h"

namespace fdt {
namespace {

constexpr uint32_t kLBO = 128;   // bytes between the two 16-byte K chunks of one MMA (adjacent core matrices)
constexpr int kEpiWarps = 4;

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

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
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

__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
  a.x = fmaf(v.x, w.x, a.x); a.y = fmaf(v.y, w.y, a.y); a.z = fmaf(v.z, w.z, a.z); a.w = fmaf(v.w, w.w, a.w);
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
__device__ __forceinline__ void add4(float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; }
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }

// Shared-memory carve-up (must match ws_smem_bytes in plan.cpp):
//   [W (w_parts x Npad x K8)] [bias Npad] [alpha Npad] [dw taps+bias 10 x K8] [dtab n_items x 8 B] [barriers 128 B]
//   | 128-byte aligned: [A ring: NA x (hi, lo) x 128 x K8] [input ring: NS x in_stage_bytes]
template <int ND>
__global__ void __launch_bounds__((ND + kEpiWarps + 2) * 32, 1)
k_block_ws(const __grid_constant__ CUtensorMap tmap, DwPwTcP p, int B, int ntiles) {
  extern __shared__ __align__(128) float smem[];
  __shared__ uint32_t tmem_base_s;
  constexpr int kThreads = (ND + kEpiWarps + 2) * 32;
  constexpr int kDwThreads = ND * 32;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int Q8 = p.K8 >> 2;                       // 16-byte K chunks per row
  const uint32_t SBO = (uint32_t)Q8 * 128u;       // bytes between 8-row groups
  const int NS = p.ns, NA = p.na;
  float* sB = smem;
  float* sBias = sB + (size_t)p.w_parts * p.Npad * p.K8;
  float* sAlpha = sBias + p.Npad;
  float* sDw = sAlpha + p.Npad;
  uint2* dtab = reinterpret_cast<uint2*>(sDw + (p.has_dw ? 10 * p.K8 : 0));
  uint64_t* bars = reinterpret_cast<uint64_t*>(dtab + p.n_items);
  const uint32_t a_stage_floats = 2u * 128u * (uint32_t)p.K8;
  float* sA = reinterpret_cast<float*>(((uintptr_t)(bars + 16) + 127) & ~(uintptr_t)127);
  float* sIn0 = sA + (size_t)NA * a_stage_floats;
  const uint32_t in_stage_floats = (uint32_t)p.in_floats;     // multiple of 32 floats (128 B)
  // barrier slots: full_in[NS] | empty_in[NS] | a_full[NA] | a_empty[NA] | d_full[2] | d_empty[2]   (NS <= 4, NA <= 2)
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t full_in = bar0, empty_in = bar0 + 8u * 4, a_full = bar0 + 8u * 8, a_empty = bar0 + 8u * 10,
                 d_full = bar0 + 8u * 12, d_empty = bar0 + 8u * 14;
  const int thw = p.TH * p.TW;
  const int nslots = p.G * thw;
  const bool epi_reads_stage = p.res_mode == 1;

  // ---- prologue (all threads): weights, bias, tables, barriers, TMEM -------------------------------
  const uint32_t sB_u32 = smem_u32(sB);
  for (int i = tid; i < p.w_parts * p.Npad * Q8; i += kThreads) cp_async16_u32(sB_u32 + 16u * i, p.wB + 4 * (size_t)i);
  for (int i = tid; i < p.Npad; i += kThreads) {
    sBias[i] = p.bias[i];
    sAlpha[i] = p.alpha ? p.alpha[i] : 0.f;
  }
  if (p.has_dw) {
    for (int i = tid; i < 10 * p.K8; i += kThreads) sDw[i] = i < 9 * p.K8 ? p.dww[i] : p.dwb[i - 9 * p.K8];
    // depthwise table: item = (((g*Q8 + qq)*nstrips + st)*TW + tx)  (tx fastest => conflict-free LDS)
    for (int it = tid; it < p.n_items; it += kThreads) {
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
    for (int i = 0; i < NS; ++i) {
      mbar_init(full_in + 8u * i, 1);
      mbar_init(empty_in + 8u * i, ND + (epi_reads_stage ? kEpiWarps : 0));
    }
    for (int i = 0; i < NA; ++i) {
      mbar_init(a_full + 8u * i, ND);
      mbar_init(a_empty + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(d_full + 8u * i, 1);
      mbar_init(d_empty + 8u * i, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
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

  if (warp < kEpiWarps) {
    // =============================== epilogue warps =================================================
    const int slot = warp * 32 + lane;            // TMEM lane == pixel slot
    int e_g, e_r, e_ty, e_tx;
    p.fd_thw.divmod(slot, e_g, e_r);
    p.fd_TW.divmod(e_r, e_ty, e_tx);
    const bool slot_ok = slot < nslots;
    const long long o_rel = slot_ok ? (long long)e_g * p.out_istride + ((long long)e_ty * p.OW + e_tx) * p.CoutS : 0;
    const int rs = p.res_pool ? 2 : 1;
    const uint32_t res_off = slot_ok ? (uint32_t)((((size_t)e_g * p.IH + e_ty * rs + p.dpt) * p.IW + e_tx * rs + p.dpl) * p.KS) : 0u;
    const uint32_t row_f = (uint32_t)p.IW * p.KS;
    int si = 0, sph = 0, di = 0, dph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int grp, trem, tyi, txi;
      p.fd_tpg.divmod(tile, grp, trem);
      p.fd_tilesX.divmod(trem, tyi, txi);
      const int ty0 = tyi * p.TH, tx0 = txi * p.TW;
      const int b0 = grp * p.G;
      const float* res_s = sIn0 + (size_t)si * in_stage_floats + res_off;
      if (epi_reads_stage) mbar_wait(full_in + 8u * si, (uint32_t)sph);   // visibility of the TMA writes to this thread
      mbar_wait(d_full + 8u * di, (uint32_t)dph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int oy = ty0 + e_ty, ox = tx0 + e_tx, b = b0 + e_g;
      const bool valid = slot_ok && b < B && oy < p.OH && ox < p.OW;
      float* orow = p.out + (long long)b0 * p.out_istride + ((long long)ty0 * p.OW + tx0) * p.CoutS + o_rel;
      const float* rbase = p.res_mode == 2 ? p.res + (size_t)(valid ? b : 0) * p.res_istride : nullptr;
      const uint32_t tcol0 = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(di * p.Npad);
      for (int c0 = 0; c0 < p.Npad; c0 += 16) {
        uint32_t u[16];
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
              "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
            : "r"(tcol0 + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        if (!valid) continue;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int c = c0 + 4 * q;
          if (c >= p.CoutS) break;
          const float4 bv = ld4(sBias + c);
          float4 v = make_float4(__uint_as_float(u[4 * q]) + bv.x, __uint_as_float(u[4 * q + 1]) + bv.y,
                                 __uint_as_float(u[4 * q + 2]) + bv.z, __uint_as_float(u[4 * q + 3]) + bv.w);
          if (p.res_mode == 1) {
            // residual straight from the staged tile (channels >= Cin are the zero channel pad)
            if (c < p.res_lim) {
              if (p.res_pool) {
                const float* r1p = res_s + p.KS;
                const float* r2p = res_s + row_f;
                const float* r3p = r2p + p.KS;
                add4(v, max4(max4(ld4(res_s + c), ld4(r1p + c)), max4(ld4(r2p + c), ld4(r3p + c))));
              } else {
                add4(v, ld4(res_s + c));
              }
            }
          } else if (p.res_mode == 2) {
            float rv[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
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
            add4(v, make_float4(rv[0], rv[1], rv[2], rv[3]));
          }
          if (p.act == kActRelu) {
            v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
          } else if (p.act == kActPrelu) {
            const float4 a0 = ld4(sAlpha + c);
            v.x = v.x >= 0.f ? v.x : v.x * a0.x; v.y = v.y >= 0.f ? v.y : v.y * a0.y;
            v.z = v.z >= 0.f ? v.z : v.z * a0.z; v.w = v.w >= 0.f ? v.w : v.w * a0.w;
          }
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
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(d_empty + 8u * di);
        if (epi_reads_stage) mbar_arrive(empty_in + 8u * si);
      }
      if (++si == NS) { si = 0; sph ^= 1; }
      if (++di == 2) { di = 0; dph ^= 1; }
    }
  } else if (warp < kEpiWarps + ND) {
    // =============================== depthwise / A-operand warps ====================================
    const int dtid = tid - kEpiWarps * 32;
    const uint32_t row_f = (uint32_t)p.IW * p.KS;
    int si = 0, sph = 0, ai = 0, aph = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const float* sIn = sIn0 + (size_t)si * in_stage_floats;
      float* sAhi = sA + (size_t)ai * a_stage_floats;
      float* sAlo = sAhi + 128 * p.K8;
      mbar_wait(full_in + 8u * si, (uint32_t)sph);
      mbar_wait(a_empty + 8u * ai, (uint32_t)(aph ^ 1));
      if (p.has_dw) {
        for (int it = dtid; it < p.n_items; it += kDwThreads) {
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
        for (int it = dtid; it < nslots * Q8; it += kDwThreads) {
          int qq, sl;
          p.fd_nslots.divmod(it, qq, sl);
          const float4 a = ld4(sIn + (size_t)sl * p.KS + 4 * qq);
          split_store(sAhi, sAlo, (((uint32_t)sl >> 3) * SBO + (uint32_t)qq * kLBO + ((uint32_t)sl & 7u) * 16u) >> 2, a);
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of A -> visible to the MMA (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full + 8u * ai);
        mbar_arrive(empty_in + 8u * si);
      }
      if (++si == NS) { si = 0; sph ^= 1; }
      if (++ai == NA) { ai = 0; aph ^= 1; }
    }
  } else if (warp == kEpiWarps + ND) {
    // =============================== TMA producer ===================================================
    if (lane == 0) {
      const uint32_t stage_bytes = (uint32_t)(p.G * p.IH * p.IW * p.KS) * 4u;
      int si = 0, sph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        int grp, trem, tyi, txi;
        p.fd_tpg.divmod(tile, grp, trem);
        p.fd_tilesX.divmod(trem, tyi, txi);
        const int b0 = grp * p.G;
        const int iy0 = tyi * p.TH * p.s - p.dpt, ix0 = txi * p.TW * p.s - p.dpl;
        mbar_wait(empty_in + 8u * si, (uint32_t)(sph ^ 1));
        const uint32_t bar = full_in + 8u * si;
        mbar_expect_tx(bar, stage_bytes);
        const uint32_t dst = smem_u32(sIn0 + (size_t)si * in_stage_floats);
        asm volatile(
            "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
            ::"r"(dst), "l"(&tmap), "r"(0), "r"(ix0), "r"(iy0), "r"(b0), "r"(bar) : "memory");
        if (++si == NS) { si = 0; sph ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // =============================== MMA issuer =====================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);
      const int ksteps = p.K8 >> 3;
      const uint32_t b_lo = sB_u32 + (uint32_t)p.Npad * p.K8 * 4u;
      int ai = 0, aph = 0, di = 0, dph = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        mbar_wait(a_full + 8u * ai, (uint32_t)aph);
        mbar_wait(d_empty + 8u * di, (uint32_t)(dph ^ 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t a_hi = smem_u32(sA + (size_t)ai * a_stage_floats), a_lo = a_hi + 128u * (uint32_t)p.K8 * 4u;
        const uint32_t dcol = tmem_base + (uint32_t)(di * p.Npad);
        for (int ks = 0; ks < ksteps; ++ks) {
          const uint64_t db = make_desc(sB_u32 + ks * 2 * kLBO, SBO);
          const uint64_t dah = make_desc(a_hi + ks * 2 * kLBO, SBO);
          mma_tf32(dcol, dah, db, idesc, ks > 0 ? 1u : 0u);
          mma_tf32(dcol, make_desc(a_lo + ks * 2 * kLBO, SBO), db, idesc, 1u);
          // fp32 weights (face_landmark): W = W_hi + W_lo, third product A_hi * W_lo (A_lo * W_lo ~ 2^-22, dropped)
          if (p.w_parts > 1) mma_tf32(dcol, dah, make_desc(b_lo + ks * 2 * kLBO, SBO), idesc, 1u);
        }
        mma_commit(a_empty + 8u * ai);   // operand buffer reusable once these MMAs have read it
        mma_commit(d_full + 8u * di);    // accumulator complete
        if (++ai == NA) { ai = 0; aph ^= 1; }
        if (++di == 2) { di = 0; dph ^= 1; }
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
  Key key(p.in, cap, p.H, p.W, p.CinS, p.KS, p.IW, p.IH, p.G, p.in_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.CinS, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.CinS * 4, (cuuint64_t)p.W * p.CinS * 4, (cuuint64_t)p.in_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.KS, (cuuint32_t)p.IW, (cuuint32_t)p.IH, (cuuint32_t)p.G};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  cache[key] = tm;
  *out = tm;
  return true;
}

template <int ND>
void launch_ws_nd(const CUtensorMap& tm, const DwPwTcP& p, int B, int ntiles, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      cudaFuncSetAttribute(k_block_ws<ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
      c = p.smem_bytes;
    }
  }
  int grid = std::min(ntiles, 148);
  if (grid < 1) grid = 1;
  k_block_ws<ND><<<grid, (ND + kEpiWarps + 2) * 32, p.smem_bytes, s>>>(tm, p, B, ntiles);
}

}  // namespace

bool launch_block_ws(const DwPwTcP& p, int B, int cap, cudaStream_t s) {
  CUtensorMap tm;
  if (!input_tensor_map(p, cap, &tm)) return false;
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  if (p.nd == 12) launch_ws_nd<12>(tm, p, B, ntiles, s);
  else launch_ws_nd<8>(tm, p, B, ntiles, s);
  return true;
}

}  // namespace fdt
