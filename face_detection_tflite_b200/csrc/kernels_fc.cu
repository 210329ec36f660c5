// k_fc_tc — a convolution whose window is the whole input map (the face-landmark head: 3x3x32 -> 1404), i.e. one dense
// contraction per image: out[b][n] = sum_k in[b][k] * W[n][k] + bias[n], as a tcgen05 GEMM over the images of a chunk.
//
//   grid = (N tiles of 128 output channels, M tiles of 128 images); 8 warps.
//   W tile [w_parts x 128 x K16] fp16 (UMMA K-major core matrices, prepared by the planner; fp32-origin weights are scaled by
//     a power of two and split into hi + lo) -> shared memory by ONE bulk copy (<= 144 KB).
//   A: thread = image row (= TMEM lane); its K floats -> fp16 hi + lo -> tcgen05.st into TENSOR MEMORY (hi halves in the first
//     K16 / 2 columns of the operand area, lo halves in the next K16 / 2); the two warps of a lane quarter take alternate quads.
//   MMA (one thread): kind::f16, A from TMEM: A_hi x W_hi + A_lo x W_hi [+ A_hi x W_lo]  -> 128 fp32 accumulator columns.
//   epilogue: tcgen05.ld, * wscale + bias, stores guarded by n < N and image < B.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <map>
#include <mutex>

#include "kernels.h"

namespace fdt {
namespace {

constexpr int kFcThreads = 256;
constexpr uint32_t kLBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
__device__ __forceinline__ void split_store_tmem(uint32_t taddr_hi, uint32_t lo_off, const float4& a) {
  const uint32_t h0 = pack_h2(a.x, a.y), h1 = pack_h2(a.z, a.w);
  const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
  const uint32_t l0 = pack_h2(a.x - f0.x, a.y - f0.y), l1 = pack_h2(a.z - f1.x, a.w - f1.y);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi), "r"(h0), "r"(h1) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi + lo_off), "r"(l0), "r"(l1) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
}

// TMEM columns: accumulator [0, 128), operand [128, 128 + K16): K16 <= 384
__global__ void __launch_bounds__(kFcThreads, 1) k_fc_tc(FcP p, int B) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bars[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = (int)blockIdx.x * 128, b0 = (int)blockIdx.y * 128;
  const uint32_t w_a = (smem_u32(smem_raw) + 127u) & ~127u;
  const uint32_t w_full = smem_u32(&bars[0]), d_full = smem_u32(&bars[1]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(w_full));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(d_full));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(w_full), "r"((uint32_t)p.tile_bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(w_a), "l"(reinterpret_cast<const unsigned char*>(p.w) + (size_t)blockIdx.x * p.tile_bytes), "r"((uint32_t)p.tile_bytes), "r"(w_full) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int lq = warp & 3, g = warp >> 2;
  const int row = lq * 32 + lane, img = b0 + row;
  const uint32_t tm_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
  const uint32_t acol = tm_lane + 128u, lo_off = (uint32_t)(p.K16 >> 1);
  // ---- A operand: this image's K floats -> fp16 hi / lo -> TMEM -----------------------------------------------------
  {
    const float* src = p.in + (size_t)(img < B ? img : 0) * p.in_istride;
    const int nq = p.K16 >> 2, nq_real = p.K >> 2;
    for (int q = g; q < nq; q += 2) {
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (img < B && q < nq_real) v = __ldg(reinterpret_cast<const float4*>(src) + q);
      split_store_tmem(acol + 2u * (uint32_t)q, lo_off, v);
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  // ---- MMA ---------------------------------------------------------------------------------------------------------------
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    mbar_wait(w_full, 0u);
    const uint32_t idesc = (1u << 4) | ((128u >> 3) << 17) | ((128u >> 4) << 24);     // D f32, A / B f16, K-major, N = 128, M = 128
    const uint32_t sbo = (uint32_t)(p.K16 >> 3) * 128u;
    const uint32_t b_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t b_lo0 = ((w_a & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
    const uint32_t b_lo1 = (((w_a + 128u * (uint32_t)p.K16 * 2u) & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
    const uint32_t a0 = tmem_base + 128u;
    const int ksteps = p.K16 >> 4;
#pragma unroll 1
    for (int ks = 0; ks < ksteps; ++ks) {
      mma_ts_f16(tmem_base, a0 + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, ks ? 1u : 0u);
      mma_ts_f16(tmem_base, a0 + lo_off + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
      if (p.w_parts == 2) mma_ts_f16(tmem_base, a0 + 8u * (uint32_t)ks, b_lo1 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(d_full) : "memory");
  }
  // ---- epilogue: each warp group takes 64 of the 128 columns ---------------------------------------------------------
  mbar_wait(d_full, 0u);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  float* orow = p.out + (size_t)(img < B ? img : 0) * p.out_istride;
#pragma unroll 1
  for (int c16 = 0; c16 < 4; ++c16) {
    const int c0 = g * 64 + c16 * 16;
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
          "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(tm_lane + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (img < B) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int n = n0 + c0 + j;
        if (n < p.N) orow[n] = fmaf(__uint_as_float(u[j]), p.wscale, __ldg(p.bias + n));
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

}  // namespace

bool launch_fc_tc(const FcP& p, int B, cudaStream_t s) {
  if (B <= 0) return true;
  static std::mutex mu;
  static std::map<int, size_t> cur;
  const size_t smem = (size_t)p.tile_bytes + 256;
  int dev = 0;
  cudaGetDevice(&dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (smem > c) {
      if (cudaFuncSetAttribute(k_fc_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return false;
      c = smem;
    }
  }
  dim3 grid((unsigned)((p.N + 127) / 128), (unsigned)((B + 127) / 128));
  k_fc_tc<<<grid, kFcThreads, smem, s>>>(p, B);
  return true;
}

}  // namespace fdt
