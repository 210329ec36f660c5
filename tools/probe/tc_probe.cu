// Stand-alone probe of the tcgen05 TF32 GEMM building block used by the BlazeBlock kernel:
// D[128 x N] = A[128 x K] * W[N x K]^T with A split into hi + lo TF32 parts (2 MMAs per K step, weights
// exactly representable), operands in the no-swizzle K-major core-matrix layout, accumulator in TMEM.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tc_probe tc_probe.cu ; run: ./tc_probe
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
  return d;                  // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(acc));
}

__global__ void __launch_bounds__(128) tc_probe(const float* A, const float* W, float* D, int K, int N) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int KQ = K / 4;
  const uint32_t LBO = 128, SBO = KQ * 128;
  float* sAhi = reinterpret_cast<float*>(smem);
  float* sAlo = sAhi + 128 * K;
  float* sB = sAlo + 128 * K;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  // operands -> canonical layout
  for (int kq = 0; kq < KQ; ++kq) {
    float4 a = *reinterpret_cast<const float4*>(A + (size_t)tid * K + 4 * kq);
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(a.x) & 0xFFFFE000u); lo.x = a.x - hi.x;
    hi.y = __uint_as_float(__float_as_uint(a.y) & 0xFFFFE000u); lo.y = a.y - hi.y;
    hi.z = __uint_as_float(__float_as_uint(a.z) & 0xFFFFE000u); lo.z = a.z - hi.z;
    hi.w = __uint_as_float(__float_as_uint(a.w) & 0xFFFFE000u); lo.w = a.w - hi.w;
    size_t off = ((size_t)(tid >> 3) * SBO + (size_t)kq * LBO + (tid & 7) * 16) / 4;
    *reinterpret_cast<float4*>(sAhi + off) = hi;
    *reinterpret_cast<float4*>(sAlo + off) = lo;
  }
  for (int i = tid; i < N * KQ; i += 128) {
    int n = i / KQ, kq = i - n * KQ;
    float4 w = *reinterpret_cast<const float4*>(W + (size_t)n * K + 4 * kq);
    size_t off = ((size_t)(n >> 3) * SBO + (size_t)kq * LBO + (n & 7) * 16) / 4;
    *reinterpret_cast<float4*>(sB + off) = w;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(128u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t a_hi = smem_u32(sAhi), a_lo = smem_u32(sAlo), b = smem_u32(sB);
    for (int ks = 0; ks < K / 8; ++ks) {
      uint64_t db = make_desc(b + ks * 2 * LBO, LBO, SBO);
      mma_tf32(tmem_base, make_desc(a_hi + ks * 2 * LBO, LBO, SBO), db, idesc, ks > 0 ? 1u : 0u);
      mma_tf32(tmem_base, make_desc(a_lo + ks * 2 * LBO, LBO, SBO), db, idesc, 1u);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  // wait for the MMAs
  {
    uint32_t ok = 0;
    while (!ok) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  // epilogue: warp w reads TMEM lanes 32w..32w+31
  for (int c = 0; c < N; c += 8) {
    uint32_t v[8];
    uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)(warp * 32 + lane) * N + c + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(128u));
}

static float half_round(float x) {   // value exactly representable in fp16 (like the detector weights)
  int e; float m = frexpf(x, &e); return ldexpf(roundf(m * 2048.f) / 2048.f, e);
}

int main() {
  int cfgs[][2] = {{24, 32}, {32, 32}, {48, 48}, {56, 64}, {96, 96}, {88, 16}, {40, 48}, {8, 16}};
  int bad = 0;
  for (auto& c : cfgs) {
    int K = c[0], N = c[1];
    std::vector<float> A(128 * K), W(N * K), D(128 * N, -777.f);
    srand(K * 131 + N);
    for (auto& v : A) v = ((rand() % 20001) - 10000) / 977.0f;
    for (auto& v : W) v = half_round(((rand() % 20001) - 10000) / 9777.0f);
    float *dA, *dW, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice);
    size_t smem = (size_t)(2 * 128 * K + N * K) * 4;
    cudaFuncSetAttribute(tc_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tc_probe<<<1, 128, smem>>>(dA, dW, dD, K, N);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("K=%d N=%d CUDA error %s\n", K, N, cudaGetErrorString(e)); return 2; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double maxerr = 0, maxref = 0, maxerr32 = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double ref = 0; float r32 = 0;
        for (int k = 0; k < K; ++k) { ref += (double)A[m * K + k] * W[n * K + k]; r32 = fmaf(A[m * K + k], W[n * K + k], r32); }
        maxerr = fmax(maxerr, fabs(D[m * N + n] - ref)); maxref = fmax(maxref, fabs(ref));
        maxerr32 = fmax(maxerr32, fabs(r32 - ref));
      }
    printf("K=%3d N=%3d  max|err|=%.3e  max|ref|=%.3f  rel=%.3e   (fp32 fma chain rel=%.3e)\n", K, N, maxerr, maxref, maxerr / maxref, maxerr32 / maxref);
    if (!(maxerr / maxref < 1e-5)) ++bad;
    cudaFree(dA); cudaFree(dW); cudaFree(dD);
  }
  printf(bad ? "PROBE FAILED (%d configs)\n" : "PROBE OK\n", bad);
  return bad ? 1 : 0;
}
