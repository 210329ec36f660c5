// (fp16 variant of ts_probe.cu: kind::f16, A = hi + lo fp16 terms, two halves per TMEM column, W exact fp16)
// Stand-alone probe of tcgen05.mma with the A operand in TENSOR MEMORY (the ".ts" form):
//   D[128 x N] = A[128 x K] * W[N x K]^T,  A = A_hi + A_lo (TF32 split) written by the CUDA cores with tcgen05.st
//   (thread = TMEM lane = row m, K consecutive 32-bit columns), W in shared memory (no-swizzle K-major core matrices).
// It answers two questions the image-resident tail kernel depends on: (1) the A layout in TMEM is row m -> lane m,
// element k -> column k; (2) tcgen05.st -> tcgen05.wait::st -> fence -> barrier -> tcgen05.mma ordering is sufficient.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o ts_probe ts_probe.cu ; run: ./ts_probe [K N]
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstring>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}

__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(acc));
}

__global__ void __launch_bounds__(128) tsh_probe(const float* A, const float* W, float* D, int K, int N, int a_col0) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  const int KC = K / 8;                       // 16-byte chunks (8 halves) per row
  const uint32_t LBO = 128, SBO = KC * 128;
  __half* sB = reinterpret_cast<__half*>(smem);
  __shared__ __align__(8) uint64_t mbar;
  __shared__ uint32_t tmem_base_s;

  for (int i = tid; i < N * K; i += 128) {
    int n = i / K, k = i - n * K;
    size_t off = ((size_t)(n >> 3) * SBO + (size_t)(k >> 3) * LBO + (n & 7) * 16 + (k & 7) * 2) / 2;
    sB[off] = __float2half_rn(W[(size_t)n * K + k]);
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  // A: thread tid owns row m = tid = TMEM lane; hi part at columns [a_col0, a_col0 + K), lo part right after it
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  for (int part = 0; part < 2; ++part)
    for (int k0 = 0; k0 < K; k0 += 16) {
      uint32_t u[8];
      for (int j = 0; j < 8; ++j) {
        float a0 = A[(size_t)tid * K + k0 + 2 * j], a1 = A[(size_t)tid * K + k0 + 2 * j + 1];
        __half h0 = __float2half_rn(a0), h1 = __float2half_rn(a1);
        __half l0 = __float2half_rn(a0 - __half2float(h0)), l1 = __float2half_rn(a1 - __half2float(h1));
        __half2 v = part == 0 ? __halves2half2(h0, h1) : __halves2half2(l0, l1);      // low 16 bits = even k
        u[j] = *reinterpret_cast<uint32_t*>(&v);
      }
      asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_base + (uint32_t)(a_col0 + part * (K / 2) + k0 / 2)),
                   "r"(u[0]), "r"(u[1]), "r"(u[2]), "r"(u[3]), "r"(u[4]), "r"(u[5]), "r"(u[6]), "r"(u[7]) : "memory");
    }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (tid == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);     // D f32, A/B f16, K-major
    for (int part = 0; part < 2; ++part)
      for (int ks = 0; ks < K / 16; ++ks) {
        const uint64_t db = make_desc(smem_u32(sB) + ks * 256, LBO, SBO);
        mma_ts_f16(tmem_base, tmem_base + (uint32_t)(a_col0 + part * (K / 2) + ks * 8), db, idesc, (part | ks) ? 1u : 0u);
      }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
  }
  {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(smem_u32(&mbar)), "r"(0u) : "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < N; c0 += 8) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(lane_base + (uint32_t)c0));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 8; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(u[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

int main(int argc, char** argv) {
  int ok_all = 1;
  const int cases[][2] = {{16, 16}, {32, 32}, {48, 64}, {96, 96}, {96, 112}, {128, 128}};
  for (auto& c : cases) {
    const int K = argc > 2 ? atoi(argv[1]) : c[0], N = argc > 2 ? atoi(argv[2]) : c[1];
    std::vector<float> A(128 * K), W((size_t)N * K), D(128 * (size_t)N, 0.f);
    srand(1);
    for (auto& v : A) v = (float)rand() / RAND_MAX * 4.f - 2.f;
    for (auto& v : W) {                                   // fp16-origin weights: exact in TF32
      float f = (float)rand() / RAND_MAX * 2.f - 1.f;
      v = __half2float(__float2half_rn(f));
    }
    float *dA, *dW, *dD;
    cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
    cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dD, 0, D.size() * 4);
    const size_t smem = (size_t)N * K * 2 + 1024;
    cudaFuncSetAttribute(tsh_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    tsh_probe<<<1, 128, smem>>>(dA, dW, dD, K, N, 256);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, ref_max = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < N; ++n) {
        double r = 0;
        for (int k = 0; k < K; ++k) r += (double)A[(size_t)m * K + k] * W[(size_t)n * K + k];
        worst = fmax(worst, fabs(r - D[(size_t)m * N + n]));
        ref_max = fmax(ref_max, fabs(r));
      }
    const bool ok = e == cudaSuccess && worst <= 4e-6 * ref_max;
    printf("tsh_probe K=%d N=%d: %s, max|err| %.3e (rel %.3e) %s\n", K, N, cudaGetErrorString(e), worst, worst / ref_max, ok ? "OK" : "MISMATCH");
    if (!ok) {
      ok_all = 0;
      printf("  D[0][0..3] = %g %g %g %g ; D[1][0..3] = %g %g %g %g\n", D[0], D[1], D[2], D[3], D[N], D[N + 1], D[N + 2], D[N + 3]);
    }
    cudaFree(dA); cudaFree(dW); cudaFree(dD);
    if (argc > 2) break;
  }
  return ok_all ? 0 : 1;
}
