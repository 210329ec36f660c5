// Tiled fp32 kernels for the conv stacks (XNNPACK conv2d / dwconv2d / add / clamp / prelu
// equivalents, fused).  Both kernels are persistent: a CTA keeps its weight chunk in shared
// memory and strides over tiles of 128 (or 64) output pixels.
//
//   k_gemm_conv : dense k x k convolution as im2col-in-shared-memory x register-tiled GEMM
//                 (BlazeFace stems; optionally reads the u8 letterboxed image and applies the
//                 [-1,1] normalisation + BGR->RGB swap on load).
//   k_dwpw      : BlazeBlock = [depthwise 3x3 (stride 1|2, TFLite SAME)] -> pointwise 1x1 + bias
//                 [+ residual (optional 2x2/2 max-pool, zero channel pad)] + ReLU/PReLU, the
//                 depthwise result staying in shared memory.
//
// GEMM micro-kernel: thread (pg, ng) owns TM pixel slots {pg + NPG*i} and 8 output channels
// (quads ng and ng+NNG of the current NC-wide chunk); A is pixel-major in smem with a row stride
// KS chosen so that KS/4 is odd (conflict-free 128-bit loads), W is k-major [KP][NC].
#include "kernels.h"

namespace fdt {
namespace {

__device__ __forceinline__ float act1(float v, int act, float alpha) {
  if (act == kActRelu) return fmaxf(v, 0.f);
  if (act == kActPrelu) return v >= 0.f ? v : v * alpha;
  return v;
}

template <int TM>
__device__ __forceinline__ void gemm_core(const float* __restrict__ sA, int KS, const float* __restrict__ sW,
                                          int NC, int KP, int pg, int NPG, int ng, int NNG,
                                          float (&acc)[TM][8]) {
  const float* a0 = sA + (size_t)pg * KS;
  const float* w0 = sW + 4 * ng;
  const float* w1 = sW + 4 * (ng + NNG);
  const size_t astep = (size_t)NPG * KS;
#pragma unroll 1
  for (int k = 0; k < KP; k += 4) {
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * astep + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b0 = *reinterpret_cast<const float4*>(w0 + (size_t)(k + kk) * NC);
      const float4 b1 = *reinterpret_cast<const float4*>(w1 + (size_t)(k + kk) * NC);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
        acc[i][0] = fmaf(av, b0.x, acc[i][0]);
        acc[i][1] = fmaf(av, b0.y, acc[i][1]);
        acc[i][2] = fmaf(av, b0.z, acc[i][2]);
        acc[i][3] = fmaf(av, b0.w, acc[i][3]);
        acc[i][4] = fmaf(av, b1.x, acc[i][4]);
        acc[i][5] = fmaf(av, b1.y, acc[i][5]);
        acc[i][6] = fmaf(av, b1.z, acc[i][6]);
        acc[i][7] = fmaf(av, b1.w, acc[i][7]);
      }
    }
  }
}

// Loads W[KP][c0 .. c0+NC) (global row stride CoutP) into smem [KP][NC].
__device__ __forceinline__ void load_weights(float* sW, const float* __restrict__ w, int KP, int CoutP, int c0,
                                             int NC, int tid, int nt) {
  const int nq = NC >> 2;
  for (int i = tid; i < KP * nq; i += nt) {
    int k = i / nq, q = i - k * nq;
    int c = c0 + 4 * q;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < CoutP) v = *reinterpret_cast<const float4*>(w + (size_t)k * CoutP + c);
    *reinterpret_cast<float4*>(sW + (size_t)k * NC + 4 * q) = v;
  }
}

// Epilogue store of one pixel's quad.  `dst` points at channel c of the pixel.
__device__ __forceinline__ void store_quad(float* dst, int c, int Cout, int CoutS, int vec, float4 v) {
  if (vec && c + 3 < CoutS) {
    if (c + 0 >= Cout) v.x = 0.f;
    if (c + 1 >= Cout) v.y = 0.f;
    if (c + 2 >= Cout) v.z = 0.f;
    if (c + 3 >= Cout) v.w = 0.f;
    *reinterpret_cast<float4*>(dst) = v;
  } else {
    if (c + 0 < CoutS) dst[0] = c + 0 < Cout ? v.x : 0.f;
    if (c + 1 < CoutS) dst[1] = c + 1 < Cout ? v.y : 0.f;
    if (c + 2 < CoutS) dst[2] = c + 2 < Cout ? v.z : 0.f;
    if (c + 3 < CoutS) dst[3] = c + 3 < Cout ? v.w : 0.f;
  }
}

// ------------------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(512) k_gemm_conv(GemmConvP p, int B, int ntiles) {
  extern __shared__ __align__(16) float smem[];
  const int P = TM * p.NPG;
  float* sA = smem;                          // [P][KS]
  float* sW = smem + (size_t)P * p.KS;       // [KP][NC]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ng = tid % p.NNG, pg = tid / p.NNG;
  const long long total_px = (long long)B * p.OH * p.OW;
  const int kwc = p.kw * p.Cin;

  for (int chunk = 0; chunk < p.nchunks; ++chunk) {
    const int c0 = chunk * p.NC;
    __syncthreads();
    load_weights(sW, p.w, p.KP, p.CoutP, c0, p.NC, tid, nt);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      __syncthreads();  // previous tile's GEMM done with sA (and weights visible on first pass)
      // ---- im2col: sA[slot][k], k = (ky*kw + kx)*Cin + c
      const long long px0 = (long long)tile * P;
      for (int i = tid; i < P * p.KP; i += nt) {
        int slot = i / p.KP, k = i - slot * p.KP;
        float v = 0.f;
        long long px = px0 + slot;
        if (k < p.K && px < total_px) {
          int ox = (int)(px % p.OW);
          long long r = px / p.OW;
          int oy = (int)(r % p.OH);
          int b = (int)(r / p.OH);
          int ky = k / kwc, rem = k - ky * kwc;
          int kx = rem / p.Cin, c = rem - kx * p.Cin;
          int iy = oy * p.sh + ky - p.pt, ix = ox * p.sw + kx - p.pl;
          if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
            if (p.in8) {
              // bgrMatToSignedFloat32 (helpers.dart:401-406): RGB channel c = BGR byte 2-c
              unsigned char u = p.in8[(size_t)b * p.in_istride + ((size_t)iy * p.W + ix) * 3 + (2 - c)];
              v = fmaf((float)u, 1.0f / 127.5f, -1.0f);
            } else {
              v = p.in[(size_t)b * p.in_istride + ((size_t)iy * p.W + ix) * p.CinS + c];
            }
          }
        }
        sA[(size_t)slot * p.KS + k] = v;
      }
      __syncthreads();
      float acc[TM][8];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_core<TM>(sA, p.KS, sW, p.NC, p.KP, pg, p.NPG, ng, p.NNG, acc);
      // ---- epilogue
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        long long px = px0 + pg + p.NPG * i;
        if (px >= total_px) continue;
        long long b = px / ((long long)p.OH * p.OW);
        long long sp = px - b * (long long)p.OH * p.OW;
        float* orow = p.out + b * p.out_istride + sp * p.CoutS;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int c = c0 + 4 * (ng + h * p.NNG);
          if (c >= p.CoutS) continue;
          float4 v;
          float* vv = &v.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int cc = c + j;
            float bias = p.bias[cc];   // bias / alpha are padded to CoutP + NC
            float al = p.alpha ? p.alpha[cc] : 0.f;
            vv[j] = act1(acc[i][4 * h + j] + bias, p.act, al);
          }
          store_quad(orow + c, c, p.Cout, p.CoutS, p.vec_store, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(512) k_dwpw(DwPwP p, int B, int ntiles) {
  extern __shared__ __align__(16) float smem[];
  const int P = TM * p.NPG;
  const int in_elems = p.G * p.IH * p.IW * p.KS;
  float* sIn = smem;                                  // [G][IH][IW][KS]
  float* sA = p.has_dw ? smem + in_elems : smem;      // [P][KS]  (aliases sIn without DW)
  float* sW = sA + (size_t)P * p.KS;                  // [KP][NC]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int ng = tid % p.NNG, pg = tid / p.NNG;
  const int Q = p.KP >> 2;
  const int tiles_per_group = p.tilesX * p.tilesY;
  const int thw = p.TH * p.TW;

  for (int chunk = 0; chunk < p.nchunks; ++chunk) {
    const int c0 = chunk * p.NC;
    __syncthreads();
    load_weights(sW, p.w, p.KP, p.CoutP, c0, p.NC, tid, nt);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int grp = tile / tiles_per_group;
      const int trem = tile - grp * tiles_per_group;
      const int ty0 = (trem / p.tilesX) * p.TH, tx0 = (trem % p.tilesX) * p.TW;
      const int b0 = grp * p.G;
      const int iy0 = ty0 * p.s - p.dpt, ix0 = tx0 * p.s - p.dpl;  // input coords of sIn(0,0)
      __syncthreads();
      // ---- stage the input tile (+halo), zero outside the image (TFLite SAME zero padding)
      for (int i = tid; i < p.G * p.IH * p.IW * Q; i += nt) {
        int q = i % Q;
        int r = i / Q;
        int lx = r % p.IW; r /= p.IW;
        int ly = r % p.IH;
        int g = r / p.IH;
        int b = b0 + g, y = iy0 + ly, x = ix0 + lx;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B && y >= 0 && y < p.H && x >= 0 && x < p.W && 4 * q < p.CinS)
          v = *reinterpret_cast<const float4*>(p.in + (size_t)b * p.in_istride + ((size_t)y * p.W + x) * p.CinS + 4 * q);
        *reinterpret_cast<float4*>(sIn + ((size_t)(g * p.IH + ly) * p.IW + lx) * p.KS + 4 * q) = v;
      }
      __syncthreads();
      if (p.has_dw) {
        // ---- depthwise 3x3: item = (g, tx, q) column strip, TH outputs each
        for (int i = tid; i < p.G * p.TW * Q; i += nt) {
          int q = i % Q;
          int r = i / Q;
          int tx = r % p.TW;
          int g = r / p.TW;
          float4 w[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const float4*>(p.dww + (size_t)t * p.KP + 4 * q);
          const float4 bias = *reinterpret_cast<const float4*>(p.dwb + 4 * q);
          const float* base = sIn + ((size_t)g * p.IH * p.IW + (size_t)tx * p.s) * p.KS + 4 * q;
          for (int ty = 0; ty < p.TH; ++ty) {
            float4 a = bias;
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              const float* row = base + (size_t)(ty * p.s + ky) * p.IW * p.KS;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) {
                const float4 v = *reinterpret_cast<const float4*>(row + (size_t)kx * p.KS);
                const float4 ww = w[ky * 3 + kx];
                a.x = fmaf(v.x, ww.x, a.x);
                a.y = fmaf(v.y, ww.y, a.y);
                a.z = fmaf(v.z, ww.z, a.z);
                a.w = fmaf(v.w, ww.w, a.w);
              }
            }
            int slot = g * thw + ty * p.TW + tx;
            *reinterpret_cast<float4*>(sA + (size_t)slot * p.KS + 4 * q) = a;
          }
        }
        __syncthreads();
      }
      float acc[TM][8];
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
      gemm_core<TM>(sA, p.KS, sW, p.NC, p.KP, pg, p.NPG, ng, p.NNG, acc);
      // ---- epilogue: bias + residual + activation
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        int slot = pg + p.NPG * i;
        if (slot >= p.G * thw) continue;
        int g = slot / thw, r = slot - g * thw;
        int oy = ty0 + r / p.TW, ox = tx0 + r % p.TW;
        int b = b0 + g;
        if (b >= B || oy >= p.OH || ox >= p.OW) continue;
        float* orow = p.out + (size_t)b * p.out_istride + ((size_t)oy * p.OW + ox) * p.CoutS;
        const float* rbase = p.res ? p.res + (size_t)b * p.res_istride : nullptr;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          int c = c0 + 4 * (ng + h * p.NNG);
          if (c >= p.CoutS) continue;
          float4 v;
          float* vv = &v.x;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            int cc = c + j;
            float x = acc[i][4 * h + j] + p.bias[cc];
            if (rbase && cc < p.res_C) {
              float rv;
              if (p.res_pool) {
                rv = -INFINITY;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                  for (int dx = 0; dx < 2; ++dx) {
                    int ry = 2 * oy + dy, rx = 2 * ox + dx;
                    if (ry < p.res_H && rx < p.res_W)
                      rv = fmaxf(rv, rbase[((size_t)ry * p.res_W + rx) * p.res_Cs + cc]);
                  }
              } else {
                rv = rbase[((size_t)oy * p.res_W + ox) * p.res_Cs + cc];
              }
              x += rv;
            }
            vv[j] = act1(x, p.act, p.alpha ? p.alpha[cc] : 0.f);
          }
          store_quad(orow + c, c, p.Cout, p.CoutS, p.vec_store, v);
        }
      }
    }
  }
}

}  // namespace

void launch_gemm_conv(const GemmConvP& p, int B, cudaStream_t s, int max_ctas) {
  const int P = p.TM * p.NPG;
  long long total_px = (long long)B * p.OH * p.OW;
  int ntiles = (int)((total_px + P - 1) / P);
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  int nt = p.NPG * p.NNG;
  if (p.TM == 8) {
    cudaFuncSetAttribute(k_gemm_conv<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    k_gemm_conv<8><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  } else {
    cudaFuncSetAttribute(k_gemm_conv<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    k_gemm_conv<4><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  }
}

void launch_dwpw(const DwPwP& p, int B, cudaStream_t s, int max_ctas) {
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  int nt = p.NPG * p.NNG;
  if (p.TM == 8) {
    cudaFuncSetAttribute(k_dwpw<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    k_dwpw<8><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  } else {
    cudaFuncSetAttribute(k_dwpw<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes);
    k_dwpw<4><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  }
}

}  // namespace fdt
