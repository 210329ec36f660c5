// Tiled fp32 kernels for the conv stacks (XNNPACK conv2d / dwconv2d / add / clamp / prelu
// equivalents, fused).  All three are persistent: a CTA keeps its weight chunk in shared memory
// and strides over tiles of 128 (or 64) output pixels.
//
//   k_stem      : k x k / stride-2 convolution of the 3-channel input image, read as u8x4 BGRX
//                 (letterboxed frame or face crop) with the [-1,1] normalisation + BGR->RGB swap
//                 applied while staging the patch into shared memory; each thread owns a strip of 4
//                 horizontally adjacent output pixels x 4 output channels and slides over the patch.
//   k_dwpw      : BlazeBlock = [depthwise 3x3 (stride 1|2, TFLite SAME)] -> pointwise 1x1 + bias
//                 [+ residual (optional 2x2/2 max-pool, zero channel pad)] + ReLU/PReLU.  The input
//                 tile (+halo) is staged with cp.async, the depthwise result and the residual never
//                 leave shared memory.
//   k_gemm_conv : any other dense convolution as im2col-in-shared-memory x GEMM (mesh head etc.).
//
// GEMM micro-kernel: thread (pg, q) owns TM pixel slots {pg + NPG*i} and the 4 output channels of
// quad q; A is pixel-major in smem with a row stride KS such that KS/4 is odd (conflict-free 128-bit
// loads for consecutive pixels), W is k-major [KP][NC].  A warp covers ~32/NQ pixel groups x NQ
// quads, so both operand loads are mostly broadcasts (1-3 wavefronts per LDS.128).
#include <map>
#include <mutex>
#include <utility>

#include <algorithm>

#include "kernels.h"

#ifndef FDT_MINB
#define FDT_MINB 1   // min resident CTAs per SM the register allocator must allow (tuning knob)
#endif

namespace fdt {
namespace {

__device__ __forceinline__ float act1(float v, int act, float alpha) {
  if (act == kActRelu) return fmaxf(v, 0.f);
  if (act == kActPrelu) return v >= 0.f ? v : v * alpha;
  return v;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;  // src-size 0 => 16 bytes of zeros (TFLite SAME zero padding)
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\n" ::);
  asm volatile("cp.async.wait_group 0;\n" ::: "memory");
}

template <int TM>
__device__ __forceinline__ void gemm_core(const float* __restrict__ sA, int KS, const float* __restrict__ sW,
                                          int NC, int KP, int pg, int NPG, int q, float (&acc)[TM][4]) {
  const float* a0 = sA + (size_t)pg * KS;
  const float* w0 = sW + 4 * q;
  const size_t astep = (size_t)NPG * KS;
#pragma unroll 2
  for (int k = 0; k < KP; k += 4) {
    float4 a[TM];
#pragma unroll
    for (int i = 0; i < TM; ++i) a[i] = *reinterpret_cast<const float4*>(a0 + i * astep + k);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const float4 b = *reinterpret_cast<const float4*>(w0 + (size_t)(k + kk) * NC);
#pragma unroll
      for (int i = 0; i < TM; ++i) {
        const float av = kk == 0 ? a[i].x : (kk == 1 ? a[i].y : (kk == 2 ? a[i].z : a[i].w));
        acc[i][0] = fmaf(av, b.x, acc[i][0]);
        acc[i][1] = fmaf(av, b.y, acc[i][1]);
        acc[i][2] = fmaf(av, b.z, acc[i][2]);
        acc[i][3] = fmaf(av, b.w, acc[i][3]);
      }
    }
  }
}

// Loads W[KP][c0 .. c0+NC) (global row stride CoutP) into smem [KP][NC].
__device__ __forceinline__ void load_weights(float* sW, const float* __restrict__ w, int KP, int CoutP, int c0,
                                             int NC, int tid, int nt) {
  const int nq = NC >> 2;
  for (int i = tid; i < KP * nq; i += nt) {   // once per CTA and chunk: plain division is fine here
    int k = i / nq, q = i - k * nq;
    int c = c0 + 4 * q;
    cp_async16(sW + (size_t)k * NC + 4 * q, c < CoutP ? w + (size_t)k * CoutP + c : w, c < CoutP);
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

// ------------------------------------------------------------------------------------------------
// Stem: KW x KW, stride 2, Cin = 3 from a u8x4 BGRX image.  Tile = 8 x 16 output pixels = 32 strips
// of 4 pixels; thread = (strip, quad).  Strips are numbered row-fastest so that the strips of one
// warp sit in different patch rows (row stride PW float4, PW odd => conflict-free).
template <int KW>
__global__ void __launch_bounds__(512) k_stem(StemP p, int B, int ntiles) {
  constexpr int TH = 8, TW = 16, SW = 4;
  constexpr int PH = (TH - 1) * 2 + KW, PW = (TW - 1) * 2 + KW, NCOL = (SW - 1) * 2 + KW;
  extern __shared__ __align__(16) float smem[];
  float4* sP = reinterpret_cast<float4*>(smem);        // [PH][PW] normalised RGB0
  float* sW = smem + PH * PW * 4;                      // [KP][NC]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int NQ = p.NC >> 2;
  const int q = tid % NQ, strip = tid / NQ;
  const int sy = strip % TH, sx = strip / TH;          // row-fastest
  const int tilesX = (p.OW + TW - 1) / TW, tilesY = (p.OH + TH - 1) / TH;
  const int tiles_per_img = tilesX * tilesY;

  load_weights(sW, p.w, p.KP, p.CoutP, 0, p.NC, tid, nt);
  cp_async_wait_all();
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int b = tile / tiles_per_img;
    const int trem = tile - b * tiles_per_img;
    const int ty0 = (trem / tilesX) * TH, tx0 = (trem % tilesX) * TW;
    const int iy0 = ty0 * 2 - p.pt, ix0 = tx0 * 2 - p.pl;
    __syncthreads();
    const uchar4* img = reinterpret_cast<const uchar4*>(p.in8) + (size_t)b * p.H * p.W;
    for (int i = tid; i < PH * PW; i += nt) {
      int ly = i / PW, lx = i - ly * PW;
      int y = iy0 + ly, x = ix0 + lx;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (y >= 0 && y < p.H && x >= 0 && x < p.W) {
        uchar4 u = img[(size_t)y * p.W + x];
        // bgrMatToSignedFloat32 (helpers.dart:401-406): RGB = (u.z, u.y, u.x), v * (1/127.5) - 1
        v.x = fmaf((float)u.z, 1.0f / 127.5f, -1.0f);
        v.y = fmaf((float)u.y, 1.0f / 127.5f, -1.0f);
        v.z = fmaf((float)u.x, 1.0f / 127.5f, -1.0f);
      }
      sP[i] = v;
    }
    __syncthreads();
    float acc[SW][4];
#pragma unroll
    for (int j = 0; j < SW; ++j) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
    const float* wq = sW + 4 * q;
#pragma unroll 1
    for (int ky = 0; ky < KW; ++ky) {
      float4 px[NCOL];
      const float4* row = sP + (2 * sy + ky) * PW + sx * (2 * SW);
#pragma unroll
      for (int c = 0; c < NCOL; ++c) px[c] = row[c];
#pragma unroll
      for (int kx = 0; kx < KW; ++kx) {
        const float* wk = wq + (size_t)((ky * KW + kx) * 3) * p.NC;
        const float4 w0 = *reinterpret_cast<const float4*>(wk);
        const float4 w1 = *reinterpret_cast<const float4*>(wk + p.NC);
        const float4 w2 = *reinterpret_cast<const float4*>(wk + 2 * p.NC);
#pragma unroll
        for (int j = 0; j < SW; ++j) {
          const float4 a = px[2 * j + kx];
          acc[j][0] = fmaf(a.x, w0.x, acc[j][0]); acc[j][1] = fmaf(a.x, w0.y, acc[j][1]);
          acc[j][2] = fmaf(a.x, w0.z, acc[j][2]); acc[j][3] = fmaf(a.x, w0.w, acc[j][3]);
          acc[j][0] = fmaf(a.y, w1.x, acc[j][0]); acc[j][1] = fmaf(a.y, w1.y, acc[j][1]);
          acc[j][2] = fmaf(a.y, w1.z, acc[j][2]); acc[j][3] = fmaf(a.y, w1.w, acc[j][3]);
          acc[j][0] = fmaf(a.z, w2.x, acc[j][0]); acc[j][1] = fmaf(a.z, w2.y, acc[j][1]);
          acc[j][2] = fmaf(a.z, w2.z, acc[j][2]); acc[j][3] = fmaf(a.z, w2.w, acc[j][3]);
        }
      }
    }
    const int oy = ty0 + sy, c = 4 * q;
    if (oy < p.OH && c < p.CoutS) {
      const float4 bias = *reinterpret_cast<const float4*>(p.bias + c);
      float4 al = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.alpha) al = *reinterpret_cast<const float4*>(p.alpha + c);
#pragma unroll
      for (int j = 0; j < SW; ++j) {
        int ox = tx0 + sx * SW + j;
        if (ox >= p.OW) continue;
        float4 v;
        v.x = act1(acc[j][0] + bias.x, p.act, al.x);
        v.y = act1(acc[j][1] + bias.y, p.act, al.y);
        v.z = act1(acc[j][2] + bias.z, p.act, al.z);
        v.w = act1(acc[j][3] + bias.w, p.act, al.w);
        store_quad(p.out + (size_t)b * p.out_istride + ((size_t)oy * p.OW + ox) * p.CoutS + c, c, p.Cout, p.CoutS, p.vec_store, v);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(384, FDT_MINB) k_gemm_conv(GemmConvP p, int B, int ntiles) {
  extern __shared__ __align__(16) float smem[];
  const int P = TM * p.NPG;
  float* sA = smem;                          // [P][KS]
  float* sW = smem + (size_t)P * p.KS;       // [KP][NC]
  const int tid = threadIdx.x, nt = blockDim.x;
  const int NQ = p.NC >> 2;
  int q, pg;
  p.fd_NQ.divmod(tid, pg, q);          // threads beyond NPG*NQ only help with the im2col staging
  const long long total_px = (long long)B * p.OH * p.OW;

  // output-channel chunks are spread over gridDim.y (a [faces x 288] x [288 x 1404] head has only a handful of pixel tiles)
  for (int chunk = blockIdx.y; chunk < p.nchunks; chunk += gridDim.y) {
    const int c0 = chunk * p.NC;
    __syncthreads();
    load_weights(sW, p.w, p.KP, p.CoutP, c0, p.NC, tid, nt);
    cp_async_wait_all();
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      __syncthreads();  // previous tile's GEMM done with sA (and weights visible on first pass)
      const long long px0 = (long long)tile * P;
      if (p.flat) {
        // the window is the whole dense input: row = the image itself, copied 16 bytes at a time
        const int kq = p.KP >> 2;
        for (int i = tid; i < P * kq; i += nt) {
          int slot = i / kq, k4 = i - slot * kq;
          long long px = px0 + slot;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (px < total_px && 4 * k4 < p.K) v = *reinterpret_cast<const float4*>(p.in + (size_t)px * p.in_istride + 4 * k4);
          *reinterpret_cast<float4*>(sA + (size_t)slot * p.KS + 4 * k4) = v;
        }
      } else {
        // ---- im2col: sA[slot][k], k = (ky*kw + kx)*Cin + c
        for (int i = tid; i < P * p.KP; i += nt) {
          int slot, k;
          p.fd_KP.divmod(i, slot, k);
          float v = 0.f;
          long long px = px0 + slot;
          if (k < p.K && px < total_px) {
            int r, ox, b, oy, ky, rem, kx, c;
            p.fd_OW.divmod((int)px, r, ox);
            p.fd_OH.divmod(r, b, oy);
            p.fd_kwc.divmod(k, ky, rem);
            p.fd_Cin.divmod(rem, kx, c);
            int iy = oy * p.sh + ky - p.pt, ix = ox * p.sw + kx - p.pl;
            if (iy >= 0 && iy < p.H && ix >= 0 && ix < p.W) {
              if (p.in8) {
                unsigned char u = p.in8[(size_t)b * p.in_istride + ((size_t)iy * p.W + ix) * 4 + (2 - c)];
                v = fmaf((float)u, 1.0f / 127.5f, -1.0f);
              } else {
                v = p.in[(size_t)b * p.in_istride + ((size_t)iy * p.W + ix) * p.CinS + c];
              }
            }
          }
          sA[(size_t)slot * p.KS + k] = v;
        }
      }
      __syncthreads();
      if (pg >= p.NPG) continue;      // staging helper threads (no barrier below inside this tile)
      float acc[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      gemm_core<TM>(sA, p.KS, sW, p.NC, p.KP, pg, p.NPG, q, acc);
      const int c = c0 + 4 * q;
      if (c < p.CoutS) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          long long px = px0 + pg + p.NPG * i;
          if (px >= total_px) continue;
          long long b = px / ((long long)p.OH * p.OW);
          long long sp = px - b * (long long)p.OH * p.OW;
          float4 v;
          float* vv = &v.x;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            vv[j] = act1(acc[i][j] + p.bias[c + j], p.act, p.alpha ? p.alpha[c + j] : 0.f);
          store_quad(p.out + b * p.out_istride + sp * p.CoutS + c, c, p.Cout, p.CoutS, p.vec_store, v);
        }
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
template <int TM>
__global__ void __launch_bounds__(384, FDT_MINB) k_dwpw(DwPwP p, int B, int ntiles) {
  extern __shared__ __align__(16) float smem[];
  const int P = TM * p.NPG;
  const int in_elems = p.G * p.IH * p.IW * p.KS;
  float* sIn = smem;                                  // [G][IH][IW][KS]
  float* sA = p.has_dw ? smem + in_elems : smem;      // [P][KS]  (aliases sIn without DW)
  float* sW = sA + (size_t)P * p.KS;                  // [KP][NC]
  const int tid = threadIdx.x, nt = blockDim.x;
  int q, pg;
  p.fd_NQ.divmod(tid, pg, q);
  const int Q = p.KP >> 2;
  const int thw = p.TH * p.TW;

  // output-channel chunks are spread over gridDim.y (a [faces x 288] x [288 x 1404] head has only a handful of pixel tiles)
  for (int chunk = blockIdx.y; chunk < p.nchunks; chunk += gridDim.y) {
    const int c0 = chunk * p.NC;
    __syncthreads();
    load_weights(sW, p.w, p.KP, p.CoutP, c0, p.NC, tid, nt);
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      int grp, trem, tyi, txi;
      p.fd_tpg.divmod(tile, grp, trem);
      p.fd_tilesX.divmod(trem, tyi, txi);
      const int ty0 = tyi * p.TH, tx0 = txi * p.TW;
      const int b0 = grp * p.G;
      const int iy0 = ty0 * p.s - p.dpt, ix0 = tx0 * p.s - p.dpl;  // input coords of sIn(0,0)
      __syncthreads();
      // ---- stage the input tile (+halo) with cp.async; zero outside the image (TFLite SAME padding)
      {
        const int total = p.G * p.IH * p.IW * Q;
        for (int i = tid; i < total; i += nt) {
          int pix, qq, gy, lx, g, ly;
          p.fd_Q.divmod(i, pix, qq);
          p.fd_IW.divmod(pix, gy, lx);
          p.fd_IH.divmod(gy, g, ly);
          int b = b0 + g, y = iy0 + ly, x = ix0 + lx;
          bool ok = b < B && y >= 0 && y < p.H && x >= 0 && x < p.W;
          const float* src = ok ? p.in + (size_t)b * p.in_istride + ((size_t)y * p.W + x) * p.CinS + 4 * qq : p.in;
          cp_async16(sIn + ((size_t)gy * p.IW + lx) * p.KS + 4 * qq, src, ok);
        }
      }
      cp_async_wait_all();
      __syncthreads();
      if (p.has_dw) {
        // ---- depthwise 3x3: item = (g, q, tx) column strip (tx fastest: consecutive lanes read
        //      consecutive pixels, stride KS floats => conflict-free), TH outputs each
        const int nitems = p.G * Q * p.TW;
        for (int it = tid; it < nitems; it += nt) {
          int tx, r, qq, g;
          p.fd_TW.divmod(it, r, tx);
          p.fd_Q.divmod(r, g, qq);
          float4 w[9];
#pragma unroll
          for (int t = 0; t < 9; ++t) w[t] = *reinterpret_cast<const float4*>(p.dww + (size_t)t * p.KP + 4 * qq);
          const float4 bias = *reinterpret_cast<const float4*>(p.dwb + 4 * qq);
          const float* base = sIn + ((size_t)g * p.IH * p.IW + (size_t)tx * p.s) * p.KS + 4 * qq;
          float* dst = sA + ((size_t)g * thw + tx) * p.KS + 4 * qq;
          const size_t rstride = (size_t)p.IW * p.KS;
          if (p.s == 1) {
            // rolling 3-row window
            float4 r0[3], r1[3], r2[3];
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
              r0[kx] = *reinterpret_cast<const float4*>(base + (size_t)kx * p.KS);
              r1[kx] = *reinterpret_cast<const float4*>(base + rstride + (size_t)kx * p.KS);
            }
            for (int ty = 0; ty < p.TH; ++ty) {
#pragma unroll
              for (int kx = 0; kx < 3; ++kx)
                r2[kx] = *reinterpret_cast<const float4*>(base + (size_t)(ty + 2) * rstride + (size_t)kx * p.KS);
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
                a.x = fmaf(r2[kx].x, w[6 + kx].x, a.x); a.y = fmaf(r2[kx].y, w[6 + kx].y, a.y);
                a.z = fmaf(r2[kx].z, w[6 + kx].z, a.z); a.w = fmaf(r2[kx].w, w[6 + kx].w, a.w);
              }
              *reinterpret_cast<float4*>(dst + (size_t)ty * p.TW * p.KS) = a;
#pragma unroll
              for (int kx = 0; kx < 3; ++kx) { r0[kx] = r1[kx]; r1[kx] = r2[kx]; }
            }
          } else {
            for (int ty = 0; ty < p.TH; ++ty) {
              float4 a = bias;
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
                const float* row = base + (size_t)(ty * p.s + ky) * rstride;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  const float4 v = *reinterpret_cast<const float4*>(row + (size_t)kx * p.KS);
                  const float4 ww = w[ky * 3 + kx];
                  a.x = fmaf(v.x, ww.x, a.x); a.y = fmaf(v.y, ww.y, a.y);
                  a.z = fmaf(v.z, ww.z, a.z); a.w = fmaf(v.w, ww.w, a.w);
                }
              }
              *reinterpret_cast<float4*>(dst + (size_t)ty * p.TW * p.KS) = a;
            }
          }
        }
        __syncthreads();
      }
      float acc[TM][4];
#pragma unroll
      for (int i = 0; i < TM; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
      gemm_core<TM>(sA, p.KS, sW, p.NC, p.KP, pg, p.NPG, q, acc);
      // ---- epilogue: bias + residual + activation
      const int c = c0 + 4 * q;
      if (c < p.CoutS) {
        const float4 bias = *reinterpret_cast<const float4*>(p.bias + c);
        float4 al = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.alpha) al = *reinterpret_cast<const float4*>(p.alpha + c);
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          int slot = pg + p.NPG * i;
          if (slot >= p.G * thw) continue;
          int g, r, ty, tx;
          p.fd_thw.divmod(slot, g, r);
          p.fd_TW.divmod(r, ty, tx);
          int oy = ty0 + ty, ox = tx0 + tx;
          int b = b0 + g;
          if (b >= B || oy >= p.OH || ox >= p.OW) continue;
          float4 v = make_float4(acc[i][0] + bias.x, acc[i][1] + bias.y, acc[i][2] + bias.z, acc[i][3] + bias.w);
          if (p.res_mode == 1) {
            // residual straight from the staged tile (channels >= Cin are the zero channel pad)
            if (c < p.KP) {
              if (p.res_pool) {
                const float* rb = sIn + (((size_t)g * p.IH + 2 * ty + p.dpt) * p.IW + 2 * tx + p.dpl) * p.KS + c;
                float4 m0 = *reinterpret_cast<const float4*>(rb);
                float4 m1 = *reinterpret_cast<const float4*>(rb + p.KS);
                float4 m2 = *reinterpret_cast<const float4*>(rb + (size_t)p.IW * p.KS);
                float4 m3 = *reinterpret_cast<const float4*>(rb + (size_t)p.IW * p.KS + p.KS);
                v.x += fmaxf(fmaxf(m0.x, m1.x), fmaxf(m2.x, m3.x));
                v.y += fmaxf(fmaxf(m0.y, m1.y), fmaxf(m2.y, m3.y));
                v.z += fmaxf(fmaxf(m0.z, m1.z), fmaxf(m2.z, m3.z));
                v.w += fmaxf(fmaxf(m0.w, m1.w), fmaxf(m2.w, m3.w));
              } else {
                const float4 rv = *reinterpret_cast<const float4*>(
                    sIn + (((size_t)g * p.IH + ty + p.dpt) * p.IW + tx + p.dpl) * p.KS + c);
                v.x += rv.x; v.y += rv.y; v.z += rv.z; v.w += rv.w;
              }
            }
          } else if (p.res_mode == 2) {
            const float* rbase = p.res + (size_t)b * p.res_istride;
            float* vv = &v.x;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              int cc = c + j;
              if (cc >= p.res_C) continue;
              float rv;
              if (p.res_pool) {
                rv = -INFINITY;
#pragma unroll
                for (int dy = 0; dy < 2; ++dy)
#pragma unroll
                  for (int dx = 0; dx < 2; ++dx) {
                    int ry = 2 * oy + dy, rx = 2 * ox + dx;
                    if (ry < p.res_H && rx < p.res_W) rv = fmaxf(rv, rbase[((size_t)ry * p.res_W + rx) * p.res_Cs + cc]);
                  }
              } else {
                rv = rbase[((size_t)oy * p.res_W + ox) * p.res_Cs + cc];
              }
              vv[j] += rv;
            }
          }
          v.x = act1(v.x, p.act, al.x); v.y = act1(v.y, p.act, al.y);
          v.z = act1(v.z, p.act, al.z); v.w = act1(v.w, p.act, al.w);
          store_quad(p.out + (size_t)b * p.out_istride + ((size_t)oy * p.OW + ox) * p.CoutS + c, c, p.Cout, p.CoutS, p.vec_store, v);
        }
      }
    }
  }
}

// Raises a kernel's dynamic shared-memory limit once per (device, kernel, size high-water mark).
void set_smem(const void* kernel, size_t bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, size_t> cur;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(mu);
  size_t& c = cur[{dev, kernel}];
  if (bytes > c) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    c = bytes;
  }
}

}  // namespace

void launch_stem(const StemP& p, int B, cudaStream_t s, int max_ctas) {
  const int tiles = ((p.OW + 15) / 16) * ((p.OH + 7) / 8);
  int ntiles = tiles * B;
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  int nt = 32 * (p.NC >> 2);
  if (p.kw == 5) {
    set_smem((const void*)k_stem<5>, p.smem_bytes);
    k_stem<5><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  } else {
    set_smem((const void*)k_stem<3>, p.smem_bytes);
    k_stem<3><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  }
}

void launch_gemm_conv(const GemmConvP& p, int B, cudaStream_t s, int max_ctas) {
  const int P = p.TM * p.NPG;
  long long total_px = (long long)B * p.OH * p.OW;
  int ntiles = (int)((total_px + P - 1) / P);
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  // few pixel tiles and many channel chunks: give each chunk its own CTAs (at least ~2 waves of the 148 SMs in total)
  int gy = 1;
  if (p.nchunks > 1 && grid < 148) gy = std::min(p.nchunks, (296 + grid - 1) / grid);
  int nt = p.NPG * (p.NC >> 2);
  if (nt < 128) nt = 128;               // extra threads help with the im2col staging
  nt = (nt + 31) / 32 * 32;
  if (p.TM == 8) {
    set_smem((const void*)k_gemm_conv<8>, p.smem_bytes);
    k_gemm_conv<8><<<dim3(grid, gy), nt, p.smem_bytes, s>>>(p, B, ntiles);
  } else {
    set_smem((const void*)k_gemm_conv<4>, p.smem_bytes);
    k_gemm_conv<4><<<dim3(grid, gy), nt, p.smem_bytes, s>>>(p, B, ntiles);
  }
}

void launch_dwpw(const DwPwP& p, int B, cudaStream_t s, int max_ctas) {
  int groups = (B + p.G - 1) / p.G;
  int ntiles = groups * p.tilesX * p.tilesY;
  int grid = ntiles < max_ctas ? ntiles : max_ctas;
  if (grid < 1) grid = 1;
  int nt = p.NPG * (p.NC >> 2);
  if (p.TM == 8) {
    set_smem((const void*)k_dwpw<8>, p.smem_bytes);
    k_dwpw<8><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  } else {
    set_smem((const void*)k_dwpw<4>, p.smem_bytes);
    k_dwpw<4><<<grid, nt, p.smem_bytes, s>>>(p, B, ntiles);
  }
}

}  // namespace fdt
