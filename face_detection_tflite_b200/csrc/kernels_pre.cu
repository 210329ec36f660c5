// Preprocessing kernels: fused letterbox (cv::resize INTER_LINEAR 8U fixed point +
// copyMakeBorder(0) + BGRA/GRAY->BGR) and the [-1,1] normalisation.
// Reference: convertImageToTensor / bgrMatToSignedFloat32 (lib/src/util/helpers.dart:303-421).
#include "kernels.h"

namespace fdt {
namespace {

__device__ __forceinline__ void load_bgr(const uint8_t* px, int channels, int* v) {
  if (channels == 1) { v[0] = v[1] = v[2] = px[0]; }
  else { v[0] = px[0]; v[1] = px[1]; v[2] = px[2]; }
}

// One thread per output pixel: four 3-byte taps (byte loads; neighbouring threads share the sectors through L1) and one uchar4
// store; grid = (columns / 64, rows / 4, frames), so no index is divided (the flat 64-bit index this kernel started with cost two
// 64-bit divisions per pixel: 125 -> 122 us per 1024 1280x720 frames).
// Measured alternatives (round 2), both SLOWER on B200:
//  * the 6 bytes of the two neighbouring taps of a row from two aligned 32-bit loads + a funnel shift instead of six byte loads
//    (a third of the load instructions, same sectors): 168-172 us vs 122;
//  * a row-staged kernel - one warp per output row, the two source rows fetched with coalesced 16-byte streaming loads into shared
//    memory, taps read from there: 148 vs 126 us (97 vs 73 us per 256 1920x1080 frames).
// The DRAM traffic is the same in all three (every 32-byte sector of the 2-in-10 source rows) and this form already sits at 68 % of
// peak DRAM throughput.
__global__ void __launch_bounds__(256) k_letterbox(LetterboxP p, int B) {
  const int x = blockIdx.x * 64 + (threadIdx.x & 63), y = blockIdx.y * 4 + (threadIdx.x >> 6);
  if (x >= p.dst_w || y >= p.dst_h) return;
  for (int b = blockIdx.z; b < B; b += gridDim.z) {
    const int dy = y - p.pad_top, dx = x - p.pad_left;
    uchar3 o = make_uchar3(0, 0, 0);
    if (dy >= 0 && dy < p.new_h && dx >= 0 && dx < p.new_w) {
      const uint8_t* src = p.frames + (size_t)b * p.frame_stride;
      if (p.identity) {
        int v[3];
        load_bgr(src + (size_t)dy * p.row_stride + (size_t)dx * p.channels, p.channels, v);
        o = make_uchar3((unsigned char)v[0], (unsigned char)v[1], (unsigned char)v[2]);
      } else {
        const int xa = p.x0[dx], xb = p.x1[dx], wa = p.ax0[dx], wb = p.ax1[dx];
        const int ya = p.y0[dy], yb = p.y1[dy], va = p.by0[dy], vb = p.by1[dy];
        int t00[3], t01[3], t10[3], t11[3];
        const uint8_t* r0 = src + (size_t)ya * p.row_stride;
        const uint8_t* r1 = src + (size_t)yb * p.row_stride;
        load_bgr(r0 + (size_t)xa * p.channels, p.channels, t00);
        load_bgr(r0 + (size_t)xb * p.channels, p.channels, t01);
        load_bgr(r1 + (size_t)xa * p.channels, p.channels, t10);
        load_bgr(r1 + (size_t)xb * p.channels, p.channels, t11);
        int res[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          // HResizeLinear (11-bit) then VResizeLinear<uchar,int,short> fixed point (resize.cpp)
          int h0 = t00[c] * wa + t01[c] * wb;
          int h1 = t10[c] * wa + t11[c] * wb;
          int v = ((va * (h0 >> 4)) >> 16) + ((vb * (h1 >> 4)) >> 16);
          v = (v + 2) >> 2;
          res[c] = v < 0 ? 0 : (v > 255 ? 255 : v);
        }
        o = make_uchar3((unsigned char)res[0], (unsigned char)res[1], (unsigned char)res[2]);
      }
    }
    reinterpret_cast<uchar4*>(p.out)[((size_t)b * p.dst_h + y) * p.dst_w + x] = make_uchar4(o.x, o.y, o.z, 0);
  }
}

__global__ void k_normalize(const uint8_t* in, TV out, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % out.Cs);
    long long pix = idx / out.Cs;
    long long hw = (long long)out.H * out.W;
    int b = (int)(pix / hw);
    long long sp = pix % hw;
    float v = 0.f;
    if (c < 3) v = fmaf((float)in[((size_t)b * hw + sp) * 4 + (2 - c)], 1.0f / 127.5f, -1.0f);
    out.p[b * out.istride + sp * out.Cs + c] = v;
  }
}

}  // namespace

void launch_letterbox(const LetterboxP& p, int B, cudaStream_t s) {
  if (B <= 0) return;
  dim3 grid((unsigned)((p.dst_w + 63) / 64), (unsigned)((p.dst_h + 3) / 4), (unsigned)(B < 65535 ? B : 65535));
  k_letterbox<<<grid, 256, 0, s>>>(p, B);
}

void launch_normalize(const uint8_t* in, TV out, int B, cudaStream_t s) {
  long long total = (long long)B * out.H * out.W * out.Cs;
  long long g = (total + 255) / 256;
  if (g > 148LL * 32) g = 148LL * 32;
  k_normalize<<<(int)(g < 1 ? 1 : g), 256, 0, s>>>(in, out, total);
}

}  // namespace fdt
