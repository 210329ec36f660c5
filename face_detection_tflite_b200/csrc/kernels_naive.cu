// One-thread-per-output-element kernels with plain TFLite op semantics.  They execute the graph
// op by op at fuse_level 0 (the device-side parity reference for every intermediate tensor) and
// cover shapes the tiled kernels do not take.  XNNPACK equivalents: conv2d / dwconv2d / add /
// clamp / prelu / max-pool / resize-bilinear / constant-pad.
#include "kernels.h"

namespace fdt {
namespace {

__device__ __forceinline__ float act_fn(float v, int act, const float* alpha, int c) {
  if (act == kActRelu) return fmaxf(v, 0.f);
  if (act == kActPrelu) return v >= 0.f ? v : v * alpha[c];
  return v;
}

__global__ void k_naive_conv(NaiveConvP p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int co = (int)(idx % p.out.Cs);
    long long r = idx / p.out.Cs;
    int ox = (int)(r % p.out.W); r /= p.out.W;
    int oy = (int)(r % p.out.H);
    int b = (int)(r / p.out.H);
    float* dst = p.out.p + b * p.out.istride + ((long long)oy * p.out.W + ox) * p.out.Cs + co;
    if (co >= p.out.C) { *dst = 0.f; continue; }
    const float* src = p.in.p + b * p.in.istride;
    float acc = p.bias ? p.bias[co] : 0.f;
    for (int ky = 0; ky < p.kh; ++ky) {
      int iy = oy * p.sh + ky - p.pt;
      if (iy < 0 || iy >= p.in.H) continue;
      for (int kx = 0; kx < p.kw; ++kx) {
        int ix = ox * p.sw + kx - p.pl;
        if (ix < 0 || ix >= p.in.W) continue;
        const float* px = src + ((long long)iy * p.in.W + ix) * p.in.Cs;
        if (p.depthwise) {
          acc = fmaf(px[co], p.w[(ky * p.kw + kx) * p.out.C + co], acc);
        } else {
          const float* wr = p.w + ((long long)(co * p.kh + ky) * p.kw + kx) * p.in.C;
          for (int c = 0; c < p.in.C; ++c) acc = fmaf(px[c], wr[c], acc);
        }
      }
    }
    *dst = act_fn(acc, p.act, p.alpha, co);
  }
}

__global__ void k_add(EltP p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.out.Cs);
    long long pix = idx / p.out.Cs;
    long long hw = (long long)p.out.H * p.out.W;
    int b = (int)(pix / hw);
    long long sp = pix % hw;
    float* dst = p.out.p + b * p.out.istride + sp * p.out.Cs + c;
    if (c >= p.out.C) { *dst = 0.f; continue; }
    float v = p.a.p[b * p.a.istride + sp * p.a.Cs + c] + p.b.p[b * p.b.istride + sp * p.b.Cs + c];
    *dst = act_fn(v, p.act, p.alpha, c);
  }
}

__global__ void k_act(EltP p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.out.Cs);
    long long pix = idx / p.out.Cs;
    long long hw = (long long)p.out.H * p.out.W;
    int b = (int)(pix / hw);
    long long sp = pix % hw;
    float* dst = p.out.p + b * p.out.istride + sp * p.out.Cs + c;
    if (c >= p.out.C) { *dst = 0.f; continue; }
    *dst = act_fn(p.a.p[b * p.a.istride + sp * p.a.Cs + c], p.act, p.alpha, c);
  }
}

// channel zero-pad [[0,0],[0,0],[0,0],[0,d]] (the only PAD form on the path)
__global__ void k_padc(EltP p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.out.Cs);
    long long pix = idx / p.out.Cs;
    long long hw = (long long)p.out.H * p.out.W;
    int b = (int)(pix / hw);
    long long sp = pix % hw;
    float v = c < p.a.C ? p.a.p[b * p.a.istride + sp * p.a.Cs + c] : 0.f;
    p.out.p[b * p.out.istride + sp * p.out.Cs + c] = v;
  }
}

__global__ void k_maxpool(PoolP p, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.out.Cs);
    long long r = idx / p.out.Cs;
    int ox = (int)(r % p.out.W); r /= p.out.W;
    int oy = (int)(r % p.out.H);
    int b = (int)(r / p.out.H);
    float* dst = p.out.p + b * p.out.istride + ((long long)oy * p.out.W + ox) * p.out.Cs + c;
    if (c >= p.out.C) { *dst = 0.f; continue; }
    float m = -INFINITY;
    for (int ky = 0; ky < p.fh; ++ky) {
      int iy = oy * p.sh + ky - p.pt;
      if (iy < 0 || iy >= p.in.H) continue;
      for (int kx = 0; kx < p.fw; ++kx) {
        int ix = ox * p.sw + kx - p.pl;
        if (ix < 0 || ix >= p.in.W) continue;
        m = fmaxf(m, p.in.p[b * p.in.istride + ((long long)iy * p.in.W + ix) * p.in.Cs + c]);
      }
    }
    *dst = m;
  }
}

// TFLite RESIZE_BILINEAR (reference kernel resize_bilinear.h): fp32 lerp, x first then y.
__global__ void k_resize_bilinear(ResizeP p, long long total) {
  float sy_scale = (p.align_corners && p.out.H > 1) ? (float)(p.in.H - 1) / (p.out.H - 1) : (float)p.in.H / p.out.H;
  float sx_scale = (p.align_corners && p.out.W > 1) ? (float)(p.in.W - 1) / (p.out.W - 1) : (float)p.in.W / p.out.W;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % p.out.Cs);
    long long r = idx / p.out.Cs;
    int ox = (int)(r % p.out.W); r /= p.out.W;
    int oy = (int)(r % p.out.H);
    int b = (int)(r / p.out.H);
    float* dst = p.out.p + b * p.out.istride + ((long long)oy * p.out.W + ox) * p.out.Cs + c;
    if (c >= p.out.C) { *dst = 0.f; continue; }
    float fy = p.half_pixel ? (oy + 0.5f) * sy_scale - 0.5f : oy * sy_scale;
    float fx = p.half_pixel ? (ox + 0.5f) * sx_scale - 0.5f : ox * sx_scale;
    float y0f = floorf(fy), x0f = floorf(fx);
    int y0 = max(0, min((int)y0f, p.in.H - 1)), y1 = max(0, min((int)y0f + 1, p.in.H - 1));
    int x0 = max(0, min((int)x0f, p.in.W - 1)), x1 = max(0, min((int)x0f + 1, p.in.W - 1));
    float wy = fy - y0f, wx = fx - x0f;
    const float* src = p.in.p + b * p.in.istride + c;
    float v00 = src[((long long)y0 * p.in.W + x0) * p.in.Cs], v01 = src[((long long)y0 * p.in.W + x1) * p.in.Cs];
    float v10 = src[((long long)y1 * p.in.W + x0) * p.in.Cs], v11 = src[((long long)y1 * p.in.W + x1) * p.in.Cs];
    float top = v00 * (1.f - wx) + v01 * wx;
    float bot = v10 * (1.f - wx) + v11 * wx;
    *dst = top * (1.f - wy) + bot * wy;
  }
}

// The FPN up-path of the full-range detector: out = RESIZE_BILINEAR(in) + skip [ReLU] in one pass, four channels per thread
// (same fp32 lerp order as above, then the ADD).  Saves the round trip of the upsampled tensor through HBM.
__global__ void k_resize_add4(ResizeP p, long long total4) {
  float sy_scale = (p.align_corners && p.out.H > 1) ? (float)(p.in.H - 1) / (p.out.H - 1) : (float)p.in.H / p.out.H;
  float sx_scale = (p.align_corners && p.out.W > 1) ? (float)(p.in.W - 1) / (p.out.W - 1) : (float)p.in.W / p.out.W;
  const int q4 = p.out.Cs >> 2;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total4; idx += (long long)gridDim.x * blockDim.x) {
    int c = (int)(idx % q4) * 4;
    long long r = idx / q4;
    int ox = (int)(r % p.out.W); r /= p.out.W;
    int oy = (int)(r % p.out.H);
    int b = (int)(r / p.out.H);
    float fy = p.half_pixel ? (oy + 0.5f) * sy_scale - 0.5f : oy * sy_scale;
    float fx = p.half_pixel ? (ox + 0.5f) * sx_scale - 0.5f : ox * sx_scale;
    float y0f = floorf(fy), x0f = floorf(fx);
    int y0 = max(0, min((int)y0f, p.in.H - 1)), y1 = max(0, min((int)y0f + 1, p.in.H - 1));
    int x0 = max(0, min((int)x0f, p.in.W - 1)), x1 = max(0, min((int)x0f + 1, p.in.W - 1));
    float wy = fy - y0f, wx = fx - x0f;
    const float* src = p.in.p + b * p.in.istride + c;
    const float4 v00 = *reinterpret_cast<const float4*>(src + ((long long)y0 * p.in.W + x0) * p.in.Cs);
    const float4 v01 = *reinterpret_cast<const float4*>(src + ((long long)y0 * p.in.W + x1) * p.in.Cs);
    const float4 v10 = *reinterpret_cast<const float4*>(src + ((long long)y1 * p.in.W + x0) * p.in.Cs);
    const float4 v11 = *reinterpret_cast<const float4*>(src + ((long long)y1 * p.in.W + x1) * p.in.Cs);
    const long long o = b * p.out.istride + ((long long)oy * p.out.W + ox) * p.out.Cs + c;
    const float4 a = *reinterpret_cast<const float4*>(p.add.p + b * p.add.istride + ((long long)oy * p.out.W + ox) * p.add.Cs + c);
    auto lerp = [&](float t00, float t01, float t10, float t11, float ad) {
      float top = t00 * (1.f - wx) + t01 * wx;
      float bot = t10 * (1.f - wx) + t11 * wx;
      float v = ad + (top * (1.f - wy) + bot * wy);      // ADD(in0 = skip, in1 = resized)
      return p.act == kActRelu ? fmaxf(v, 0.f) : v;
    };
    float4 out;
    out.x = lerp(v00.x, v01.x, v10.x, v11.x, a.x); out.y = lerp(v00.y, v01.y, v10.y, v11.y, a.y);
    out.z = lerp(v00.z, v01.z, v10.z, v11.z, a.z); out.w = lerp(v00.w, v01.w, v10.w, v11.w, a.w);
    *reinterpret_cast<float4*>(p.out.p + o) = out;
  }
}

inline int grid_for(long long total, int block) {
  long long g = (total + block - 1) / block;
  long long cap = 148LL * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

void launch_naive_conv(const NaiveConvP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  k_naive_conv<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}
void launch_add(const EltP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  k_add<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}
void launch_act(const EltP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  k_act<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}
void launch_padc(const EltP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  k_padc<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}
void launch_maxpool(const PoolP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  k_maxpool<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}
void launch_resize_bilinear(const ResizeP& p, int B, cudaStream_t s) {
  long long total = (long long)B * p.out.H * p.out.W * p.out.Cs;
  if (p.has_add) k_resize_add4<<<grid_for(total / 4, 256), 256, 0, s>>>(p, total / 4);
  else k_resize_bilinear<<<grid_for(total, 256), 256, 0, s>>>(p, total);
}

}  // namespace fdt
