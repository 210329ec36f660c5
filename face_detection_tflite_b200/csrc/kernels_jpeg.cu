// Device half of the JPEG front end: quantised DCT coefficients -> BGR frame, bit-exact with libjpeg-turbo's default
// decompression path (the one cv::imdecode runs; reference call site /root/reference/lib/src/face_detector.dart:477-485):
//   k_jpeg_idct   dequantisation + the "slow integer" 8x8 inverse DCT (jidctint.c jpeg_idct_islow: 13-bit constants,
//                 2 extra bits between the passes, range limit = +128 and clamp), one thread per block;
//   k_jpeg_color  "fancy" (triangle filter) chroma upsampling for 2x1 / 2x2 / 1x2 subsampling (jdsample.c h2v1 / h2v2 / h1v2
//                 fancy upsample, with libjpeg's edge rules), fixed-point YCbCr -> RGB (jdcolor.c: 16-bit tables written
//                 out as integer arithmetic), EXIF orientation (OpenCV's ExifTransform) folded into the store address.
// Integer work throughout: results equal cv2.imdecode byte for byte (tests/test_gpu_jpeg.py).
#include "kernels.h"

namespace fdt {
namespace {

constexpr int CONST_BITS = 13, PASS1_BITS = 2;
constexpr int FIX_0_298631336 = 2446, FIX_0_390180644 = 3196, FIX_0_541196100 = 4433, FIX_0_765366865 = 6270,
              FIX_0_899976223 = 7373, FIX_1_175875602 = 9633, FIX_1_501321110 = 12299, FIX_1_847759065 = 15137,
              FIX_1_961570560 = 16069, FIX_2_053119869 = 16819, FIX_2_562915447 = 20995, FIX_3_072711026 = 25172;

__device__ __forceinline__ int descale(int x, int n) { return (x + (1 << (n - 1))) >> n; }
// libjpeg's range_limit table indexed with (x & RANGE_MASK): +128, clamp to [0, 255]; far out-of-range values wrap as in the table
__device__ __forceinline__ int range_limit(int x) {
  x &= 1023;
  return x < 128 ? x + 128 : (x < 512 ? 255 : (x < 896 ? 0 : x - 896));
}

// one 1-D pass of the LL&M inverse DCT on eight values (already dequantised / from the workspace)
__device__ __forceinline__ void idct8(const int (&in)[8], int (&o)[8], int shift) {
  int z2 = in[2], z3 = in[6];
  int z1 = (z2 + z3) * FIX_0_541196100;
  int tmp2 = z1 + z3 * (-FIX_1_847759065);
  int tmp3 = z1 + z2 * FIX_0_765366865;
  z2 = in[0]; z3 = in[4];
  int tmp0 = (z2 + z3) * (1 << CONST_BITS);
  int tmp1 = (z2 - z3) * (1 << CONST_BITS);
  const int tmp10 = tmp0 + tmp3, tmp13 = tmp0 - tmp3, tmp11 = tmp1 + tmp2, tmp12 = tmp1 - tmp2;
  tmp0 = in[7]; tmp1 = in[5]; tmp2 = in[3]; tmp3 = in[1];
  z1 = tmp0 + tmp3; z2 = tmp1 + tmp2; z3 = tmp0 + tmp2;
  int z4 = tmp1 + tmp3;
  const int z5 = (z3 + z4) * FIX_1_175875602;
  tmp0 *= FIX_0_298631336; tmp1 *= FIX_2_053119869; tmp2 *= FIX_3_072711026; tmp3 *= FIX_1_501321110;
  z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447; z3 *= -FIX_1_961570560; z4 *= -FIX_0_390180644;
  z3 += z5; z4 += z5;
  tmp0 += z1 + z3; tmp1 += z2 + z4; tmp2 += z2 + z3; tmp3 += z1 + z4;
  o[0] = descale(tmp10 + tmp3, shift); o[7] = descale(tmp10 - tmp3, shift);
  o[1] = descale(tmp11 + tmp2, shift); o[6] = descale(tmp11 - tmp2, shift);
  o[2] = descale(tmp12 + tmp1, shift); o[5] = descale(tmp12 - tmp1, shift);
  o[3] = descale(tmp13 + tmp0, shift); o[4] = descale(tmp13 - tmp0, shift);
}

__global__ void __launch_bounds__(128) k_jpeg_idct(JpegIdctP p) {
  const int nb = p.bw * p.bh;
  for (int b = blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += gridDim.x * blockDim.x) {
    const int by = b / p.bw, bx = b - by * p.bw;
    const int4* src = reinterpret_cast<const int4*>(p.coef + (size_t)b * 64);
    int ws[64];
#pragma unroll
    for (int r = 0; r < 8; ++r) {                       // row r of the coefficient block: 8 int16
      const int4 v = __ldg(src + r);
      const int w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        ws[r * 8 + 2 * j] = (int)(short)(w[j] & 0xFFFF) * (int)p.q[r * 8 + 2 * j];
        ws[r * 8 + 2 * j + 1] = (int)(short)(w[j] >> 16) * (int)p.q[r * 8 + 2 * j + 1];
      }
    }
    // pass 1: columns
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      int in[8], o[8];
#pragma unroll
      for (int r = 0; r < 8; ++r) in[r] = ws[r * 8 + c];
      idct8(in, o, CONST_BITS - PASS1_BITS);
#pragma unroll
      for (int r = 0; r < 8; ++r) ws[r * 8 + c] = o[r];
    }
    // pass 2: rows, range limit, store
    uint8_t* dst = p.plane + (size_t)(by * 8) * p.pitch + bx * 8;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int in[8], o[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) in[c] = ws[r * 8 + c];
      idct8(in, o, CONST_BITS + PASS1_BITS + 3);
      uint32_t lo = 0, hi = 0;
#pragma unroll
      for (int c = 0; c < 4; ++c) { lo |= (uint32_t)range_limit(o[c]) << (8 * c); hi |= (uint32_t)range_limit(o[4 + c]) << (8 * c); }
      *reinterpret_cast<uint2*>(dst + (size_t)r * p.pitch) = make_uint2(lo, hi);
    }
  }
}

// chroma sample for output pixel (x, y) of a plane subsampled by (hs, vs) in {1, 2}: libjpeg's fancy upsampling
__device__ __forceinline__ int chroma_at(const uint8_t* pl, int pitch, int dw, int dh, int hs, int vs, int x, int y) {
  if (hs == 1 && vs == 1) return pl[(size_t)y * pitch + x];
  if (hs == 2 && vs == 1) {
    const uint8_t* row = pl + (size_t)y * pitch;
    const int cx = x >> 1, v = row[cx];
    if (dw <= 2) return v;                                                       // h2v1_upsample (replication)
    if ((x & 1) == 0) return cx == 0 ? v : (3 * v + row[cx - 1] + 1) >> 2;
    return cx == dw - 1 ? v : (3 * v + row[cx + 1] + 2) >> 2;
  }
  const int cy = y >> 1;
  const int oy = (y & 1) ? min(cy + 1, dh - 1) : max(cy - 1, 0);                 // the nearer neighbouring row (edge rows repeat)
  const uint8_t* r0 = pl + (size_t)cy * pitch;
  const uint8_t* r1 = pl + (size_t)oy * pitch;
  if (hs == 1) {                                                                 // h1v2_fancy_upsample
    return (3 * r0[x] + r1[x] + ((y & 1) ? 2 : 1)) >> 2;
  }
  const int cx = x >> 1;
  if (dw <= 2) return r0[cx];                                                    // h2v2_upsample (replication)
  const int cur = 3 * r0[cx] + r1[cx];
  if ((x & 1) == 0) {
    if (cx == 0) return (cur * 4 + 8) >> 4;
    return (cur * 3 + (3 * r0[cx - 1] + r1[cx - 1]) + 8) >> 4;
  }
  if (cx == dw - 1) return (cur * 4 + 7) >> 4;
  return (cur * 3 + (3 * r0[cx + 1] + r1[cx + 1]) + 7) >> 4;
}

__device__ __forceinline__ int clamp255(int v) { return min(max(v, 0), 255); }

__global__ void __launch_bounds__(256) k_jpeg_color(JpegColorP p) {
  const long long total = (long long)p.W * p.H;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / p.W), x = (int)(i - (long long)y * p.W);
    const int Y = p.py[(size_t)y * p.pitch_y + x];
    int r = Y, g = Y, b = Y;
    if (p.ncomp == 3) {
      const int cb = chroma_at(p.pcb, p.pitch_c, p.cdw, p.cdh, p.hs, p.vs, x, y) - 128;
      const int cr = chroma_at(p.pcr, p.pitch_c, p.cdw, p.cdh, p.hs, p.vs, x, y) - 128;
      r = clamp255(Y + ((91881 * cr + 32768) >> 16));
      b = clamp255(Y + ((116130 * cb + 32768) >> 16));
      g = clamp255(Y + ((-22554 * cb + 32768 - 46802 * cr) >> 16));
    }
    int ox, oy;
    switch (p.orientation) {
      default: ox = x; oy = y; break;
      case 2: ox = p.W - 1 - x; oy = y; break;
      case 3: ox = p.W - 1 - x; oy = p.H - 1 - y; break;
      case 4: ox = x; oy = p.H - 1 - y; break;
      case 5: ox = y; oy = x; break;
      case 6: ox = p.H - 1 - y; oy = x; break;
      case 7: ox = p.H - 1 - y; oy = p.W - 1 - x; break;
      case 8: ox = y; oy = p.W - 1 - x; break;
    }
    uint8_t* o = p.out + ((size_t)oy * p.out_w + ox) * 3;
    o[0] = (uint8_t)b; o[1] = (uint8_t)g; o[2] = (uint8_t)r;
  }
}

}  // namespace

void launch_jpeg_idct(const JpegIdctP& p, cudaStream_t s) {
  const int nb = p.bw * p.bh;
  if (nb <= 0) return;
  k_jpeg_idct<<<std::min((nb + 127) / 128, 148 * 16), 128, 0, s>>>(p);
}
void launch_jpeg_color(const JpegColorP& p, cudaStream_t s) {
  const long long total = (long long)p.W * p.H;
  if (total <= 0) return;
  k_jpeg_color<<<(int)std::min<long long>((total + 255) / 256, 148LL * 32), 256, 0, s>>>(p);
}

}  // namespace fdt
