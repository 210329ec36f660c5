// Scalar geometry / decode arithmetic shared by host code and device kernels.  Every function
// restates a piece of the reference's Dart glue with its f32/f64 mix (Dart double == f64).
#pragma once
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define FDT_HD __host__ __device__ __forceinline__
#else
#define FDT_HD inline
#endif

namespace fdt {

// Dart double.round(): half away from zero.
FDT_HD long long dart_round(double x) { return (long long)(x >= 0 ? floor(x + 0.5) : -floor(-x + 0.5)); }

// flutter_litert sigmoidClipped (call site lib/src/models/face_detection_model.dart:485).
FDT_HD double sigmoid_clipped(double x, double limit) {
  x = x < -limit ? -limit : (x > limit ? limit : x);
  return 1.0 / (1.0 + exp(-x));
}

// _decodeBoxesForIndices (lib/src/models/face_detection_model.dart:431-467): `tmp` is a
// Float32List, so every store rounds to f32 while each operation is evaluated in f64.
// raw: 16 f32; out box[4] (xmin,ymin,xmax,ymax) and kp[12] in f64.
FDT_HD void decode_box(const float* raw, double ax, double ay, double scale, double* box, double* kp) {
  float tmp[16];
  for (int j = 0; j < 16; ++j) tmp[j] = (float)((double)raw[j] / scale);
  tmp[0] = (float)((double)tmp[0] + ax);
  tmp[1] = (float)((double)tmp[1] + ay);
  for (int j = 4; j < 16; j += 2) {
    tmp[j] = (float)((double)tmp[j] + ax);
    tmp[j + 1] = (float)((double)tmp[j + 1] + ay);
  }
  double xc = tmp[0], yc = tmp[1], w = tmp[2], h = tmp[3];
  box[0] = xc - w * 0.5;
  box[1] = yc - h * 0.5;
  box[2] = xc + w * 0.5;
  box[3] = yc + h * 0.5;
  for (int j = 0; j < 12; ++j) kp[j] = (double)tmp[4 + j];
}

// IoU used by weightedNms (flutter_litert; MediaPipe OverlapSimilarity).
FDT_HD double box_iou(const double* a, const double* b) {
  double iw = fmin(a[2], b[2]) - fmax(a[0], b[0]);
  double ih = fmin(a[3], b[3]) - fmax(a[1], b[1]);
  if (!(iw > 0) || !(ih > 0)) return 0.0;
  double inter = iw * ih;
  double uni = (a[2] - a[0]) * (a[3] - a[1]) + (b[2] - b[0]) * (b[3] - b[1]) - inter;
  return uni > 0 ? inter / uni : 0.0;
}

// boxVisibleWidthFraction (lib/src/shared/face_gates.dart:115-121).
FDT_HD double visible_width_fraction(double xmin, double xmax, double image_width) {
  if (image_width <= 0) return 0.0;
  double left = xmin * image_width, right = xmax * image_width;
  double vis = fmin(right, image_width) - fmax(left, 0.0);
  return vis > 0 ? vis / image_width : 0.0;
}

// computeFaceAlignment (lib/src/shared/face_geometry.dart:17-45). kp: 12 normalised values.
FDT_HD void face_alignment(const double* kp, double img_w, double img_h, double* theta, double* cx,
                           double* cy, double* size) {
  double lx = kp[0] * img_w, ly = kp[1] * img_h;   // leftEye  (index 0)
  double rx = kp[2] * img_w, ry = kp[3] * img_h;   // rightEye (index 1)
  double mx = kp[6] * img_w, my = kp[7] * img_h;   // mouth    (index 3)
  double ecx = (lx + rx) * 0.5, ecy = (ly + ry) * 0.5;
  double vex = rx - lx, vey = ry - ly;
  double vmx = mx - ecx, vmy = my - ecy;
  *theta = atan2(vey, vex);
  double eye = sqrt(vex * vex + vey * vey);
  double mouth = sqrt(vmx * vmx + vmy * vmy);
  *size = fmax(mouth * 3.6, eye * 4.0);
  *cx = ecx + vmx * 0.1;
  *cy = ecy + vmy * 0.1;
}

// extractAlignedSquare (lib/src/util/helpers.dart:583-625): forward matrix of
// cv::getRotationMatrix2D (centre rounded through Point2f) with the pixel-centre translation,
// then cv::warpAffine's inversion.  theta_arg is the `theta` parameter of extractAlignedSquare
// (the face path passes -theta_face).  Returns false when round(size) <= 0.
FDT_HD bool aligned_square_inverse(double cx, double cy, double size, double theta_arg, int out_size,
                                   double* A /*6: a00,a01,b0,a10,a11,b1*/) {
  long long si = dart_round(size);
  if (!(si > 0)) return false;
  double sc = (double)out_size / (double)si;
  const double kPi = 3.141592653589793;
  double angle_deg = -theta_arg * 180.0 / kPi;
  double ang = angle_deg * kPi / 180.0;
  double a = sc * cos(ang), b = sc * sin(ang);
  double cxf = (double)(float)cx, cyf = (double)(float)cy;
  double m00 = a, m01 = b, m02 = (1 - a) * cxf - b * cyf;
  double m10 = -b, m11 = a, m12 = b * cxf + (1 - a) * cyf;
  double oc = out_size / 2.0 + 0.5 * (sc - 1.0);
  m02 += oc - cx;
  m12 += oc - cy;
  double D = m00 * m11 - m01 * m10;
  D = D != 0 ? 1.0 / D : 0.0;
  double A11 = m11 * D, A22 = m00 * D;
  double i00 = A11, i01 = m01 * (-D), i10 = m10 * (-D), i11 = A22;
  A[0] = i00; A[1] = i01; A[2] = -i00 * m02 - i01 * m12;
  A[3] = i10; A[4] = i11; A[5] = -i10 * m02 - i11 * m12;
  return true;
}

// 15-bit bilinear weights of cv::warpAffine's interpolation table for fractional position
// (fx, fy) in 1/32 pixel units (imgproc/imgwarp.cpp initInterTab2D, INTER_LINEAR).
FDT_HD void warp_weights(int fxi, int fyi, int* iw) {
  float fx = (float)fxi / 32.0f, fy = (float)fyi / 32.0f;
  float w[4] = {(1.0f - fy) * (1.0f - fx), (1.0f - fy) * fx, fy * (1.0f - fx), fy * fx};
  int sum = 0;
  for (int k = 0; k < 4; ++k) {
    iw[k] = (int)rintf(w[k] * 32768.0f);
    sum += iw[k];
  }
  int diff = 32768 - sum;
  if (diff != 0) {
    int best = 0;
    if (diff < 0) { for (int k = 1; k < 4; ++k) if (iw[k] > iw[best]) best = k; }
    else          { for (int k = 1; k < 4; ++k) if (iw[k] < iw[best]) best = k; }
    iw[best] += diff;
  }
}

struct LetterboxParams { int new_w, new_h, pad_top, pad_bottom, pad_left, pad_right; };

// computeLetterboxParams (flutter_litert; call site lib/src/util/helpers.dart:312-317).
inline LetterboxParams letterbox_params(int sw, int sh, int dw, int dh) {
  LetterboxParams p;
  double sx = (double)dw / sw, sy = (double)dh / sh;
  double scale = sx < sy ? sx : sy;
  long long nw = dart_round(sw * scale), nh = dart_round(sh * scale);
  if (nw < 1) nw = 1;
  if (nh < 1) nh = 1;
  if (nw > dw) nw = dw;
  if (nh > dh) nh = dh;
  p.new_w = (int)nw;
  p.new_h = (int)nh;
  p.pad_left = (dw - p.new_w) / 2;
  p.pad_top = (dh - p.new_h) / 2;
  p.pad_right = dw - p.new_w - p.pad_left;
  p.pad_bottom = dh - p.new_h - p.pad_top;
  return p;
}

// cv::resize INTER_LINEAR 8U tap tables for one axis (imgproc/resize.cpp): index pair and
// 11-bit weights.  clamp_fraction: x axis semantics (fraction forced to 0 at the borders).
inline void resize_linear_taps(int src, int dst, bool clamp_fraction, int* i0, int* i1, short* w0, short* w1) {
  double scale = (double)src / (double)dst;
  for (int d = 0; d < dst; ++d) {
    float f = (float)((d + 0.5) * scale - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    int a, b;
    if (clamp_fraction) {
      if (s < 0) { f = 0; s = 0; }
      if (s >= src - 1) { f = 0; s = src - 1; }
      a = s;
      b = s + 1 < src - 1 ? s + 1 : src - 1;
    } else {
      a = s < 0 ? 0 : (s > src - 1 ? src - 1 : s);
      b = s + 1 < 0 ? 0 : (s + 1 > src - 1 ? src - 1 : s + 1);
    }
    i0[d] = a;
    i1[d] = b;
    w0[d] = (short)lrintf((1.0f - f) * 2048.0f);
    w1[d] = (short)lrintf(f * 2048.0f);
  }
}

}  // namespace fdt
