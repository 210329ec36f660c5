// Host half of the JPEG front end (detectFacesFromBytes, /root/reference/lib/src/face_detector.dart:477-485 ->
// cv.imdecode): marker parsing and Huffman entropy decoding (baseline sequential and progressive, 8-bit, 1 or 3
// components) into quantised DCT coefficient planes.  Everything after the entropy decoder - dequantisation, the
// integer IDCT, chroma upsampling, colour conversion, EXIF orientation - runs on the device (kernels_jpeg.cu) and
// reproduces libjpeg-turbo's default decompression path (what cv::imdecode runs) bit for bit.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace fdt {

struct JpegComp {
  int id = 0, h = 1, v = 1, tq = 0;
  int dw = 0, dh = 0;        // component size in samples: ceil(width * h / hmax), ceil(height * v / vmax)
  int bw = 0, bh = 0;        // blocks per row / column, padded to whole MCUs
  std::vector<int16_t> coef; // [bh][bw][64] quantised coefficients, natural (row-major) order
};

struct JpegImage {
  int width = 0, height = 0, ncomp = 0, hmax = 1, vmax = 1, mcux = 0, mcuy = 0;
  bool progressive = false;
  int orientation = 1;       // EXIF tag 0x0112 (1..8); cv::imdecode applies it
  JpegComp comp[3];
  uint16_t qt[4][64];        // natural order
  bool qt_set[4] = {false, false, false, false};
};

enum JpegStatus { kJpegOk = 0, kJpegBad = 1, kJpegUnsupported = 2 };

// Parses and entropy-decodes `data`.  kJpegBad: not a decodable JPEG (the caller reports FormatException like the reference);
// kJpegUnsupported: a valid stream of a kind this decoder does not take (arithmetic coding, 12-bit, CMYK / 4 components,
// lossless, hierarchical).
JpegStatus jpeg_decode_coefficients(const uint8_t* data, size_t n, JpegImage* img, std::string* err);

}  // namespace fdt
