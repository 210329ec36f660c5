// Minimal bounds-checked TFLite (schema v3) flatbuffer reader.  Replaces what the reference gets
// from flutter_litert's Interpreter.fromBuffer (reference: lib/src/models/face_detection_model.dart:156-191).
// Field slots follow tensorflow/lite/schema/schema.fbs.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace fdt {

enum TfOpCode {
  kOpAdd = 0, kOpAvgPool = 1, kOpConcat = 2, kOpConv2D = 3, kOpDwConv2D = 4, kOpDequantize = 6,
  kOpMaxPool = 17, kOpRelu = 19, kOpReshape = 22, kOpResizeBilinear = 23, kOpPad = 34, kOpPrelu = 54,
};
enum TfType { kTfF32 = 0, kTfF16 = 1, kTfI32 = 2 };

struct TfTensor {
  std::vector<int> shape;
  int dtype = 0;
  std::string name;
  const uint8_t* data = nullptr;  // points into TfModel::blob
  size_t nbytes = 0;
  int dim(int i) const { return i < (int)shape.size() ? shape[i] : 1; }
  long long numel() const { long long n = 1; for (int d : shape) n *= d; return n; }
};

struct TfOp {
  int code = -1;
  std::vector<int> in, out;
  int padding = 0;  // 0 SAME, 1 VALID
  int stride_w = 1, stride_h = 1, dil_w = 1, dil_h = 1, act = 0, depth_mult = 1;
  int filter_w = 1, filter_h = 1, align_corners = 0, half_pixel = 0, axis = 0;
};

struct TfModel {
  std::vector<uint8_t> blob;
  std::vector<TfTensor> tensors;
  std::vector<TfOp> ops;
  std::vector<int> inputs, outputs;

  bool parse(const uint8_t* data, size_t len, std::string* err);
  // Constant tensor as float32 (f16 payloads widened exactly; DEQUANTIZE outputs resolved).
  bool const_f32(int tensor, std::vector<float>* out) const;
  bool const_i32(int tensor, std::vector<int>* out) const;
  int producer(int tensor) const;                 // op index producing `tensor`, or -1
  std::vector<int> consumers(int tensor) const;   // op indices reading `tensor`
};

float half_to_float(uint16_t h);

}  // namespace fdt
