#include "tflite_model.h"

#include <cstring>

namespace fdt {

float half_to_float(uint16_t h) {
  uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
  uint32_t exp = (h >> 10) & 0x1f;
  uint32_t man = h & 0x3ffu;
  uint32_t bits;
  if (exp == 0) {
    if (man == 0) {
      bits = sign;
    } else {  // subnormal: normalise
      int e = -1;
      do { man <<= 1; ++e; } while (!(man & 0x400u));
      bits = sign | ((uint32_t)(127 - 15 - e) << 23) | ((man & 0x3ffu) << 13);
    }
  } else if (exp == 31) {
    bits = sign | 0x7f800000u | (man << 13);
  } else {
    bits = sign | ((exp + 127 - 15) << 23) | (man << 13);
  }
  float f;
  std::memcpy(&f, &bits, 4);
  return f;
}

namespace {

struct Reader {
  const uint8_t* b;
  size_t n;
  bool ok = true;

  // every offset comes from untrusted model bytes: compare without forming o + size (which could wrap)
  template <typename T> T rd(size_t o) {
    if (o > n || sizeof(T) > n - o) { ok = false; return T(0); }
    T v;
    std::memcpy(&v, b + o, sizeof(T));
    return v;
  }
  size_t indirect(size_t o) {
    const uint32_t d = rd<uint32_t>(o);
    if (!ok || d > n - o) { ok = false; return n; }      // n: a position every later read rejects
    return o + d;
  }
  // absolute position of `slot` in table, 0 if absent
  size_t field(size_t table, int slot) {
    int32_t so = rd<int32_t>(table);
    if (!ok) return 0;
    const long long vtl = (long long)table - so;
    if (vtl < 0 || (unsigned long long)vtl > n || n - (size_t)vtl < 4) { ok = false; return 0; }   // vtable must lie inside the buffer
    size_t vt = (size_t)vtl;
    uint16_t vsize = rd<uint16_t>(vt);
    if (vsize < 4 || vsize > n - vt) { ok = false; return 0; }
    size_t e = 4 + 2 * (size_t)slot;
    if (e + 2 > vsize) return 0;
    uint16_t off = rd<uint16_t>(vt + e);
    if (off && off > n - table) { ok = false; return 0; }
    return off ? table + off : 0;
  }
  bool vec(size_t table, int slot, size_t* start, uint32_t* len) {
    size_t p = field(table, slot);
    if (!p) { *start = 0; *len = 0; return false; }
    size_t v = indirect(p);
    *len = rd<uint32_t>(v);
    if (!ok) { *start = 0; *len = 0; return false; }
    *start = v + 4;
    return ok;
  }
  std::vector<size_t> tables(size_t table, int slot) {
    size_t s; uint32_t l;
    std::vector<size_t> r;
    if (!vec(table, slot, &s, &l)) return r;
    if (s > n || 4ull * l > n - s) { ok = false; return r; }
    for (uint32_t i = 0; i < l; ++i) r.push_back(indirect(s + 4ull * i));
    return r;
  }
  std::vector<int> ints(size_t table, int slot) {
    size_t s; uint32_t l;
    std::vector<int> r;
    if (!vec(table, slot, &s, &l)) return r;
    if (s > n || 4ull * l > n - s) { ok = false; return r; }
    for (uint32_t i = 0; i < l; ++i) r.push_back(rd<int32_t>(s + 4ull * i));
    return r;
  }
  template <typename T> T scalar(size_t table, int slot, T dflt) {
    size_t p = field(table, slot);
    return p ? rd<T>(p) : dflt;
  }
};

}  // namespace

bool TfModel::parse(const uint8_t* data, size_t len, std::string* err) {
  auto fail = [&](const char* m) { if (err) *err = m; return false; };
  if (!data || len < 16) return fail("tflite buffer too small");
  blob.assign(data, data + len);
  Reader r{blob.data(), blob.size()};
  size_t root = r.indirect(0);
  std::vector<int> opcodes;
  for (size_t t : r.tables(root, 1)) {
    int dep = r.scalar<int8_t>(t, 0, 0);
    int neu = r.scalar<int32_t>(t, 3, 0);
    opcodes.push_back(dep > neu ? dep : neu);
  }
  struct Buf { size_t s; uint32_t l; };
  std::vector<Buf> bufs;
  for (size_t t : r.tables(root, 4)) {
    Buf b{0, 0};
    r.vec(t, 0, &b.s, &b.l);
    if (!r.ok || b.s > blob.size() || b.l > blob.size() - b.s) return fail("buffer out of range");
    bufs.push_back(b);
  }
  auto sgs = r.tables(root, 2);
  if (!r.ok || sgs.empty()) return fail("no subgraph");
  size_t sg = sgs[0];
  for (size_t t : r.tables(sg, 0)) {
    TfTensor x;
    x.shape = r.ints(t, 0);
    x.dtype = r.scalar<int8_t>(t, 1, 0);
    uint32_t bi = r.scalar<uint32_t>(t, 2, 0);
    size_t s; uint32_t l;
    if (r.vec(t, 3, &s, &l) && s <= blob.size() && l <= blob.size() - s) x.name.assign((const char*)blob.data() + s, l);
    if (bi < bufs.size() && bufs[bi].l) { x.data = blob.data() + bufs[bi].s; x.nbytes = bufs[bi].l; }
    // shapes size every device buffer downstream: positive dims, bounded element count, constant payload of the right size
    if (x.shape.size() > 6) return fail("tensor rank above 6");
    unsigned long long numel = 1;
    for (int d : x.shape) {
      if (d <= 0) return fail("tensor with a non-positive dimension");
      numel *= (unsigned long long)d;
      if (numel > (1ull << 31)) return fail("tensor too large");
    }
    if (x.data) {
      const unsigned long long esz = x.dtype == kTfF16 ? 2 : 4;
      if ((x.dtype == kTfF32 || x.dtype == kTfF16 || x.dtype == kTfI32) && x.nbytes != numel * esz) return fail("constant tensor payload does not match its shape");
    }
    tensors.push_back(std::move(x));
  }
  inputs = r.ints(sg, 1);
  outputs = r.ints(sg, 2);
  for (size_t t : r.tables(sg, 3)) {
    TfOp op;
    uint32_t oi = r.scalar<uint32_t>(t, 0, 0);
    if (oi >= opcodes.size()) return fail("bad opcode index");
    op.code = opcodes[oi];
    op.in = r.ints(t, 1);
    op.out = r.ints(t, 2);
    size_t p = r.field(t, 4);
    if (p) {
      size_t o = r.indirect(p);
      switch (op.code) {
        case kOpConv2D:
          op.padding = r.scalar<int8_t>(o, 0, 0); op.stride_w = r.scalar<int32_t>(o, 1, 1);
          op.stride_h = r.scalar<int32_t>(o, 2, 1); op.act = r.scalar<int8_t>(o, 3, 0);
          op.dil_w = r.scalar<int32_t>(o, 4, 1); op.dil_h = r.scalar<int32_t>(o, 5, 1);
          break;
        case kOpDwConv2D:
          op.padding = r.scalar<int8_t>(o, 0, 0); op.stride_w = r.scalar<int32_t>(o, 1, 1);
          op.stride_h = r.scalar<int32_t>(o, 2, 1); op.depth_mult = r.scalar<int32_t>(o, 3, 1);
          op.act = r.scalar<int8_t>(o, 4, 0); op.dil_w = r.scalar<int32_t>(o, 5, 1);
          op.dil_h = r.scalar<int32_t>(o, 6, 1);
          break;
        case kOpMaxPool: case kOpAvgPool:
          op.padding = r.scalar<int8_t>(o, 0, 0); op.stride_w = r.scalar<int32_t>(o, 1, 1);
          op.stride_h = r.scalar<int32_t>(o, 2, 1); op.filter_w = r.scalar<int32_t>(o, 3, 1);
          op.filter_h = r.scalar<int32_t>(o, 4, 1); op.act = r.scalar<int8_t>(o, 5, 0);
          break;
        case kOpResizeBilinear:
          op.align_corners = r.scalar<int8_t>(o, 2, 0); op.half_pixel = r.scalar<int8_t>(o, 3, 0);
          break;
        case kOpAdd: op.act = r.scalar<int8_t>(o, 0, 0); break;
        case kOpConcat: op.axis = r.scalar<int32_t>(o, 0, 0); op.act = r.scalar<int8_t>(o, 1, 0); break;
        default: break;
      }
    }
    for (int i : op.in) if (i < -1 || i >= (int)tensors.size()) return fail("op input out of range");   // -1 = optional input absent
    if (op.code != kOpDequantize && op.code != kOpReshape && !op.in.empty() && op.in[0] < 0) return fail("op without a data input");
    if (op.stride_w < 1 || op.stride_h < 1 || op.stride_w > 16 || op.stride_h > 16 || op.filter_w < 1 || op.filter_h < 1 || op.filter_w > 64 ||
        op.filter_h > 64 || op.dil_w < 1 || op.dil_h < 1) return fail("op with out-of-range stride / filter / dilation");
    for (int i : op.out) if (i < 0 || i >= (int)tensors.size()) return fail("op output out of range");
    // arity of the ops the planner dereferences: required inputs present and not the "absent" marker
    auto need = [&](size_t k) {
      if (op.in.size() < k) return false;
      for (size_t j = 0; j < k; ++j) if (op.in[j] < 0) return false;
      return true;
    };
    bool arity = !op.out.empty();
    switch (op.code) {
      case kOpConv2D: case kOpDwConv2D: case kOpPrelu: case kOpPad: case kOpAdd: arity = arity && need(2); break;
      case kOpConcat: arity = arity && !op.in.empty() && need(op.in.size()); break;
      default: arity = arity && need(1); break;
    }
    if (!arity) return fail("op with missing inputs / outputs");
    ops.push_back(std::move(op));
  }
  if (!r.ok) return fail("truncated flatbuffer");
  if (inputs.empty() || outputs.empty()) return fail("graph without inputs/outputs");
  for (int i : inputs) if (i < 0 || i >= (int)tensors.size()) return fail("graph input out of range");
  for (int i : outputs) if (i < 0 || i >= (int)tensors.size()) return fail("graph output out of range");
  return true;
}

int TfModel::producer(int tensor) const {
  for (size_t i = 0; i < ops.size(); ++i)
    for (int o : ops[i].out) if (o == tensor) return (int)i;
  return -1;
}

std::vector<int> TfModel::consumers(int tensor) const {
  std::vector<int> r;
  for (size_t i = 0; i < ops.size(); ++i)
    for (int x : ops[i].in) if (x == tensor) { r.push_back((int)i); break; }
  return r;
}

bool TfModel::const_f32(int tensor, std::vector<float>* out) const {
  if (tensor < 0 || tensor >= (int)tensors.size()) return false;
  const TfTensor* t = &tensors[tensor];
  if (!t->data) {  // follow DEQUANTIZE
    int p = producer(tensor);
    if (p < 0 || ops[p].code != kOpDequantize || ops[p].in.empty() || ops[p].in[0] < 0) return false;
    t = &tensors[ops[p].in[0]];
    if (!t->data) return false;
  }
  if (t->dtype == kTfF32) {
    size_t n = t->nbytes / 4;
    out->resize(n);
    std::memcpy(out->data(), t->data, n * 4);
    return true;
  }
  if (t->dtype == kTfF16) {
    size_t n = t->nbytes / 2;
    out->resize(n);
    for (size_t i = 0; i < n; ++i) {
      uint16_t h;
      std::memcpy(&h, t->data + 2 * i, 2);
      (*out)[i] = half_to_float(h);
    }
    return true;
  }
  return false;
}

bool TfModel::const_i32(int tensor, std::vector<int>* out) const {
  if (tensor < 0 || tensor >= (int)tensors.size()) return false;
  const TfTensor& t = tensors[tensor];
  if (!t.data || t.dtype != kTfI32) return false;
  size_t n = t.nbytes / 4;
  out->resize(n);
  std::memcpy(out->data(), t.data, n * 4);
  return true;
}

}  // namespace fdt
