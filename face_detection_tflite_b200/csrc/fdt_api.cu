// C ABI of libfdt_cuda.so (see include/fdt_api.h).  Host-side orchestration of one detector
// handle: chunked, double-streamed pipeline
//   [H2D] -> letterbox -> BlazeFace conv stack -> decode + weighted NMS [-> ROI list -> warpAffine
//   -> face_landmark conv stack -> mesh unpack] -> [D2H]
// mirroring _FaceDetectorCore.detectFacesDirect (lib/src/isolate/face_detector_core.dart:215-394).
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fdt_api.h"
#include "engine.h"
#include "fdt_math.h"
#include "kernels.h"

using namespace fdt;

static_assert(sizeof(fdt_face) == 152, "fdt_face wire layout (bindings rely on it)");
static_assert(sizeof(fdt_config) == 48, "fdt_config layout");

namespace {

#ifndef FDT_STREAMS
#define FDT_STREAMS 2
#endif
constexpr int kStreams = FDT_STREAMS;   // chunks alternate over this many streams (each with its own arena)
constexpr int kMeshInput = 192;
constexpr double kMinScore = 0.5;         // lib/src/shared/face_model_config.dart:53
constexpr double kMinSuppression = 0.3;   // lib/src/shared/face_model_config.dart:77
constexpr int kNumStages = 6;

std::string g_create_error;
std::mutex g_create_mu;

struct SsdOptions { int num_layers, input_h, input_w; int strides[4]; double interp; };

// kSsdFront / kSsdBack / kSsdFull (lib/src/shared/face_model_config.dart:80-125)
SsdOptions ssd_options(int model) {
  switch (model) {
    case FDT_MODEL_BACK_CAMERA: return {4, 256, 256, {16, 32, 32, 32}, 1.0};
    case FDT_MODEL_FULL: case FDT_MODEL_FULL_SPARSE: return {1, 192, 192, {4, 0, 0, 0}, 0.0};
    default: return {4, 128, 128, {8, 16, 16, 16}, 1.0};
  }
}

// flutter_litert generateAnchors (call site lib/src/models/face_detection_model.dart:138,:178):
// layers with equal stride are merged; aspectRatios = [1.0] plus one interpolated-scale anchor.
std::vector<double> generate_anchors(const SsdOptions& o) {
  std::vector<double> out;
  int layer = 0;
  while (layer < o.num_layers) {
    int last = layer, repeats = 0;
    while (last < o.num_layers && o.strides[last] == o.strides[layer]) {
      repeats += 1 + (o.interp > 0 ? 1 : 0);
      ++last;
    }
    int stride = o.strides[layer];
    int fh = (o.input_h + stride - 1) / stride, fw = (o.input_w + stride - 1) / stride;
    for (int y = 0; y < fh; ++y) {
      double cy = (y + 0.5) / fh;
      for (int x = 0; x < fw; ++x) {
        double cx = (x + 0.5) / fw;
        for (int r = 0; r < repeats; ++r) { out.push_back(cx); out.push_back(cy); }
      }
    }
    layer = last;
  }
  return out;
}

struct LbTables {
  int w = 0, h = 0;
  LetterboxParams lp;
  int *x0 = nullptr, *x1 = nullptr, *y0 = nullptr, *y1 = nullptr;
  short *ax0 = nullptr, *ax1 = nullptr, *by0 = nullptr, *by1 = nullptr;
  bool identity = false;
  // Sparse upload (host frames, fast mode): INTER_LINEAR only reads the rows y0/y1 name.  When those
  // rows form a periodic pattern (`run` consecutive rows every `period`, e.g. rows 10y+4,10y+5 for
  // 720 -> 72) a single strided DMA copies just them; y0c/y1c index the compacted frame.
  bool sparse = false;
  int r0 = 0, run = 0, period = 0, rows_c = 0;
  int *y0c = nullptr, *y1c = nullptr;
};

template <typename T> T* dev_upload(const std::vector<T>& v) {
  T* p = nullptr;
  if (cudaMalloc(&p, std::max<size_t>(v.size(), 1) * sizeof(T)) != cudaSuccess) return nullptr;
  cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
  return p;
}

}  // namespace

struct fdt_handle {
  fdt_config cfg;
  int chunk = 256, max_faces = FDT_MAX_FACES;
  Engine det, mesh;
  bool has_mesh = false;
  cudaStream_t streams[kStreams] = {};
  EngineCtx det_ctx[kStreams];
  EngineCtx mesh_ctx;
  uint8_t* d_frames[kStreams] = {};
  size_t d_frames_cap[kStreams] = {};
  uint8_t* d_lb[kStreams] = {};
  int* d_cand_idx[kStreams] = {};
  int* d_cand_n[kStreams] = {};
  int cand_cap = 0;
  fdt_face* d_faces = nullptr;
  int* d_counts = nullptr;
  int res_cap = 0;
  double* d_anchors = nullptr;
  std::vector<double> anchors;
  int num_anchors = 0;
  std::vector<LbTables> tables;
  // mesh stage
  int mesh_cap = 0;
  int *d_total = nullptr, *d_face_img = nullptr, *d_face_slot = nullptr, *d_overflow = nullptr;
  double *d_affine = nullptr, *d_align = nullptr, *d_mesh_score = nullptr;
  uint8_t* d_crops = nullptr;
  float* d_mesh_out = nullptr;
  float* h_mesh_out = nullptr;      // pinned
  double* h_mesh_score = nullptr;   // pinned
  int* h_counts = nullptr;          // pinned [chunk]
  fdt_face* h_faces = nullptr;      // pinned [chunk*max_faces]
  int last_mesh_faces = 0;
  // bookkeeping
  std::mutex mu;
  std::string err;
  long long launches = 0;
  long long h2d_bytes = 0;
  int last_first_chunk = 0;         // images of the last call's first chunk (debug taps)
  bool stage_timing = false;
  float stage_ms[kNumStages] = {};
  int stage_launches[kNumStages] = {};
  cudaEvent_t ev[2] = {};
  cudaEvent_t tev[3] = {};
  bool ready = false;
};

namespace {

int fail(fdt_handle* h, int code, const std::string& msg) {
  if (h) h->err = msg;
  else { std::lock_guard<std::mutex> g(g_create_mu); g_create_error = msg; }
  return code;
}

bool cuda_ok(fdt_handle* h, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  h->err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

int channels_of(int mat_type) {
  switch (mat_type) {
    case FDT_MAT_8UC1: return 1;
    case FDT_MAT_8UC3: return 3;
    case FDT_MAT_8UC4: return 4;
    default: return 0;
  }
}

const LbTables* get_tables(fdt_handle* h, int w, int hh) {
  for (const LbTables& t : h->tables)
    if (t.w == w && t.h == hh) return &t;
  LbTables t;
  t.w = w; t.h = hh;
  t.lp = letterbox_params(w, hh, h->det.in_w(), h->det.in_h());
  t.identity = t.lp.new_w == w && t.lp.new_h == hh;
  std::vector<int> x0(t.lp.new_w), x1(t.lp.new_w), y0(t.lp.new_h), y1(t.lp.new_h);
  std::vector<short> ax0(t.lp.new_w), ax1(t.lp.new_w), by0(t.lp.new_h), by1(t.lp.new_h);
  resize_linear_taps(w, t.lp.new_w, true, x0.data(), x1.data(), ax0.data(), ax1.data());
  resize_linear_taps(hh, t.lp.new_h, false, y0.data(), y1.data(), by0.data(), by1.data());
  t.x0 = dev_upload(x0); t.x1 = dev_upload(x1); t.ax0 = dev_upload(ax0); t.ax1 = dev_upload(ax1);
  t.y0 = dev_upload(y0); t.y1 = dev_upload(y1); t.by0 = dev_upload(by0); t.by1 = dev_upload(by1);
  if (!t.x0 || !t.x1 || !t.ax0 || !t.ax1 || !t.y0 || !t.y1 || !t.by0 || !t.by1) return nullptr;
  if (!t.identity) {
    std::vector<char> need(hh, 0);
    for (int i = 0; i < t.lp.new_h; ++i) { need[y0[i]] = 1; need[y1[i]] = 1; }
    int first = 0;
    while (first < hh && !need[first]) ++first;
    int run = 0;
    while (first + run < hh && need[first + run]) ++run;
    int next = first + run;
    while (next < hh && !need[next]) ++next;
    int period = next < hh ? next - first : 0;
    bool ok = period > run && hh % period == 0 && first + run <= period;
    for (int r = 0; ok && r < hh; ++r) {
      int ph = r % period;
      ok = (need[r] != 0) == (ph >= first && ph < first + run);
    }
    if (ok && run * 2 <= period) {
      std::vector<int> y0c(t.lp.new_h), y1c(t.lp.new_h);
      auto compact = [&](int r) { return (r / period) * run + (r % period - first); };
      for (int i = 0; i < t.lp.new_h; ++i) { y0c[i] = compact(y0[i]); y1c[i] = compact(y1[i]); }
      t.y0c = dev_upload(y0c); t.y1c = dev_upload(y1c);
      if (t.y0c && t.y1c) { t.sparse = true; t.r0 = first; t.run = run; t.period = period; t.rows_c = hh / period * run; }
    }
  }
  h->tables.push_back(t);
  return &h->tables.back();
}

bool ensure_results(fdt_handle* h, int batch) {
  if (batch <= h->res_cap) return true;
  if (h->d_faces) cudaFree(h->d_faces);
  if (h->d_counts) cudaFree(h->d_counts);
  h->d_faces = nullptr; h->d_counts = nullptr; h->res_cap = 0;
  int cap = std::max(batch, h->chunk);
  if (!cuda_ok(h, cudaMalloc(&h->d_faces, (size_t)cap * h->max_faces * sizeof(fdt_face)), "cudaMalloc(faces)")) return false;
  if (!cuda_ok(h, cudaMalloc(&h->d_counts, (size_t)cap * sizeof(int)), "cudaMalloc(counts)")) return false;
  h->res_cap = cap;
  return true;
}

struct StageTimer {
  fdt_handle* h; int stage; cudaStream_t s; int launches;
  StageTimer(fdt_handle* hh, int st, cudaStream_t ss) : h(hh), stage(st), s(ss), launches(0) {
    if (h->stage_timing) cudaEventRecord(h->ev[0], s);
  }
  ~StageTimer() {
    h->launches += launches;
    if (!h->stage_timing) return;
    cudaEventRecord(h->ev[1], s);
    cudaEventSynchronize(h->ev[1]);
    float ms = 0;
    cudaEventElapsedTime(&ms, h->ev[0], h->ev[1]);
    h->stage_ms[stage] += ms;
    h->stage_launches[stage] += launches;
  }
};

// Mesh stage of one chunk (standard mode): ROI list -> warp -> face_landmark -> unpack; host-side
// presence gate and result assembly.  `counts`/`faces` are the chunk's host copies.
int mesh_stage(fdt_handle* h, cudaStream_t s, const uint8_t* d_frames, long long frame_stride, int row_stride,
               int channels, int w, int hh, int n, int chunk_off, fdt_face* out_faces, int32_t* out_counts,
               float* out_mesh) {
  long long total = 0;
  for (int b = 0; b < n; ++b) total += h->h_counts[b];
  std::vector<double> scores((size_t)total, 0.0);
  std::vector<float> meshes(out_mesh ? (size_t)total * FDT_MESH_FLOATS : 0);
  h->last_mesh_faces = (int)std::min<long long>(total, h->mesh_cap);
  for (long long skip = 0; skip < total; skip += h->mesh_cap) {
    int nf = (int)std::min<long long>(h->mesh_cap, total - skip);
    {
      StageTimer t(h, 3, s);
      FaceListP fl;
      fl.counts = h->d_counts + chunk_off; fl.B = n; fl.max_faces = h->max_faces;
      fl.faces = h->d_faces + (size_t)chunk_off * h->max_faces;
      fl.img_w = w; fl.img_h = hh; fl.out_size = kMeshInput; fl.cap = h->mesh_cap; fl.skip = (int)skip;
      fl.total = h->d_total; fl.face_img = h->d_face_img; fl.face_slot = h->d_face_slot;
      fl.affine = h->d_affine; fl.align = h->d_align; fl.overflow = h->d_overflow;
      launch_build_face_list(fl, s);
      WarpP wp;
      wp.frames = d_frames; wp.frame_stride = frame_stride; wp.row_stride = row_stride; wp.channels = channels;
      wp.src_w = w; wp.src_h = hh; wp.face_img = h->d_face_img; wp.affine = h->d_affine; wp.nfaces = nf;
      wp.out_size = kMeshInput; wp.crops = h->d_crops;
      launch_warp_affine(wp, s);
      t.launches = 2;
    }
    {
      StageTimer t(h, 4, s);
      static const bool dump = [] { const char* e = std::getenv("FDT_DUMP_MESH_STEPS"); return e && e[0] == '1'; }();
      if (dump) {
        // diagnostic: per-kernel device time of the mesh plan on stderr
        const size_t S = h->mesh.plan().steps.size();
        std::vector<cudaEvent_t> ev(S + 1);
        for (auto& e : ev) cudaEventCreate(&e);
        t.launches = h->mesh.run(h->mesh_ctx, h->d_crops, nf, s, ev.data());
        cudaStreamSynchronize(s);
        for (size_t i = 0; i < S; ++i) {
          float ms = 0;
          cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
          std::fprintf(stderr, "mesh step %2zu kind %2d %-28s %8.3f ms (%d faces)\n", i, (int)h->mesh.plan().steps[i].kind,
                       h->mesh.plan().steps[i].name.c_str(), ms, nf);
        }
        for (auto& e : ev) cudaEventDestroy(e);
      } else {
        t.launches = h->mesh.run(h->mesh_ctx, h->d_crops, nf, s);
      }
    }
    {
      StageTimer t(h, 5, s);
      // face_landmark.dart:154-166: landmarks = largest output divisible by 3, score = first 1-element output
      const Plan& mp = h->mesh.plan();
      int li = -1, si = -1;
      for (size_t k = 0; k < mp.out_elems.size(); ++k) {
        if (mp.out_elems[k] % 3 == 0 && (li < 0 || mp.out_elems[k] > mp.out_elems[li])) li = (int)k;
        if (mp.out_elems[k] == 1 && si < 0) si = (int)k;
      }
      MeshPostP pp;
      pp.raw = h->mesh_ctx.outputs[li]; pp.raw_istride = mp.out_elems[li];
      pp.flag = h->mesh_ctx.outputs[si]; pp.flag_istride = mp.out_elems[si];
      pp.align = h->d_align; pp.nfaces = nf; pp.in_size = kMeshInput;
      pp.mesh_out = h->d_mesh_out; pp.score_out = h->d_mesh_score;
      launch_mesh_post(pp, s);
      t.launches = 1;
    }
    cudaMemcpyAsync(h->h_mesh_score, h->d_mesh_score, (size_t)nf * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (out_mesh)
      cudaMemcpyAsync(h->h_mesh_out, h->d_mesh_out, (size_t)nf * FDT_MESH_FLOATS * sizeof(float), cudaMemcpyDeviceToHost, s);
    if (!cuda_ok(h, cudaStreamSynchronize(s), "mesh stage")) return FDT_ERR_CUDA;
    std::copy(h->h_mesh_score, h->h_mesh_score + nf, scores.begin() + skip);
    if (out_mesh) std::memcpy(meshes.data() + (size_t)skip * FDT_MESH_FLOATS, h->h_mesh_out, (size_t)nf * FDT_MESH_FLOATS * sizeof(float));
  }
  // presence gate (_passesPresence, face_detector_core.dart:101-103, :353) + assembly
  const double gate = h->cfg.min_face_presence;
  long long f = 0;
  for (int b = 0; b < n; ++b) {
    int kept = 0;
    for (int j = 0; j < h->h_counts[b]; ++j, ++f) {
      double sc = scores[(size_t)f];
      if (!(gate <= 0.0 || sc >= gate)) continue;
      fdt_face fc = h->h_faces[(size_t)b * h->max_faces + j];
      fc.mesh_score = sc;
      fc.has_mesh = 1;
      size_t slot = (size_t)(chunk_off + b) * h->max_faces + kept;
      out_faces[slot] = fc;
      if (out_mesh) std::memcpy(out_mesh + slot * FDT_MESH_FLOATS, meshes.data() + (size_t)f * FDT_MESH_FLOATS, FDT_MESH_FLOATS * sizeof(float));
      ++kept;
    }
    out_counts[chunk_off + b] = kept;
  }
  return FDT_OK;
}

int detect_impl(fdt_handle* h, const uint8_t* frames, int batch, int w, int hh, int row_stride, int mat_type, int mode,
                int mem_kind, fdt_face* out_faces, int32_t* out_counts, float* out_mesh, bool keep_on_device) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  int channels = channels_of(mat_type);
  if (!frames || batch < 0 || w <= 0 || hh <= 0 || channels == 0) return fail(h, FDT_ERR_BAD_ARG, "bad frame arguments");
  if (row_stride < w * channels) return fail(h, FDT_ERR_SIZE_MISMATCH, "row_stride smaller than width * channels");
  if (mode == FDT_MODE_FULL) return fail(h, FDT_ERR_UNSUPPORTED, "FaceDetectionMode.full (iris/blendshapes) is outside this path");
  if (mode != FDT_MODE_FAST && mode != FDT_MODE_STANDARD) return fail(h, FDT_ERR_BAD_ARG, "unknown mode");
  if (mode == FDT_MODE_STANDARD && !h->has_mesh) return fail(h, FDT_ERR_NOT_READY, "standard mode needs the face_landmark model");
  if (mode == FDT_MODE_STANDARD && keep_on_device) return fail(h, FDT_ERR_UNSUPPORTED, "device-resident results are fast-mode only");
  if (!keep_on_device && (!out_faces || !out_counts)) return fail(h, FDT_ERR_BAD_ARG, "null output buffers");
  cudaSetDevice(h->cfg.device);
  h->launches = 0;
  h->h2d_bytes = 0;
  for (int i = 0; i < kNumStages; ++i) { h->stage_ms[i] = 0; h->stage_launches[i] = 0; }
  if (batch == 0) return FDT_OK;
  if (!ensure_results(h, batch)) return FDT_ERR_CUDA;
  const LbTables* tb = get_tables(h, w, hh);
  if (!tb) return fail(h, FDT_ERR_CUDA, "letterbox table upload failed");
  const int S_w = h->det.in_w(), S_h = h->det.in_h();
  const long long frame_stride = (long long)hh * row_stride;
  h->last_first_chunk = std::min(batch, h->chunk);

  for (int off = 0, c = 0; off < batch; off += h->chunk, ++c) {
    const int n = std::min(h->chunk, batch - off);
    const int si = (mode == FDT_MODE_STANDARD) ? 0 : c % kStreams;
    cudaStream_t s = h->streams[si];
    const uint8_t* d_fr;
    const bool sparse = mem_kind == FDT_MEM_HOST && mode == FDT_MODE_FAST && tb->sparse;
    long long dev_frame_stride = frame_stride;
    if (mem_kind == FDT_MEM_HOST) {
      size_t need = (size_t)h->chunk * frame_stride;
      if (h->d_frames_cap[si] < need) {
        cudaStreamSynchronize(s);
        if (h->d_frames[si]) cudaFree(h->d_frames[si]);
        h->d_frames[si] = nullptr; h->d_frames_cap[si] = 0;
        if (!cuda_ok(h, cudaMalloc(&h->d_frames[si], need), "cudaMalloc(frame staging)")) return FDT_ERR_CUDA;
        h->d_frames_cap[si] = need;
      }
      const uint8_t* src = frames + (size_t)off * frame_stride;
      if (sparse) {
        // one strided DMA: `run` rows out of every `period`, for all n frames (frames are contiguous)
        size_t width_b = (size_t)tb->run * row_stride;
        cudaMemcpy2DAsync(h->d_frames[si], width_b, src + (size_t)tb->r0 * row_stride, (size_t)tb->period * row_stride,
                          width_b, (size_t)n * (hh / tb->period), cudaMemcpyHostToDevice, s);
        dev_frame_stride = (long long)tb->rows_c * row_stride;
        h->h2d_bytes += (long long)width_b * n * (hh / tb->period);
      } else {
        cudaMemcpyAsync(h->d_frames[si], src, (size_t)n * frame_stride, cudaMemcpyHostToDevice, s);
        h->h2d_bytes += (long long)n * frame_stride;
      }
      d_fr = h->d_frames[si];
    } else {
      d_fr = frames + (size_t)off * frame_stride;
    }
    {
      StageTimer t(h, 0, s);
      LetterboxP lb;
      lb.frames = d_fr; lb.frame_stride = dev_frame_stride; lb.row_stride = row_stride; lb.channels = channels;
      lb.src_w = w; lb.src_h = hh; lb.out = h->d_lb[si]; lb.dst_w = S_w; lb.dst_h = S_h;
      lb.new_w = tb->lp.new_w; lb.new_h = tb->lp.new_h; lb.pad_top = tb->lp.pad_top; lb.pad_left = tb->lp.pad_left;
      lb.x0 = tb->x0; lb.x1 = tb->x1; lb.ax0 = tb->ax0; lb.ax1 = tb->ax1;
      lb.y0 = sparse ? tb->y0c : tb->y0; lb.y1 = sparse ? tb->y1c : tb->y1; lb.by0 = tb->by0; lb.by1 = tb->by1;
      lb.identity = tb->identity ? 1 : 0;
      launch_letterbox(lb, n, s);
      t.launches = 1;
    }
    {
      StageTimer t(h, 1, s);
      t.launches = h->det.run(h->det_ctx[si], h->d_lb[si], n, s);
    }
    {
      StageTimer t(h, 2, s);
      const Plan& dp = h->det.plan();
      DecodeP d;
      d.boxes = h->det_ctx[si].outputs[0]; d.boxes_istride = dp.out_elems[0];   // _boundingBoxIndex = 0
      d.scores = h->det_ctx[si].outputs[1]; d.scores_istride = dp.out_elems[1]; // _scoreIndex = 1
      d.anchors = h->d_anchors; d.N = h->num_anchors; d.input_h = S_h;
      d.raw_thresh = std::log(kMinScore / (1.0 - kMinScore));
      d.score_thresh = kMinScore; d.iou_thresh = kMinSuppression;
      d.pad_t = (double)tb->lp.pad_top / S_h; d.pad_b = (double)tb->lp.pad_bottom / S_h;
      d.pad_l = (double)tb->lp.pad_left / S_w; d.pad_r = (double)tb->lp.pad_right / S_w;
      d.min_score = h->cfg.min_score; d.min_face_size = h->cfg.min_face_size;
      d.img_w = w; d.img_h = hh; d.max_faces = h->max_faces;
      d.faces = h->d_faces + (size_t)off * h->max_faces; d.counts = h->d_counts + off;
      d.cand_idx = c == 0 ? h->d_cand_idx[si] : nullptr; d.cand_cap = h->cand_cap;
      d.cand_n = c == 0 ? h->d_cand_n[si] : nullptr;
      launch_decode_nms(d, n, s);
      t.launches = 1;
    }
    if (mode == FDT_MODE_STANDARD) {
      cudaMemcpyAsync(h->h_counts, h->d_counts + off, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(h->h_faces, h->d_faces + (size_t)off * h->max_faces, (size_t)n * h->max_faces * sizeof(fdt_face),
                      cudaMemcpyDeviceToHost, s);
      if (!cuda_ok(h, cudaStreamSynchronize(s), "detector stage")) return FDT_ERR_CUDA;
      int rc = mesh_stage(h, s, d_fr, frame_stride, row_stride, channels, w, hh, n, off, out_faces, out_counts, out_mesh);
      if (rc != FDT_OK) return rc;
    } else if (!keep_on_device) {
      cudaMemcpyAsync(out_counts + off, h->d_counts + off, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(out_faces + (size_t)off * h->max_faces, h->d_faces + (size_t)off * h->max_faces,
                      (size_t)n * h->max_faces * sizeof(fdt_face), cudaMemcpyDeviceToHost, s);
    }
  }
  if (!keep_on_device) {
    for (int i = 0; i < kStreams; ++i)
      if (!cuda_ok(h, cudaStreamSynchronize(h->streams[i]), "pipeline")) return FDT_ERR_CUDA;
  }
  if (!cuda_ok(h, cudaGetLastError(), "kernel launch")) return FDT_ERR_CUDA;
  if (h->det.failed() || h->mesh.failed()) { h->err = "TMA tensor map encoding failed (cuTensorMapEncodeTiled)"; return FDT_ERR_CUDA; }
  return FDT_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

void fdt_default_config(fdt_config* cfg) {
  if (!cfg) return;
  std::memset(cfg, 0, sizeof(*cfg));
  cfg->struct_size = (int32_t)sizeof(fdt_config);
  cfg->model = FDT_MODEL_BACK_CAMERA;   // FaceDetector.create default (lib/src/face_detector.dart:85)
  cfg->device = 0;
  cfg->max_batch = 0;
  cfg->max_faces = 0;
  cfg->fuse_level = -1;
  cfg->min_score = 0.0;
  cfg->min_face_size = 0.0;
  cfg->min_face_presence = 0.5;         // kDefaultMinFacePresenceConfidence
}

int32_t fdt_create(const fdt_config* cfg_in, const uint8_t* det_tflite, size_t det_len, const uint8_t* mesh_tflite,
                   size_t mesh_len, fdt_handle** out) {
  if (!out) return fail(nullptr, FDT_ERR_BAD_ARG, "null out pointer");
  *out = nullptr;
  fdt_config cfg;
  fdt_default_config(&cfg);
  if (cfg_in) {
    if (cfg_in->struct_size != (int32_t)sizeof(fdt_config)) return fail(nullptr, FDT_ERR_BAD_ARG, "fdt_config.struct_size mismatch");
    cfg = *cfg_in;
  }
  // validateFaceGates (lib/src/shared/face_gates.dart:31-59)
  auto bad = [](double v) { return std::isnan(v) || v < 0.0 || v > 1.0; };
  if (bad(cfg.min_score)) return fail(nullptr, FDT_ERR_BAD_ARG, "minScore must be in the inclusive range [0.0, 1.0]");
  if (bad(cfg.min_face_size)) return fail(nullptr, FDT_ERR_BAD_ARG, "minFaceSize must be in the inclusive range [0.0, 1.0]");
  if (bad(cfg.min_face_presence)) return fail(nullptr, FDT_ERR_BAD_ARG, "minFacePresenceConfidence must be in the inclusive range [0.0, 1.0]");
  if (cfg.model < FDT_MODEL_FRONT_CAMERA || cfg.model > FDT_MODEL_FULL_SPARSE) return fail(nullptr, FDT_ERR_BAD_ARG, "unknown model");
  if (cfg.model == FDT_MODEL_FULL_SPARSE) return fail(nullptr, FDT_ERR_UNSUPPORTED, "fullSparse (DENSIFY) is outside this path");
  if (!det_tflite || det_len == 0) return fail(nullptr, FDT_ERR_BAD_ARG, "missing detector model bytes");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
    return fail(nullptr, FDT_ERR_CUDA, "no CUDA device: libfdt_cuda has no CPU fallback");
  if (cfg.device < 0 || cfg.device >= ndev) return fail(nullptr, FDT_ERR_BAD_ARG, "bad device ordinal");
  if (cudaSetDevice(cfg.device) != cudaSuccess) return fail(nullptr, FDT_ERR_CUDA, "cudaSetDevice failed");

  fdt_handle* h = new fdt_handle();
  h->cfg = cfg;
  h->chunk = cfg.max_batch > 0 ? cfg.max_batch : 256;
  h->max_faces = cfg.max_faces > 0 ? std::min(cfg.max_faces, (int)FDT_MAX_FACES) : FDT_MAX_FACES;
  int fuse = cfg.fuse_level < 0 ? 1 : cfg.fuse_level;
  // FDT_NO_TC=1 keeps the pointwise GEMMs on the fp32 CUDA-core kernels (A/B measurements)
  const char* no_tc = std::getenv("FDT_NO_TC");
  const bool use_tc = !(no_tc && no_tc[0] == '1');
  std::string err;
  auto bail = [&](int code, const std::string& m) {
    fail(nullptr, code, m);
    fdt_destroy(h);
    return code;
  };
  if (!h->det.init(det_tflite, det_len, fuse, &err, use_tc)) return bail(FDT_ERR_MODEL, "detector model: " + err);
  SsdOptions so = ssd_options(cfg.model);
  if (h->det.in_h() != so.input_h || h->det.in_w() != so.input_w)
    return bail(FDT_ERR_MODEL, "detector input size does not match the selected FaceDetectionModel");
  h->anchors = generate_anchors(so);
  h->num_anchors = (int)h->anchors.size() / 2;
  const Plan& dp = h->det.plan();
  if (dp.out_elems.size() < 2 || dp.out_elems[0] != (long long)h->num_anchors * 16 || dp.out_elems[1] != h->num_anchors)
    return bail(FDT_ERR_MODEL, "detector outputs do not match [N,16] boxes / [N] scores for the SSD anchors");
  h->d_anchors = dev_upload(h->anchors);
  if (!h->d_anchors) return bail(FDT_ERR_CUDA, "anchor upload failed");
  h->cand_cap = h->num_anchors;
  for (int i = 0; i < kStreams; ++i) {
    if (cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking) != cudaSuccess) return bail(FDT_ERR_CUDA, "stream creation failed");
    if (!h->det.make_ctx(h->chunk, &h->det_ctx[i], &err)) return bail(FDT_ERR_CUDA, err);
    size_t lb = (size_t)h->chunk * h->det.in_h() * h->det.in_w() * 4;   // BGRX
    if (cudaMalloc(&h->d_lb[i], lb) != cudaSuccess) return bail(FDT_ERR_CUDA, "cudaMalloc(letterboxed) failed");
    if (cudaMalloc(&h->d_cand_idx[i], (size_t)h->chunk * h->cand_cap * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&h->d_cand_n[i], (size_t)h->chunk * sizeof(int)) != cudaSuccess)
      return bail(FDT_ERR_CUDA, "cudaMalloc(candidates) failed");
    cudaMemset(h->d_cand_n[i], 0, (size_t)h->chunk * sizeof(int));
  }
  cudaEventCreate(&h->ev[0]);
  cudaEventCreate(&h->ev[1]);
  for (int i = 0; i < 3; ++i) cudaEventCreate(&h->tev[i]);
  if (mesh_tflite && mesh_len) {
    if (!h->mesh.init(mesh_tflite, mesh_len, fuse, &err, use_tc)) return bail(FDT_ERR_MODEL, "mesh model: " + err);
    if (h->mesh.in_h() != kMeshInput || h->mesh.in_w() != kMeshInput) return bail(FDT_ERR_MODEL, "mesh model input must be 192x192");
    const Plan& mp = h->mesh.plan();
    bool has3 = false, has1 = false;
    for (long long e : mp.out_elems) { has3 |= (e == FDT_MESH_FLOATS); has1 |= (e == 1); }
    if (!has3 || !has1) return bail(FDT_ERR_MODEL, "mesh model must output 1404 landmarks and a face flag");
    h->mesh_cap = std::max(64, h->chunk * 4);
    if (!h->mesh.make_ctx(h->mesh_cap, &h->mesh_ctx, &err)) return bail(FDT_ERR_CUDA, err);
    size_t mc = (size_t)h->mesh_cap;
    bool ok = cudaMalloc(&h->d_total, sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_overflow, sizeof(int)) == cudaSuccess &&
              cudaMalloc(&h->d_face_img, mc * sizeof(int)) == cudaSuccess && cudaMalloc(&h->d_face_slot, mc * sizeof(int)) == cudaSuccess &&
              cudaMalloc(&h->d_affine, mc * 6 * sizeof(double)) == cudaSuccess && cudaMalloc(&h->d_align, mc * 4 * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&h->d_mesh_score, mc * sizeof(double)) == cudaSuccess &&
              cudaMalloc(&h->d_crops, mc * kMeshInput * kMeshInput * 4) == cudaSuccess &&
              cudaMalloc(&h->d_mesh_out, mc * FDT_MESH_FLOATS * sizeof(float)) == cudaSuccess &&
              cudaMallocHost(&h->h_mesh_out, mc * FDT_MESH_FLOATS * sizeof(float)) == cudaSuccess &&
              cudaMallocHost(&h->h_mesh_score, mc * sizeof(double)) == cudaSuccess;
    if (!ok) return bail(FDT_ERR_CUDA, "mesh stage allocation failed");
    cudaMemset(h->d_overflow, 0, sizeof(int));
    h->has_mesh = true;
  }
  if (cudaMallocHost(&h->h_counts, (size_t)h->chunk * sizeof(int)) != cudaSuccess ||
      cudaMallocHost(&h->h_faces, (size_t)h->chunk * h->max_faces * sizeof(fdt_face)) != cudaSuccess)
    return bail(FDT_ERR_CUDA, "pinned allocation failed");
  if (cudaDeviceSynchronize() != cudaSuccess) return bail(FDT_ERR_CUDA, "device initialisation failed");
  h->ready = true;
  *out = h;
  return FDT_OK;
}

int32_t fdt_destroy(fdt_handle* h) {
  if (!h) return FDT_OK;
  {
    std::lock_guard<std::mutex> g(h->mu);
    h->ready = false;
    cudaSetDevice(h->cfg.device);
    cudaDeviceSynchronize();
    for (int i = 0; i < kStreams; ++i) {
      h->det.free_ctx(&h->det_ctx[i]);
      if (h->d_frames[i]) cudaFree(h->d_frames[i]);
      if (h->d_lb[i]) cudaFree(h->d_lb[i]);
      if (h->d_cand_idx[i]) cudaFree(h->d_cand_idx[i]);
      if (h->d_cand_n[i]) cudaFree(h->d_cand_n[i]);
      if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
    }
    h->mesh.free_ctx(&h->mesh_ctx);
    void* dev[] = {h->d_faces, h->d_counts, h->d_anchors, h->d_total, h->d_face_img, h->d_face_slot, h->d_overflow,
                   h->d_affine, h->d_align, h->d_mesh_score, h->d_crops, h->d_mesh_out};
    for (void* p : dev) if (p) cudaFree(p);
    void* pin[] = {h->h_mesh_out, h->h_mesh_score, h->h_counts, h->h_faces};
    for (void* p : pin) if (p) cudaFreeHost(p);
    for (LbTables& t : h->tables) {
      void* tp[] = {t.x0, t.x1, t.y0, t.y1, t.ax0, t.ax1, t.by0, t.by1, t.y0c, t.y1c};
      for (void* p : tp) if (p) cudaFree(p);
    }
    if (h->ev[0]) cudaEventDestroy(h->ev[0]);
    if (h->ev[1]) cudaEventDestroy(h->ev[1]);
    for (int i = 0; i < 3; ++i) if (h->tev[i]) cudaEventDestroy(h->tev[i]);
  }
  delete h;
  return FDT_OK;
}

int32_t fdt_detect_batch(fdt_handle* h, const uint8_t* frames, int32_t batch, int32_t width, int32_t height,
                         int32_t row_stride, int32_t mat_type, int32_t mode, int32_t mem_kind, fdt_face* out_faces,
                         int32_t* out_counts, float* out_mesh) {
  if (!h) return fail(nullptr, FDT_ERR_NOT_READY, "null handle");
  std::lock_guard<std::mutex> g(h->mu);
  return detect_impl(h, frames, batch, width, height, row_stride, mat_type, mode, mem_kind, out_faces, out_counts, out_mesh, false);
}

int32_t fdt_detect_one(fdt_handle* h, const uint8_t* bytes, size_t nbytes, int32_t width, int32_t height, int32_t mat_type,
                       int32_t mode, fdt_face* out_faces, int32_t* out_count, float* out_mesh) {
  if (!h) return fail(nullptr, FDT_ERR_NOT_READY, "null handle");
  std::lock_guard<std::mutex> g(h->mu);
  int ch = channels_of(mat_type);
  if (ch == 0 || width <= 0 || height <= 0) return fail(h, FDT_ERR_BAD_ARG, "bad frame arguments");
  // matFromPackedBytes length check (lib/src/util/helpers.dart:440-447)
  if (nbytes != (size_t)width * height * ch) return fail(h, FDT_ERR_SIZE_MISMATCH, "bytes length does not equal width * height * channels");
  return detect_impl(h, bytes, 1, width, height, width * ch, mat_type, mode, FDT_MEM_HOST, out_faces, out_count, out_mesh, false);
}

int32_t fdt_detect_batch_device(fdt_handle* h, const uint8_t* d_frames, int32_t batch, int32_t width, int32_t height,
                                int32_t row_stride, int32_t mat_type, int32_t mode, const fdt_face** d_faces,
                                const int32_t** d_counts) {
  if (!h) return fail(nullptr, FDT_ERR_NOT_READY, "null handle");
  std::lock_guard<std::mutex> g(h->mu);
  int rc = detect_impl(h, d_frames, batch, width, height, row_stride, mat_type, mode, FDT_MEM_DEVICE, nullptr, nullptr, nullptr, true);
  if (rc == FDT_OK) {
    if (d_faces) *d_faces = h->d_faces;
    if (d_counts) *d_counts = h->d_counts;
  }
  return rc;
}

int32_t fdt_synchronize(fdt_handle* h) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->cfg.device);
  for (int i = 0; i < kStreams; ++i)
    if (!cuda_ok(h, cudaStreamSynchronize(h->streams[i]), "synchronize")) return FDT_ERR_CUDA;
  return FDT_OK;
}

int32_t fdt_get_info(fdt_handle* h, int32_t* input_w, int32_t* input_h, int32_t* num_anchors, int32_t* max_faces, int32_t* max_batch) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  if (input_w) *input_w = h->det.in_w();
  if (input_h) *input_h = h->det.in_h();
  if (num_anchors) *num_anchors = h->num_anchors;
  if (max_faces) *max_faces = h->max_faces;
  if (max_batch) *max_batch = h->chunk;
  return FDT_OK;
}

int32_t fdt_get_anchors(fdt_handle* h, double* out_xy) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  if (!out_xy) return fail(h, FDT_ERR_BAD_ARG, "null output");
  std::memcpy(out_xy, h->anchors.data(), h->anchors.size() * sizeof(double));
  return FDT_OK;
}

int32_t fdt_letterbox_params(int32_t src_w, int32_t src_h, int32_t dst_w, int32_t dst_h, int32_t* out6) {
  if (!out6 || src_w <= 0 || src_h <= 0 || dst_w <= 0 || dst_h <= 0) return FDT_ERR_BAD_ARG;
  LetterboxParams p = letterbox_params(src_w, src_h, dst_w, dst_h);
  out6[0] = p.new_w; out6[1] = p.new_h; out6[2] = p.pad_top; out6[3] = p.pad_bottom; out6[4] = p.pad_left; out6[5] = p.pad_right;
  return FDT_OK;
}

int32_t fdt_alloc_pinned(size_t nbytes, void** out) {
  if (!out) return FDT_ERR_BAD_ARG;
  return cudaMallocHost(out, nbytes) == cudaSuccess ? FDT_OK : FDT_ERR_CUDA;
}
int32_t fdt_free_pinned(void* p) { return cudaFreeHost(p) == cudaSuccess ? FDT_OK : FDT_ERR_CUDA; }

int32_t fdt_alloc_device(fdt_handle* h, size_t nbytes, void** out) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  if (!out) return fail(h, FDT_ERR_BAD_ARG, "null out pointer");
  cudaSetDevice(h->cfg.device);
  return cuda_ok(h, cudaMalloc(out, nbytes), "cudaMalloc") ? FDT_OK : FDT_ERR_CUDA;
}
int32_t fdt_free_device(fdt_handle* h, void* p) {
  if (!h) return FDT_ERR_NOT_READY;
  cudaSetDevice(h->cfg.device);
  return cudaFree(p) == cudaSuccess ? FDT_OK : FDT_ERR_CUDA;
}
int32_t fdt_copy_to_device(fdt_handle* h, void* dst, const void* src, size_t nbytes) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  cudaSetDevice(h->cfg.device);
  return cuda_ok(h, cudaMemcpy(dst, src, nbytes, cudaMemcpyHostToDevice), "cudaMemcpy") ? FDT_OK : FDT_ERR_CUDA;
}

int32_t fdt_copy_to_host(fdt_handle* h, void* dst, const void* src, size_t nbytes) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  cudaSetDevice(h->cfg.device);
  for (int i = 0; i < kStreams; ++i) cudaStreamSynchronize(h->streams[i]);
  return cuda_ok(h, cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToHost), "cudaMemcpy") ? FDT_OK : FDT_ERR_CUDA;
}

// ---- parity taps ------------------------------------------------------------------------------
int32_t fdt_debug_get_letterboxed(fdt_handle* h, int32_t n, uint8_t* out) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  if (n < 0 || n > h->last_first_chunk || !out) return fail(h, FDT_ERR_BAD_ARG, "n exceeds the last call's first chunk");
  cudaSetDevice(h->cfg.device);
  size_t px = (size_t)n * h->det.in_h() * h->det.in_w();
  std::vector<uint8_t> tmp(px * 4);
  if (!cuda_ok(h, cudaMemcpy(tmp.data(), h->d_lb[0], px * 4, cudaMemcpyDeviceToHost), "tap")) return FDT_ERR_CUDA;
  for (size_t i = 0; i < px; ++i) { out[3 * i] = tmp[4 * i]; out[3 * i + 1] = tmp[4 * i + 1]; out[3 * i + 2] = tmp[4 * i + 2]; }
  return FDT_OK;
}

int32_t fdt_debug_get_input_tensor(fdt_handle* h, int32_t n, float* out) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  if (n < 0 || n > h->last_first_chunk || !out) return fail(h, FDT_ERR_BAD_ARG, "n exceeds the last call's first chunk");
  cudaSetDevice(h->cfg.device);
  float* tmp = nullptr;
  size_t elems = (size_t)n * h->det.in_h() * h->det.in_w() * 3;
  if (!cuda_ok(h, cudaMalloc(&tmp, std::max<size_t>(elems, 1) * sizeof(float)), "tap alloc")) return FDT_ERR_CUDA;
  TV v;
  v.p = tmp; v.H = h->det.in_h(); v.W = h->det.in_w(); v.C = 3; v.Cs = 3; v.istride = (long long)v.H * v.W * 3;
  launch_normalize(h->d_lb[0], v, n, h->streams[0]);
  cudaStreamSynchronize(h->streams[0]);
  bool ok = cuda_ok(h, cudaMemcpy(out, tmp, elems * sizeof(float), cudaMemcpyDeviceToHost), "tap");
  cudaFree(tmp);
  return ok ? FDT_OK : FDT_ERR_CUDA;
}

int32_t fdt_debug_get_raw_heads(fdt_handle* h, int32_t n, float* out_boxes, float* out_scores) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  if (n < 0 || n > h->last_first_chunk) return fail(h, FDT_ERR_BAD_ARG, "n exceeds the last call's first chunk");
  cudaSetDevice(h->cfg.device);
  const Plan& dp = h->det.plan();
  bool ok = true;
  if (out_boxes) ok &= cuda_ok(h, cudaMemcpy(out_boxes, h->det_ctx[0].outputs[0], (size_t)n * dp.out_elems[0] * sizeof(float), cudaMemcpyDeviceToHost), "tap");
  if (out_scores) ok &= cuda_ok(h, cudaMemcpy(out_scores, h->det_ctx[0].outputs[1], (size_t)n * dp.out_elems[1] * sizeof(float), cudaMemcpyDeviceToHost), "tap");
  return ok ? FDT_OK : FDT_ERR_CUDA;
}

int32_t fdt_debug_get_candidates(fdt_handle* h, int32_t image, int32_t* out_indices, int32_t capacity, int32_t* out_n) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  if (image < 0 || image >= h->last_first_chunk || !out_n) return fail(h, FDT_ERR_BAD_ARG, "image exceeds the last call's first chunk");
  cudaSetDevice(h->cfg.device);
  int n = 0;
  if (!cuda_ok(h, cudaMemcpy(&n, h->d_cand_n[0] + image, sizeof(int), cudaMemcpyDeviceToHost), "tap")) return FDT_ERR_CUDA;
  *out_n = n;
  int m = std::min(std::min(n, capacity), h->cand_cap);
  if (out_indices && m > 0 &&
      !cuda_ok(h, cudaMemcpy(out_indices, h->d_cand_idx[0] + (size_t)image * h->cand_cap, (size_t)m * sizeof(int), cudaMemcpyDeviceToHost), "tap"))
    return FDT_ERR_CUDA;
  return FDT_OK;
}

int32_t fdt_debug_get_tensor(fdt_handle* h, int32_t which, int32_t tflite_tensor, int32_t n, float* out,
                             size_t out_capacity_floats, int32_t* out_dims4) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  const Engine& e = which == 0 ? h->det : h->mesh;
  const EngineCtx& ctx = which == 0 ? h->det_ctx[0] : h->mesh_ctx;
  if (which == 1 && !h->has_mesh) return fail(h, FDT_ERR_NOT_READY, "no mesh model");
  const Plan& p = e.plan();
  auto it = p.tf2pt.find(tflite_tensor);
  if (it == p.tf2pt.end() || !p.tensors[it->second].materialized) return fail(h, FDT_ERR_BAD_ARG, "tensor is not materialised in this plan");
  if (p.fuse_level == 1 && p.tensors[it->second].root < 0) return fail(h, FDT_ERR_UNSUPPORTED, "activation buffers are reused at fuse_level 1; use 0 or 2");
  int cap = which == 0 ? h->last_first_chunk : h->last_mesh_faces;
  if (n < 0 || n > cap) return fail(h, FDT_ERR_BAD_ARG, "n exceeds the images of the last call");
  cudaSetDevice(h->cfg.device);
  TV v = e.view(ctx, it->second);
  if (out_dims4) { out_dims4[0] = n; out_dims4[1] = v.H; out_dims4[2] = v.W; out_dims4[3] = v.C; }
  size_t need = (size_t)n * v.H * v.W * v.C;
  if (!out) return FDT_OK;
  if (out_capacity_floats < need) return fail(h, FDT_ERR_SIZE_MISMATCH, "output buffer too small");
  // strided device layout -> dense NHWC on the host
  cudaError_t ce = cudaSuccess;
  if (v.Cs == v.C) {
    ce = cudaMemcpy2D(out, (size_t)v.H * v.W * v.C * sizeof(float), v.p, (size_t)v.istride * sizeof(float),
                      (size_t)v.H * v.W * v.C * sizeof(float), n, cudaMemcpyDeviceToHost);
  } else {
    for (int b = 0; b < n && ce == cudaSuccess; ++b)
      ce = cudaMemcpy2D(out + (size_t)b * v.H * v.W * v.C, (size_t)v.C * sizeof(float), v.p + (size_t)b * v.istride,
                        (size_t)v.Cs * sizeof(float), (size_t)v.C * sizeof(float), (size_t)v.H * v.W, cudaMemcpyDeviceToHost);
  }
  return cuda_ok(h, ce, "tap") ? FDT_OK : FDT_ERR_CUDA;
}

int32_t fdt_debug_get_mesh_stage(fdt_handle* h, int32_t n, uint8_t* out_crops, float* out_raw1404, float* out_flag, int32_t* out_n) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  if (!h->has_mesh) return fail(h, FDT_ERR_NOT_READY, "no mesh model");
  if (out_n) *out_n = h->last_mesh_faces;
  n = std::min(n, h->last_mesh_faces);
  if (n <= 0) return FDT_OK;
  cudaSetDevice(h->cfg.device);
  const Plan& mp = h->mesh.plan();
  bool ok = true;
  if (out_crops) {
    size_t px = (size_t)n * kMeshInput * kMeshInput;
    std::vector<uint8_t> tmp(px * 4);
    ok &= cuda_ok(h, cudaMemcpy(tmp.data(), h->d_crops, px * 4, cudaMemcpyDeviceToHost), "tap");
    for (size_t i = 0; i < px; ++i) { out_crops[3 * i] = tmp[4 * i]; out_crops[3 * i + 1] = tmp[4 * i + 1]; out_crops[3 * i + 2] = tmp[4 * i + 2]; }
  }
  for (size_t k = 0; k < mp.out_elems.size(); ++k) {
    if (mp.out_elems[k] == FDT_MESH_FLOATS && out_raw1404)
      ok &= cuda_ok(h, cudaMemcpy(out_raw1404, h->mesh_ctx.outputs[k], (size_t)n * FDT_MESH_FLOATS * sizeof(float), cudaMemcpyDeviceToHost), "tap");
    if (mp.out_elems[k] == 1 && out_flag)
      ok &= cuda_ok(h, cudaMemcpy(out_flag, h->mesh_ctx.outputs[k], (size_t)n * sizeof(float), cudaMemcpyDeviceToHost), "tap");
  }
  return ok ? FDT_OK : FDT_ERR_CUDA;
}

int64_t fdt_last_launch_count(fdt_handle* h) { return h ? h->launches : 0; }
int64_t fdt_last_h2d_bytes(fdt_handle* h) { return h ? h->h2d_bytes : 0; }

int32_t fdt_set_stage_timing(fdt_handle* h, int32_t enable) {
  if (!h) return FDT_ERR_NOT_READY;
  h->stage_timing = enable != 0;
  return FDT_OK;
}
int32_t fdt_get_stage_ms(fdt_handle* h, int32_t stage, float* ms, int32_t* launches) {
  if (!h || stage < 0 || stage >= kNumStages) return FDT_ERR_BAD_ARG;
  if (ms) *ms = h->stage_ms[stage];
  if (launches) *launches = h->stage_launches[stage];
  return FDT_OK;
}

int32_t fdt_timer_begin(fdt_handle* h) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->cfg.device);
  for (int i = 0; i < kStreams; ++i) cudaStreamSynchronize(h->streams[i]);
  cudaEventRecord(h->tev[0], h->streams[0]);
  for (int i = 1; i < kStreams; ++i) cudaStreamWaitEvent(h->streams[i], h->tev[0], 0);
  return FDT_OK;
}

int32_t fdt_timer_end(fdt_handle* h, float* ms) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  cudaSetDevice(h->cfg.device);
  for (int i = 1; i < kStreams; ++i) {
    cudaEventRecord(h->tev[2], h->streams[i]);
    cudaStreamWaitEvent(h->streams[0], h->tev[2], 0);
  }
  cudaEventRecord(h->tev[1], h->streams[0]);
  if (!cuda_ok(h, cudaEventSynchronize(h->tev[1]), "timer")) return FDT_ERR_CUDA;
  float t = 0;
  cudaEventElapsedTime(&t, h->tev[0], h->tev[1]);
  if (ms) *ms = t;
  return FDT_OK;
}

int32_t fdt_profile_chunk(fdt_handle* h, const uint8_t* d_frames, int32_t n, int32_t width, int32_t height,
                          int32_t row_stride, int32_t mat_type, int32_t repeats, float* out_ms, int32_t capacity,
                          int32_t* out_launches) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  std::lock_guard<std::mutex> g(h->mu);
  int channels = channels_of(mat_type);
  const int S = (int)h->det.plan().steps.size();
  if (out_launches) *out_launches = S + 2;
  if (!d_frames || n <= 0 || n > h->chunk || channels == 0 || row_stride < width * channels || repeats <= 0)
    return fail(h, FDT_ERR_BAD_ARG, "bad profile arguments");
  if (!out_ms || capacity < S + 2) return fail(h, FDT_ERR_SIZE_MISMATCH, "out_ms too small");
  cudaSetDevice(h->cfg.device);
  if (!ensure_results(h, n)) return FDT_ERR_CUDA;
  const LbTables* tb = get_tables(h, width, height);
  if (!tb) return fail(h, FDT_ERR_CUDA, "letterbox table upload failed");
  std::vector<cudaEvent_t> ev(S + 3);
  for (auto& e : ev) cudaEventCreate(&e);
  std::vector<double> acc(S + 2, 0.0);
  cudaStream_t s = h->streams[0];
  const int S_w = h->det.in_w(), S_h = h->det.in_h();
  for (int r = 0; r < repeats; ++r) {
    LetterboxP lb;
    lb.frames = d_frames; lb.frame_stride = (long long)height * row_stride; lb.row_stride = row_stride; lb.channels = channels;
    lb.src_w = width; lb.src_h = height; lb.out = h->d_lb[0]; lb.dst_w = S_w; lb.dst_h = S_h;
    lb.new_w = tb->lp.new_w; lb.new_h = tb->lp.new_h; lb.pad_top = tb->lp.pad_top; lb.pad_left = tb->lp.pad_left;
    lb.x0 = tb->x0; lb.x1 = tb->x1; lb.ax0 = tb->ax0; lb.ax1 = tb->ax1;
    lb.y0 = tb->y0; lb.y1 = tb->y1; lb.by0 = tb->by0; lb.by1 = tb->by1;
    lb.identity = tb->identity ? 1 : 0;
    cudaEventRecord(ev[0], s);
    launch_letterbox(lb, n, s);
    h->det.run(h->det_ctx[0], h->d_lb[0], n, s, ev.data() + 1);
    const Plan& dp = h->det.plan();
    DecodeP d;
    d.boxes = h->det_ctx[0].outputs[0]; d.boxes_istride = dp.out_elems[0];
    d.scores = h->det_ctx[0].outputs[1]; d.scores_istride = dp.out_elems[1];
    d.anchors = h->d_anchors; d.N = h->num_anchors; d.input_h = S_h;
    d.raw_thresh = std::log(kMinScore / (1.0 - kMinScore));
    d.score_thresh = kMinScore; d.iou_thresh = kMinSuppression;
    d.pad_t = (double)tb->lp.pad_top / S_h; d.pad_b = (double)tb->lp.pad_bottom / S_h;
    d.pad_l = (double)tb->lp.pad_left / S_w; d.pad_r = (double)tb->lp.pad_right / S_w;
    d.min_score = h->cfg.min_score; d.min_face_size = h->cfg.min_face_size;
    d.img_w = width; d.img_h = height; d.max_faces = h->max_faces;
    d.faces = h->d_faces; d.counts = h->d_counts;
    d.cand_idx = nullptr; d.cand_cap = 0; d.cand_n = nullptr;
    launch_decode_nms(d, n, s);
    cudaEventRecord(ev[S + 2], s);
    if (!cuda_ok(h, cudaStreamSynchronize(s), "profile")) { for (auto& e : ev) cudaEventDestroy(e); return FDT_ERR_CUDA; }
    for (int i = 0; i < S + 2; ++i) {
      float ms = 0;
      cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
      acc[i] += ms;
    }
  }
  for (int i = 0; i < S + 2; ++i) out_ms[i] = (float)(acc[i] / repeats);
  for (auto& e : ev) cudaEventDestroy(e);
  return FDT_OK;
}

int32_t fdt_get_step_info(fdt_handle* h, int32_t launch, char* kernel, char* tensor, int32_t str_cap, double* macs_per_image,
                          double* bytes_per_image) {
  if (!h || !h->ready) return fail(h, FDT_ERR_NOT_READY, "detector is not initialised");
  const Plan& p = h->det.plan();
  const int S = (int)p.steps.size();
  if (launch < 0 || launch > S + 1) return fail(h, FDT_ERR_BAD_ARG, "launch index out of range");
  static const char* kn[] = {"k_normalize", "k_naive_conv", "k_gemm_conv", "k_dwpw", "k_add", "k_act", "k_padc", "k_maxpool", "k_resize_bilinear", "k_stem", "k_dwpw_tc", "k_stem_tc", "k_block_ws", "k_stem_ws"};
  std::string kname, tname;
  double macs = 0, bytes = 0;
  const int S_w = h->det.in_w(), S_h = h->det.in_h();
  if (launch == 0) {
    kname = "k_letterbox"; tname = "letterboxed_u8";
    bytes = 0;  // depends on the frame size: 4 taps x 3 B per content pixel read + S*S*3 written (filled by the caller)
  } else if (launch == S + 1) {
    kname = "k_decode_nms"; tname = "faces";
    bytes = (double)h->num_anchors * 17 * 4;
  } else {
    const PStep& st = p.steps[launch - 1];
    kname = kn[st.kind]; tname = st.name; macs = st.macs;
    const PTensor& o = p.tensors[st.out];
    bytes = (double)o.H * o.W * o.C * 4;
    if (st.in_u8) bytes += (double)S_w * S_h * 4;
    else if (st.in >= 0) { const PTensor& i = p.tensors[st.in]; bytes += (double)i.H * i.W * i.C * 4; }
    if (st.in2 >= 0 && st.in2 != st.in) { const PTensor& r = p.tensors[st.in2]; bytes += (double)r.H * r.W * r.C * 4; }
  }
  if (kernel && str_cap > 0) { std::snprintf(kernel, str_cap, "%s", kname.c_str()); }
  if (tensor && str_cap > 0) { std::snprintf(tensor, str_cap, "%s", tname.c_str()); }
  if (macs_per_image) *macs_per_image = macs;
  if (bytes_per_image) *bytes_per_image = bytes;
  return FDT_OK;
}

int32_t fdt_host_anchors(int32_t model, double* out_xy, int32_t capacity_pairs) {
  std::vector<double> a = generate_anchors(ssd_options(model));
  int n = (int)a.size() / 2;
  if (out_xy) std::memcpy(out_xy, a.data(), (size_t)std::min(n, capacity_pairs) * 2 * sizeof(double));
  return n;
}

int32_t fdt_host_plan_describe(const uint8_t* tflite, size_t len, int32_t fuse_level, char* buf, size_t buf_len) {
  TfModel m;
  std::string err;
  Plan p;
  std::string text;
  int rc = FDT_OK;
  if (!m.parse(tflite, len, &err) || !p.build(m, fuse_level, &err)) { text = err; rc = FDT_ERR_MODEL; }
  else text = p.describe();
  if (buf && buf_len) {
    size_t n = std::min(buf_len - 1, text.size());
    std::memcpy(buf, text.data(), n);
    buf[n] = 0;
  }
  return rc;
}

int32_t fdt_host_resize_taps(int32_t src, int32_t dst, int32_t is_x_axis, int32_t* i0, int32_t* i1, int16_t* w0, int16_t* w1) {
  if (src <= 0 || dst <= 0 || !i0 || !i1 || !w0 || !w1) return FDT_ERR_BAD_ARG;
  resize_linear_taps(src, dst, is_x_axis != 0, i0, i1, w0, w1);
  return FDT_OK;
}

int32_t fdt_host_decode_box(const float* raw16, double ax, double ay, double scale, double* out_box4, double* out_kp12) {
  if (!raw16 || !out_box4 || !out_kp12) return FDT_ERR_BAD_ARG;
  decode_box(raw16, ax, ay, scale, out_box4, out_kp12);
  return FDT_OK;
}

int32_t fdt_host_face_roi(const double* kp12, double img_w, double img_h, int32_t out_size, double* out10) {
  if (!kp12 || !out10) return 0;
  face_alignment(kp12, img_w, img_h, &out10[0], &out10[1], &out10[2], &out10[3]);
  return aligned_square_inverse(out10[1], out10[2], out10[3], -out10[0], out_size, &out10[4]) ? 1 : 0;
}

const char* fdt_last_error(fdt_handle* h) {
  if (h) return h->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> g(g_create_mu);
  copy = g_create_error;
  return copy.c_str();
}

const char* fdt_version(void) { return "fdt-cuda 0.1.0 (sm_100a)"; }

}  // extern "C"
