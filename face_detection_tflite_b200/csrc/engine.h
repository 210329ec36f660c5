// Engine: one TFLite model lowered to a Plan, with weights resident in HBM.
// EngineCtx: the per-stream activation arena and output buffers for a chunk of images.
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "kernels.h"
#include "plan.h"
#include "tflite_model.h"

namespace fdt {

struct EngineCtx {
  float* arena = nullptr;               // [cap][arena_per_image]
  std::vector<float*> outputs;          // per graph output: [cap][out_elems]
  int cap = 0;
};

class Engine {
 public:
  ~Engine();
  bool init(const uint8_t* tflite, size_t len, int fuse_level, std::string* err, bool use_tc = true);
  bool make_ctx(int cap, EngineCtx* ctx, std::string* err) const;
  void free_ctx(EngineCtx* ctx) const;
  // Runs the plan on B <= ctx.cap images whose u8 BGR input lives at `in_u8` ([B][H][W][3]).
  // Returns the number of kernel launches.
  // `evs` (optional): steps()+1 events, evs[i] recorded before step i and evs[steps()] after the last.
  int run(const EngineCtx& ctx, const uint8_t* in_u8, int B, cudaStream_t s, cudaEvent_t* evs = nullptr) const;
  TV view(const EngineCtx& ctx, int ptensor) const;

  // true once a launch could not be issued (tensor map encoding failed): results are invalid, callers must fail
  bool failed() const { return failed_; }
  const Plan& plan() const { return plan_; }
  const TfModel& model() const { return model_; }
  int in_h() const { return plan_.in_h; }
  int in_w() const { return plan_.in_w; }

 private:
  TfModel model_;
  Plan plan_;
  float* d_blob_ = nullptr;
  std::vector<TailLayerD*> d_tail_;     // per plan step: device copy of a k_tail_ws layer program (or null)
  std::vector<TailBlk*> d_blks_;        // per plan step: device copy of a k_chain_wide W block table (or null)
  int max_ctas_ = 148 * 4;
  mutable bool failed_ = false;
};

}  // namespace fdt
