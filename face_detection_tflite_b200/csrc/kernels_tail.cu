// k_tail_ws — the low-resolution tail of a BlazeFace detector (every 16x16 / 8x8 BlazeBlock and both head pairs)
// in ONE launch with the activations of an image resident in shared memory.
//
// Launched layer by layer these steps move ~1.2 MB per image through HBM at 1-4 tiles per CTA and are bound by launch
// prologues and pipeline fill, not by bandwidth (round-1 profile: 12 launches = 32 % of the step at 0.16-0.31 of the
// HBM roofline).  Here a persistent CTA per SM takes one image at a time:
//
//   input (16x16xC, HBM) --TMA--> buffer A [256 px][KSA] -- blocks @16x16 (in place) --> heads@16x16 -> HBM outputs
//                                  '-- stride-2 block --> buffer B [64 px][KSB] -- blocks @8x8 (in place) --> heads@8x8 -> HBM
//
//   warps 0..11 (compute): per layer, thread = (TMEM lane, a third of the channel quads).  A 16x16 map is two M-tiles, its
//       even and its odd rows, so a thread owns two vertically adjacent pixels and feeds both depthwise outputs from one
//       4 x 3 window of loads; an 8x8 map is one M-tile.  depthwise 3x3 from shared memory (fp32 FFMA2) -> fp16 hi + lo
//       split -> tcgen05.st straight into TENSOR MEMORY (the A operand never touches shared memory), in two K halves so
//       that the first half's MMAs overlap the second half's depthwise; then the epilogue: tcgen05.ld -> + bias
//       + residual (same pixel / 2x2 max-pool of the source buffer) -> ReLU -> back into the activation buffer
//       (heads: + bias -> graph outputs in HBM).
//   warp 12, one lane (control): TMA of the next image, one bulk copy per layer of its weight record
//       [W fp16 | depthwise taps | depthwise bias] into a two-deep ring (the next layer's weights stream in from L2
//       during the current layer), and the tcgen05.mma issue: kind::f16, A from TMEM, W from shared memory, two
//       passes (A_hi, A_lo) - the detector's weights are fp16-origin and therefore exact, so the products carry
//       ~22 mantissa bits like the TF32 hi/lo split of k_block_ws at half the MMA count.
//
// Hand-offs are mbarriers: w_full[2] (weights landed), a_full[2] (first / second K half of the operand written, one
// arrival per warp), d_full (tcgen05.commit), in_full / a_free (image buffer).  One named barrier per layer orders the in-place epilogue
// against the next layer's depthwise reads.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>

#include "kernels.h"

namespace fdt {
namespace {

#ifndef FDT_TAIL_WARPS
#define FDT_TAIL_WARPS 12                      // measured: 16 warps (96 registers, spills) run the detector tail 9 % slower
#endif
constexpr int kTC = FDT_TAIL_WARPS;           // compute warps (a multiple of 4: one group of kTG warps per TMEM lane quarter)
constexpr int kTG = kTC / 4;
constexpr int kTComputeThreads = kTC * 32;
constexpr int kTThreads = (kTC + 1) * 32;
// TMEM columns: accumulators of tile 0 / 1 at 0 / 128 (Npad <= 128); operand of tile 0 / 1 at 256 / 384: hi halves at +0,
// lo halves at +64 (K16 / 2 <= 64 columns)
__device__ __forceinline__ uint32_t col_d(int t) { return t ? 128u : 0u; }
__device__ __forceinline__ uint32_t col_a(int t) { return t ? 384u : 256u; }
constexpr uint32_t kLBO = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(uint32_t bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
}
// Wait of the single-lane control role, with a suspend-time hint: the bare try_wait loop re-issues YIELD / TRYWAIT / BRA every ~16
// cycles - 12 % of all issued instructions of k_tail_ws<0> (ncu source counters) on a scheduler that also hosts three compute warps.
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
// Plain (non-volatile) shared-memory load: the compiler may batch these ahead of the FMAs that use them (the depthwise
// loop is LDS-latency bound with three warps per scheduler); mbarrier waits / named barriers carry "memory" clobbers, so
// no load moves across a hand-off.
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void sts4(uint32_t addr, const float4& v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// two packed fp32 FMAs (sm_100 FFMA2): same rounding as four scalar fmaf
__device__ __forceinline__ void fma4(float4& a, const float4& v, const float4& w) {
  asm("{\n\t.reg .b64 ra, rv, rw;\n\t"
      "mov.b64 ra, {%0, %1};\n\tmov.b64 rv, {%4, %5};\n\tmov.b64 rw, {%8, %9};\n\t"
      "fma.rn.f32x2 ra, rv, rw, ra;\n\tmov.b64 {%0, %1}, ra;\n\t"
      "mov.b64 ra, {%2, %3};\n\tmov.b64 rv, {%6, %7};\n\tmov.b64 rw, {%10, %11};\n\t"
      "fma.rn.f32x2 ra, rv, rw, ra;\n\tmov.b64 {%2, %3}, ra;\n\t}"
      : "+f"(a.x), "+f"(a.y), "+f"(a.z), "+f"(a.w)
      : "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "f"(w.x), "f"(w.y), "f"(w.z), "f"(w.w));
}
__device__ __forceinline__ float4 max4(const float4& a, const float4& b) {
  return make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
}
// (e0, e1) -> packed fp16 pair, low half = e0; saturating (an activation beyond the fp16 range must not turn into inf)
__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
// a = hi + lo with hi, lo fp16: ~22 mantissa bits; written as two 2-column TMEM stores (4 halves each)
__device__ __forceinline__ void split_store_tmem(uint32_t taddr_hi, const float4& a) {
  const uint32_t h0 = pack_h2(a.x, a.y), h1 = pack_h2(a.z, a.w);
  const __half2 hh0 = *reinterpret_cast<const __half2*>(&h0), hh1 = *reinterpret_cast<const __half2*>(&h1);
  const float2 f0 = __half22float2(hh0), f1 = __half22float2(hh1);
  const uint32_t l0 = pack_h2(a.x - f0.x, a.y - f0.y), l1 = pack_h2(a.z - f1.x, a.w - f1.y);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi), "r"(h0), "r"(h1) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi + 64u), "r"(l0), "r"(l1) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
}

// Shared-memory carve-up (must match tail_smem_bytes below):
//   [layers 16 x 128 B] [bias 16 x 128 floats] [slopes 16 x 128 floats] [barriers 16 x 8 B] [scratch 16 floats]
//   | 128-byte aligned: [activation buffers: act_floats] [zero slot 128 floats]
//   | 128-byte aligned: [depthwise-record ring 2 x tbuf_bytes] [pointwise-weight ring wdepth x wbuf_bytes]
// GEN = false: the detector tails (blocks with ReLU + head pairs, exact fp16 weights, residual = the block's own input);
// GEN = true adds what the face-landmark chain needs (PReLU, W = hi + lo, weight scale, pointwise-only and dot layers, residual
// from another buffer, results mirrored to HBM).
// Build with FDT_NVCC_FLAGS=-DFDT_TAIL_TRACE: warp 0 of CTA 0 stamps clock64 at the phases of every layer of its second image and
// prints the per-layer cycle counts (wait for taps / depthwise + operand stores / wait for the MMAs / epilogue / layer barrier).
#ifdef FDT_TAIL_TRACE
#define TT(k) do { if (blockIdx.x == 0 && tid == 0 && it == 1) tt[l * 6 + (k)] = clock64(); } while (0)
#else
#define TT(k) do { } while (0)
#endif
template <bool GEN>
__global__ void __launch_bounds__(kTThreads, 1) k_tail_ws(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ TailP p, int B) {
#ifdef FDT_TAIL_TRACE
  __shared__ long long tt[16 * 6];
#endif
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TailLayerD* sL = reinterpret_cast<TailLayerD*>(smem_raw);
  float* sBias = reinterpret_cast<float*>(smem_raw + kTailMaxLayers * sizeof(TailLayerD));
  float* sAlpha = sBias + kTailMaxLayers * 128;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAlpha + kTailMaxLayers * 128);
  float* scratch = reinterpret_cast<float*>(bars + 16);
  float* act0 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch + 16) + 127) & ~(uintptr_t)127);
  float* zslot = act0 + p.act_floats;
  unsigned char* tring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(zslot + 128) + 127) & ~(uintptr_t)127);
  unsigned char* wring = tring + 2 * (size_t)p.tbuf_bytes;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t in_full = bar0, a_free = bar0 + 8u, w_full = bar0 + 16u, a_full = bar0 + 32u, d_full = bar0 + 48u, t_full = bar0 + 56u;
  const int nl = p.nlayers;
  const int my_images = blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- prologue ---------------------------------------------------------------------------------------------------
  for (int i = tid; i < nl * (int)(sizeof(TailLayerD) / 16); i += kTThreads)
    reinterpret_cast<uint4*>(sL)[i] = reinterpret_cast<const uint4*>(p.layers)[i];
  // every buffer but the first (the TMA fills that one, zero channel pad included) starts as zeros: channel pads stay zero
  for (int i = (p.in_bytes >> 2) + tid; i < p.act_floats + 128; i += kTThreads) act0[i] = 0.f;
  if (warp == kTC) {
    if (lane == 0) {
      mbar_init(in_full, 1);
      mbar_init(a_free, kTC);
      for (int i = 0; i < 2; ++i) { mbar_init(w_full + 8u * i, 1); mbar_init(a_full + 8u * i, kTC); mbar_init(t_full + 8u * i, 1); }
      mbar_init(d_full, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  __syncthreads();
  for (int l = 0; l < nl; ++l)
    for (int i = tid; i < sL[l].Npad; i += kTThreads) {
      sBias[l * 128 + i] = p.blob[(size_t)sL[l].bias_off + i];
      if (GEN && sL[l].act == 2) sAlpha[l * 128 + i] = p.blob[(size_t)sL[l].alpha_off + i];
    }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t act_a = smem_u32(act0), zero_a = smem_u32(zslot), wring_a = smem_u32(wring), tring_a = smem_u32(tring);

  if (warp < kTC) {
    // =============================== compute warps ==============================================================
    const int lq = warp & 3, g = warp >> 2;                // TMEM lane quarter (hardware: warp % 4), channel-quad / column group
    const int r = lq * 32 + lane;                           // row of the 128-pixel M-tile = TMEM lane
    const uint32_t tm_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    uint32_t ph_d = 0u, ph_t0 = 0u, ph_t1 = 0u;
    int gl = 0, it = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      mbar_wait(in_full, (uint32_t)(it & 1));
      for (int l = 0; l < nl; ++l, ++gl) {
        const TailLayerD L = sL[l];                           // registers: the loops below must not re-read it from shared memory
        TT(0);
        const uint32_t src_a = act_a + 4u * (uint32_t)p.buf_off[L.src], kss_b = 4u * (uint32_t)p.buf_ks[L.src];
        const uint32_t dst_a = L.dst >= 0 ? act_a + 4u * (uint32_t)p.buf_off[L.dst] : 0u, ksd_b = L.dst >= 0 ? 4u * (uint32_t)p.buf_ks[L.dst] : 0u;
        const int src_q = (int)(kss_b >> 4), dst_q = (int)(ksd_b >> 4);       // quads per pixel record
        const int npix = L.OH * L.OW;
        const uint32_t tb_a = tring_a + (uint32_t)(gl & 1) * (uint32_t)p.tbuf_bytes;
        if (L.tap_bytes) {                                    // this layer's depthwise record (kind 3: its filter) has landed
          if (gl & 1) { mbar_wait(t_full + 8u, ph_t1); ph_t1 ^= 1u; } else { mbar_wait(t_full, ph_t0); ph_t0 ^= 1u; }
        }
        TT(1);
        if (GEN && L.kind == 3) {
          // ---- whole-map dot product (a k x k VALID convolution over a k x k map with one filter) ------------------------
          const int n = npix * L.Cin;
          float sacc = 0.f;
          for (int i = tid; i < n; i += kTComputeThreads) {
            const int px = i / L.Cin, c = i - px * L.Cin;
            float a, w;
            asm("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(src_a + (uint32_t)px * kss_b + 4u * (uint32_t)c) : "memory");
            asm("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(tb_a + 4u * (uint32_t)i) : "memory");
            sacc = fmaf(a, w, sacc);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
          if (lane == 0) { scratch[warp] = sacc; mbar_arrive(a_full); mbar_arrive(a_full + 8u); }    // the filter and the map are read
          asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
          if (tid == 0) {
            float tot = sBias[l * 128];
            for (int w2 = 0; w2 < kTC; ++w2) tot += scratch[w2];
            p.outs[L.o1][(size_t)img * p.out_istride[L.o1]] = tot;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
          if (l == p.last_a_layer && lane == 0) mbar_arrive(a_free);
          continue;
        }
        // Maps of more than 128 pixels (even height, at most 256 pixels) are split into the M-tiles "even rows" / "odd rows":
        // a thread then owns the vertically adjacent pixels (2yy, x) and (2yy + 1, x) and feeds both depthwise outputs from
        // ONE 4 x 3 window (6 instead of 9 LDS.128 per output quad); smaller maps are one M-tile in row-major order
        const bool paired = npix > 128;
        const int ntiles = paired ? 2 : 1;
        const int rows = paired ? (npix >> 1) : npix;           // rows of an M-tile that hold pixels
        const int nq = L.K16 >> 2, nq_real = (L.Cin + 3) >> 2, nq_half = (L.K16 >> 5) << 2;   // quads of the first K half (whole 16-wide K steps)
        const uint32_t dww_a = tb_a, dwb_a = dww_a + 9u * (uint32_t)L.K16 * 4u;
        const uint32_t k16_b = (uint32_t)L.K16 * 4u;
        const uint32_t acol0 = tm_lane + col_a(0), acol1 = tm_lane + col_a(1);
        // this thread's output pixel(s): (oy, ox) of tile 0; tile 1 is (oy + 1, ox) when paired
        int oy = r / L.OW;
        const int ox = r - oy * L.OW;
        if (paired) oy *= 2;
        const bool act = r < rows;
        const bool warp_on = lq * 32 < rows;                  // lane quarters without pixels only keep the hand-offs going
        const int pix0 = oy * L.OW + ox;
        // ---- operand tiles: depthwise 3x3 (blocks) or the activation itself (pointwise layers) -> fp16 hi / lo -> TMEM ----
        // The K range is produced in two halves with an mbarrier each, so that the MMAs of the first half run while the
        // second half is still being computed.
        bool half_done = false;
        if (!warp_on) {
        } else if (L.kind == 0) {
          // window offsets: rows oy*s - pad .. (+2, +3 when paired), cols ox*s - pad .. +2; SAME padding reads the zero slot
          uint32_t off[12];
          const int iy0 = oy * L.stride - L.pad, ix0 = ox * L.stride - L.pad;
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const int iy = iy0 + k / 3, ix = ix0 + k % 3;
            const bool ok = act && (unsigned)iy < (unsigned)L.IH && (unsigned)ix < (unsigned)L.IW;
            off[k] = ok ? src_a + (uint32_t)(iy * L.IW + ix) * kss_b : zero_a;
          }
          for (int q = g; q < nq; q += kTG) {
            if (!half_done && q >= nq_half) {
              half_done = true;
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(a_full);
            }
            float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
            if (q < nq_real) {
              const uint32_t qo = 16u * (uint32_t)q;
              float4 w[9];
#pragma unroll
              for (int k = 0; k < 9; ++k) w[k] = lds4(dww_a + (uint32_t)k * k16_b + qo);
              const float4 bias = lds4(dwb_a + qo);
              float4 v[3];
              // each output: bias, then the taps in (ky, kx) order
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = lds4(off[c] + qo);
              a0 = bias;
#pragma unroll
              for (int c = 0; c < 3; ++c) fma4(a0, v[c], w[c]);
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = lds4(off[3 + c] + qo);
              a1 = bias;
#pragma unroll
              for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[3 + c]); if (paired) fma4(a1, v[c], w[c]); }
#pragma unroll
              for (int c = 0; c < 3; ++c) v[c] = lds4(off[6 + c] + qo);
#pragma unroll
              for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[6 + c]); if (paired) fma4(a1, v[c], w[3 + c]); }
              if (paired) {
#pragma unroll
                for (int c = 0; c < 3; ++c) v[c] = lds4(off[9 + c] + qo);
#pragma unroll
                for (int c = 0; c < 3; ++c) fma4(a1, v[c], w[6 + c]);
              }
            }
            split_store_tmem(acol0 + 2u * (uint32_t)q, a0);
            if (paired) split_store_tmem(acol1 + 2u * (uint32_t)q, a1);
          }
        } else {
          const uint32_t px0 = act ? src_a + (uint32_t)pix0 * kss_b : zero_a;
          const uint32_t px1 = px0 + (uint32_t)L.OW * kss_b;
          for (int q = g; q < nq; q += kTG) {
            if (!half_done && q >= nq_half) {
              half_done = true;
              asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
              asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
              __syncwarp();
              if (lane == 0) mbar_arrive(a_full);
            }
            const bool in = q < src_q;
            split_store_tmem(acol0 + 2u * (uint32_t)q, in ? lds4(px0 + 16u * (uint32_t)q) : make_float4(0.f, 0.f, 0.f, 0.f));
            if (paired) split_store_tmem(acol1 + 2u * (uint32_t)q, in ? lds4(px1 + 16u * (uint32_t)q) : make_float4(0.f, 0.f, 0.f, 0.f));
          }
        }
        if (warp_on) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        TT(2);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) { if (!half_done) mbar_arrive(a_full); mbar_arrive(a_full + 8u); }
        // ---- epilogue ---------------------------------------------------------------------------------------------------
        // In-place layers: every depthwise read of the layer (all threads) must precede the first write; d_full implies it
        // (the MMAs were issued after both operand barriers completed).
        if (warp_on) {
          const uint32_t bias_a = smem_u32(sBias + l * 128), alpha_a = smem_u32(sAlpha + l * 128);
          const uint32_t rsrc_a = GEN ? act_a + 4u * (uint32_t)p.buf_off[L.rbuf] : src_a, ksr_b = GEN ? 4u * (uint32_t)p.buf_ks[L.rbuf] : kss_b;
          const int res_q = (int)(ksr_b >> 4);
          mbar_wait(d_full, ph_d);
          TT(3);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t row_b = (uint32_t)L.IW * ksr_b;
#pragma unroll 1
          for (int t = 0; t < ntiles; ++t) {
            const int pix = pix0 + t * L.OW;                     // tile 1 = the row below
            const uint32_t dcol = tm_lane + col_d(t);
            // residual source pixel(s)
            uint32_t res_a = zero_a;
            if (act && L.res == 1) res_a = rsrc_a + (uint32_t)pix * ksr_b;
            if (act && L.res == 2) res_a = rsrc_a + (uint32_t)((2 * oy) * L.IW + 2 * ox) * ksr_b;
            const uint32_t out_a = dst_a + (uint32_t)(act ? pix : 0) * ksd_b;
            float* o1 = nullptr; float* o2 = nullptr;
            int o1_q = 0;
            if (GEN ? L.o1 >= 0 : L.kind == 1) {
              o1 = p.outs[L.o1] + (size_t)img * p.out_istride[L.o1] + (size_t)(act ? pix : 0) * p.out_pix[L.o1];
              o1_q = p.out_pix[L.o1] >> 2;
              if (L.o2 >= 0) o2 = p.outs[L.o2] + (size_t)img * p.out_istride[L.o2] + (size_t)(act ? pix : 0) * p.out_pix[L.o2];
            }
            for (int c16 = g; c16 < (L.Npad >> 4); c16 += kTG) {
              uint32_t u[16];
              asm volatile(
                  "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                  : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                    "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                  : "r"(dcol + 16u * (uint32_t)c16));
              float4 bv[4], rv[4];
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cq = 4 * c16 + j;
                const uint32_t qo = 16u * (uint32_t)cq;
                bv[j] = lds4(bias_a + qo);
                rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (L.res == 1) {
                  if (cq < res_q) rv[j] = lds4(res_a + qo);
                } else if (L.res == 2) {
                  if (cq < res_q) rv[j] = max4(max4(lds4(res_a + qo), lds4(res_a + ksr_b + qo)), max4(lds4(res_a + row_b + qo), lds4(res_a + row_b + ksr_b + qo)));
                }
              }
              asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int cq = 4 * c16 + j;
                float4 v;
                if (GEN) v = make_float4(fmaf(__uint_as_float(u[4 * j]), L.wscale, bv[j].x), fmaf(__uint_as_float(u[4 * j + 1]), L.wscale, bv[j].y),
                                         fmaf(__uint_as_float(u[4 * j + 2]), L.wscale, bv[j].z), fmaf(__uint_as_float(u[4 * j + 3]), L.wscale, bv[j].w));
                else v = make_float4(__uint_as_float(u[4 * j]) + bv[j].x, __uint_as_float(u[4 * j + 1]) + bv[j].y,
                                     __uint_as_float(u[4 * j + 2]) + bv[j].z, __uint_as_float(u[4 * j + 3]) + bv[j].w);
                if (L.kind != 1) {
                  v.x += rv[j].x; v.y += rv[j].y; v.z += rv[j].z; v.w += rv[j].w;
                  if (L.act == 1) {
                    v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
                  } else if (GEN && L.act == 2) {
                    const float4 al = lds4(alpha_a + 16u * (uint32_t)cq);
                    v = make_float4(fmaf(al.x, fminf(v.x, 0.f), fmaxf(v.x, 0.f)), fmaf(al.y, fminf(v.y, 0.f), fmaxf(v.y, 0.f)),
                                    fmaf(al.z, fminf(v.z, 0.f), fmaxf(v.z, 0.f)), fmaf(al.w, fminf(v.w, 0.f), fmaxf(v.w, 0.f)));
                  }
                  if (act && cq < dst_q) sts4(out_a + 16u * (uint32_t)cq, v);       // columns >= Cout come out as exact zeros
                  if (GEN && act && cq < o1_q) *reinterpret_cast<float4*>(o1 + 4 * cq) = v;   // the tensor is also read outside the chain
                } else if (act) {
                  const int c = 4 * cq;
                  if (c + 4 <= L.c1) {
                    *reinterpret_cast<float4*>(o1 + c) = v;
                  } else {
                    const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                      if (c + e >= L.c1 && c + e - L.c1 < L.c2) o2[c + e - L.c1] = vv[e];
                  }
                }
              }
            }
          }
        }
        ph_d ^= 1u;
        TT(4);
        // the layer's writes (shared memory: generic proxy; TMEM reads done) before anybody starts the next layer
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
        TT(5);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (l == p.last_a_layer && lane == 0) mbar_arrive(a_free);              // buffer 0 may take the next image
      }
#ifdef FDT_TAIL_TRACE
      if (blockIdx.x == 0 && tid == 0 && it == 1)
        for (int l2 = 0; l2 < nl; ++l2)
          printf("tail layer %2d: taps-wait %5lld  depthwise %6lld  mma-wait %5lld  epilogue %5lld  barrier %5lld  | total %6lld\n", l2, tt[l2 * 6 + 1] - tt[l2 * 6],
                 tt[l2 * 6 + 2] - tt[l2 * 6 + 1], tt[l2 * 6 + 3] - tt[l2 * 6 + 2], tt[l2 * 6 + 4] - tt[l2 * 6 + 3], tt[l2 * 6 + 5] - tt[l2 * 6 + 4],
                 tt[l2 * 6 + 5] - tt[l2 * 6]);
#endif
    }
  } else if (lane == 0) {
    // =============================== control lane: TMA + weight rings + MMA issue =====================================
    const int total = my_images * nl;
    const int D = p.wdepth;
    auto load_weights = [&](int glayer) {                      // pointwise record of (global) layer glayer -> ring slot glayer % D
      const TailLayerD& L = sL[glayer % nl];
      if (!L.rec_bytes) return;
      const uint32_t slot = (uint32_t)(glayer % D), bar = w_full + 8u * slot;
      mbar_expect_tx(bar, (uint32_t)L.rec_bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(wring_a + slot * (uint32_t)p.wbuf_bytes), "l"(p.blob + (size_t)L.rec_off), "r"((uint32_t)L.rec_bytes), "r"(bar) : "memory");
    };
    auto load_taps = [&](int glayer) {                         // depthwise record -> ring slot glayer & 1
      const TailLayerD& L = sL[glayer % nl];
      if (!L.tap_bytes) return;
      const uint32_t slot = (uint32_t)(glayer & 1), bar = t_full + 8u * slot;
      mbar_expect_tx(bar, (uint32_t)L.tap_bytes);
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                   ::"r"(tring_a + slot * (uint32_t)p.tbuf_bytes), "l"(p.blob + (size_t)L.tap_off), "r"((uint32_t)L.tap_bytes), "r"(bar) : "memory");
    };
    auto load_image = [&](int img) {
      mbar_expect_tx(in_full, (uint32_t)p.in_bytes);
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(act_a), "l"(&tmap), "r"(0), "r"(0), "r"(0), "r"(img), "r"(in_full) : "memory");
    };
    if (my_images > 0) {
      load_image(blockIdx.x);
      for (int i = 0; i < D && i < total; ++i) load_weights(i);
      for (int i = 0; i < 2 && i < total; ++i) load_taps(i);
    }
    uint32_t ph_a = 0u, ph_d = 0u, ph_w0 = 0u, ph_w1 = 0u;
    int gl = 0, it = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      for (int l = 0; l < nl; ++l, ++gl) {
        const TailLayerD& L = sL[l];
        if (!GEN || L.kind != 3) {
          const int ntiles = L.OH * L.OW > 128 ? 2 : 1;
          const uint32_t idesc = (1u << 4) | ((uint32_t)(L.Npad >> 3) << 17) | ((128u >> 4) << 24);     // D f32, A / B f16, K-major
          const uint32_t sbo = (uint32_t)(L.K16 >> 3) * 128u;
          const uint32_t b_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
          const int slot = gl % D;
          const uint32_t wb_a = wring_a + (uint32_t)slot * (uint32_t)p.wbuf_bytes;
          const uint32_t b_lo0 = ((wb_a & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
          const uint32_t b_lo1 = (((wb_a + (uint32_t)(L.Npad * L.K16) * 2u) & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);   // W_lo (w_parts == 2)
          const int ksteps = L.K16 >> 4, khalf = L.K16 >> 5;          // K steps of the first operand half
          if (slot) { mbar_wait_idle(w_full + 8u, ph_w1); ph_w1 ^= 1u; } else { mbar_wait_idle(w_full, ph_w0); ph_w0 ^= 1u; }
          for (int half = 0; half < 2; ++half) {
            mbar_wait_idle(a_full + 8u * (uint32_t)half, ph_a);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int k0 = half ? khalf : 0, k1 = half ? ksteps : khalf;
            for (int t = 0; t < ntiles; ++t) {
              const uint32_t dcol = tmem_base + col_d(t), acol = tmem_base + col_a(t);
#pragma unroll 1
              for (int ks = k0; ks < k1; ++ks) {
                mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, ks ? 1u : 0u);
                mma_ts_f16(dcol, acol + 64u + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
                if (GEN && L.w_parts == 2) mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, b_lo1 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
              }
            }
          }
          ph_a ^= 1u;
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(d_full) : "memory");
          // this layer's MMAs have read the weight buffer and every thread has read its taps: their ring slots are free
          mbar_wait_idle(d_full, ph_d);
          ph_d ^= 1u;
        } else {
          mbar_wait_idle(a_full, ph_a);
          mbar_wait_idle(a_full + 8u, ph_a);
          ph_a ^= 1u;
        }
        if (gl + D < total) load_weights(gl + D);
        if (gl + 2 < total) load_taps(gl + 2);
        if (l == p.last_a_layer && img + (int)gridDim.x < B) {
          mbar_wait_idle(a_free, (uint32_t)(it & 1));
          load_image(img + (int)gridDim.x);
        }
      }
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kTC) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}


// ---- k_chain_wide: the same image-resident layer program for layers wider than 128 channels ---------------------------------
// (full-range BlazeFace below 24x24: 12x12x144 ... 256 and 6x6x384).  What changes against k_tail_ws<true>:
//   * the pointwise product of a layer runs as column groups x K chunks.  A map of two M-tiles (129..256 pixels) keeps 128
//     accumulator columns per tile and recomputes its (narrow) operand per column group; a one-tile map keeps up to 384
//     accumulator columns (TMEM 0..383, operand chunk at 384..511) and walks K in chunks of 128 with the accumulators resident.
//     The operand region is reused chunk after chunk: the compute warps wait for `a_done` (tcgen05.commit after the chunk's MMAs).
//   * W streams as blocks [<= 128 columns][<= 128 K] in consumption order through a ring of wdepth slots (w_full / w_free).
//   * the residual may come from an HBM tensor (same pixel or 2x2 max-pool), the input map is fetched by one bulk copy per
//     pixel (a 256-channel pixel record exceeds the TMA box limit) issued by a loader warp, so the next image's input
//     overlaps the current image's last layers.
constexpr int kWThreads = (kTC + 2) * 32;
constexpr int kWideZero = 512;
__global__ void __launch_bounds__(kWThreads, 1) k_chain_wide(const __grid_constant__ TailP p, int B) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  TailLayerD* sL = reinterpret_cast<TailLayerD*>(smem_raw);
  TailBlk* sBlk = reinterpret_cast<TailBlk*>(smem_raw + kTailMaxLayers * sizeof(TailLayerD));
  float* sBias = reinterpret_cast<float*>(sBlk + kTailMaxBlks);
  float* sAlpha = sBias + p.bias_floats;
  uint64_t* bars = reinterpret_cast<uint64_t*>(sAlpha + p.bias_floats);
  float* scratch = reinterpret_cast<float*>(bars + 32);
  float* act0 = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(scratch + 16) + 127) & ~(uintptr_t)127);
  float* zslot = act0 + p.act_floats;                      // kWideZero zero floats: SAME padding / inactive rows of up to 384 channels
  unsigned char* tring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(zslot + kWideZero) + 127) & ~(uintptr_t)127);
  unsigned char* wring = tring + 2 * (size_t)p.tbuf_bytes;
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t in_full = bar0, a_free = bar0 + 8u, a_full = bar0 + 16u, d_full = bar0 + 32u, t_full = bar0 + 40u, a_done = bar0 + 56u,
                 w_full = bar0 + 64u, w_free = bar0 + 96u;
  const int nl = p.nlayers;
  const int my_images = blockIdx.x < B ? (B - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- prologue ---------------------------------------------------------------------------------------------------
  for (int i = tid; i < nl * (int)(sizeof(TailLayerD) / 16); i += kWThreads)
    reinterpret_cast<uint4*>(sL)[i] = reinterpret_cast<const uint4*>(p.layers)[i];
  for (int i = tid; i < p.nblks; i += kWThreads) reinterpret_cast<uint4*>(sBlk)[i] = reinterpret_cast<const uint4*>(p.blks)[i];
  // every buffer starts as zeros: the loader writes only the real channels of buffer 0, channel pads stay zero for good
  for (int i = tid; i < p.act_floats + kWideZero; i += kWThreads) act0[i] = 0.f;
  if (warp == kTC) {
    if (lane == 0) {
      mbar_init(in_full, 1);
      mbar_init(a_free, kTC);
      for (int i = 0; i < 2; ++i) { mbar_init(a_full + 8u * i, kTC); mbar_init(t_full + 8u * i, 1); }
      mbar_init(d_full, 1);
      mbar_init(a_done, 1);
      for (int i = 0; i < 4; ++i) { mbar_init(w_full + 8u * i, 1); mbar_init(w_free + 8u * i, 1); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  __syncthreads();
  for (int l = 0; l < nl; ++l)
    for (int i = tid; i < sL[l].Npad; i += kWThreads) {
      sBias[sL[l].bias_s + i] = p.blob[(size_t)sL[l].bias_off + i];
      if (sL[l].act == 2) sAlpha[sL[l].bias_s + i] = p.blob[(size_t)sL[l].alpha_off + i];
    }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");       // the zero fill precedes the loader's bulk copies into buffer 0
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const uint32_t act_a = smem_u32(act0), zero_a = smem_u32(zslot), wring_a = smem_u32(wring), tring_a = smem_u32(tring);

  if (warp < kTC) {
    // =============================== compute warps ==============================================================
    const int lq = warp & 3, g = warp >> 2;
    const int r = lq * 32 + lane;
    const uint32_t tm_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    uint32_t ph_d = 0u, ph_t0 = 0u, ph_t1 = 0u;
    int gl = 0, it = 0, gchunk = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      mbar_wait(in_full, (uint32_t)(it & 1));
      for (int l = 0; l < nl; ++l, ++gl) {
        const TailLayerD L = sL[l];
        const uint32_t src_a = act_a + 4u * (uint32_t)p.buf_off[L.src], kss_b = 4u * (uint32_t)p.buf_ks[L.src];
        const uint32_t dst_a = L.dst >= 0 ? act_a + 4u * (uint32_t)p.buf_off[L.dst] : 0u, ksd_b = L.dst >= 0 ? 4u * (uint32_t)p.buf_ks[L.dst] : 0u;
        const int src_q = (int)(kss_b >> 4), dst_q = (int)(ksd_b >> 4);
        const int npix = L.OH * L.OW;
        const uint32_t tb_a = tring_a + (uint32_t)(gl & 1) * (uint32_t)p.tbuf_bytes;
        if (L.tap_bytes) {
          if (gl & 1) { mbar_wait(t_full + 8u, ph_t1); ph_t1 ^= 1u; } else { mbar_wait(t_full, ph_t0); ph_t0 ^= 1u; }
        }
        if (L.kind == 3) {
          const int n = npix * L.Cin;
          float sacc = 0.f;
          for (int i = tid; i < n; i += kTComputeThreads) {
            const int px = i / L.Cin, c = i - px * L.Cin;
            float a, w;
            asm("ld.shared.f32 %0, [%1];" : "=f"(a) : "r"(src_a + (uint32_t)px * kss_b + 4u * (uint32_t)c) : "memory");
            asm("ld.shared.f32 %0, [%1];" : "=f"(w) : "r"(tb_a + 4u * (uint32_t)i) : "memory");
            sacc = fmaf(a, w, sacc);
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) sacc += __shfl_xor_sync(0xffffffffu, sacc, o);
          if (lane == 0) { scratch[warp] = sacc; mbar_arrive(a_full); mbar_arrive(a_full + 8u); }
          asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
          if (tid == 0) {
            float tot = sBias[L.bias_s];
            for (int w2 = 0; w2 < kTC; ++w2) tot += scratch[w2];
            p.outs[L.o1][(size_t)img * p.out_istride[L.o1]] = tot;
          }
          asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
          if (l == p.last_a_layer && lane == 0) mbar_arrive(a_free);
          continue;
        }
        const bool paired = npix > 128;
        const int ntiles = paired ? 2 : 1;
        const int rows = paired ? (npix >> 1) : npix;
        const int nq = L.K16 >> 2, nq_real = (L.Cin + 3) >> 2;
        const uint32_t dww_a = tb_a, dwb_a = dww_a + 9u * (uint32_t)L.K16 * 4u;
        const uint32_t k16_b = (uint32_t)L.K16 * 4u;
        const uint32_t abase = (!paired && L.Npad > 128) ? 384u : 256u;        // operand chunk: hi halves at +0, lo halves at +64
        const uint32_t acol0 = tm_lane + abase, acol1 = tm_lane + abase + 128u;
        int oy = r / L.OW;
        const int ox = r - oy * L.OW;
        if (paired) oy *= 2;
        const bool act = r < rows;
        const bool warp_on = lq * 32 < rows;
        const int pix0 = oy * L.OW + ox;
        uint32_t off[12];
        if (L.kind == 0) {
          const int iy0 = oy * L.stride - L.pad, ix0 = ox * L.stride - L.pad;
#pragma unroll
          for (int k = 0; k < 12; ++k) {
            const int iy = iy0 + k / 3, ix = ix0 + k % 3;
            const bool ok = act && (unsigned)iy < (unsigned)L.IH && (unsigned)ix < (unsigned)L.IW;
            off[k] = ok ? src_a + (uint32_t)(iy * L.IW + ix) * kss_b : zero_a;
          }
        } else {
          off[0] = act ? src_a + (uint32_t)pix0 * kss_b : zero_a;
          off[1] = off[0] + (uint32_t)L.OW * kss_b;
        }
        const int ng = L.ng & 0xff, nbg = L.ng >> 8;
        for (int grp = 0; grp < ng; ++grp) {
          const TailBlk gb0 = sBlk[L.blk0 + grp * L.nk * nbg];
          const int n0 = paired ? gb0.n0 : 0, ncols = paired ? gb0.ncols : L.Npad;
          for (int kc = 0; kc < L.nk; ++kc, ++gchunk) {
            const int q0 = kc * 32, q1 = min(nq, q0 + 32);
            const int q_half = q0 + (((q1 - q0) >> 3) << 2);          // quads of the chunk's first half (whole 16-wide K steps)
            if (gchunk > 0) {                                        // the previous chunk's MMAs have read the operand region
              mbar_wait(a_done, (uint32_t)((gchunk - 1) & 1));
              asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            bool half_done = false;
            if (!warp_on) {
            } else if (L.kind == 0) {
              for (int q = q0 + g; q < q1; q += kTG) {
                if (!half_done && q >= q_half) {
                  half_done = true;
                  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                  __syncwarp();
                  if (lane == 0) mbar_arrive(a_full);
                }
                float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
                if (q < nq_real) {
                  const uint32_t qo = 16u * (uint32_t)q;
                  float4 w[9];
#pragma unroll
                  for (int k = 0; k < 9; ++k) w[k] = lds4(dww_a + (uint32_t)k * k16_b + qo);
                  const float4 bias = lds4(dwb_a + qo);
                  float4 v[3];
#pragma unroll
                  for (int c = 0; c < 3; ++c) v[c] = lds4(off[c] + qo);
                  a0 = bias;
#pragma unroll
                  for (int c = 0; c < 3; ++c) fma4(a0, v[c], w[c]);
#pragma unroll
                  for (int c = 0; c < 3; ++c) v[c] = lds4(off[3 + c] + qo);
                  a1 = bias;
#pragma unroll
                  for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[3 + c]); if (paired) fma4(a1, v[c], w[c]); }
#pragma unroll
                  for (int c = 0; c < 3; ++c) v[c] = lds4(off[6 + c] + qo);
#pragma unroll
                  for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[6 + c]); if (paired) fma4(a1, v[c], w[3 + c]); }
                  if (paired) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) v[c] = lds4(off[9 + c] + qo);
#pragma unroll
                    for (int c = 0; c < 3; ++c) fma4(a1, v[c], w[6 + c]);
                  }
                }
                split_store_tmem(acol0 + 2u * (uint32_t)(q - q0), a0);
                if (paired) split_store_tmem(acol1 + 2u * (uint32_t)(q - q0), a1);
              }
            } else {
              for (int q = q0 + g; q < q1; q += kTG) {
                if (!half_done && q >= q_half) {
                  half_done = true;
                  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                  __syncwarp();
                  if (lane == 0) mbar_arrive(a_full);
                }
                const bool in = q < src_q;
                split_store_tmem(acol0 + 2u * (uint32_t)(q - q0), in ? lds4(off[0] + 16u * (uint32_t)q) : make_float4(0.f, 0.f, 0.f, 0.f));
                if (paired) split_store_tmem(acol1 + 2u * (uint32_t)(q - q0), in ? lds4(off[1] + 16u * (uint32_t)q) : make_float4(0.f, 0.f, 0.f, 0.f));
              }
            }
            if (warp_on) asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) { if (!half_done) mbar_arrive(a_full); mbar_arrive(a_full + 8u); }
          }
          // ---- epilogue of the column group ----------------------------------------------------------------------------
          if (warp_on) {
            const uint32_t bias_a = smem_u32(sBias + L.bias_s), alpha_a = smem_u32(sAlpha + L.bias_s);
            const bool gsrc = L.res >= 3;
            const uint32_t rsrc_a = gsrc ? 0u : act_a + 4u * (uint32_t)p.buf_off[L.rbuf], ksr_b = gsrc ? 0u : 4u * (uint32_t)p.buf_ks[L.rbuf];
            // channel quads the residual tensor really has (the buffer may be wider and hold an earlier tenant's channels there)
            const int res_q = gsrc ? (p.rs_c[L.rbuf] + 3) >> 2 : min((int)(ksr_b >> 4), (L.c1 + 3) >> 2);
            const float* gbase = gsrc ? p.rsrc[L.rbuf] + (size_t)img * p.rs_istride[L.rbuf] : nullptr;
            const int gcs = gsrc ? p.rs_cs[L.rbuf] : 0, grow = gsrc ? p.rs_w[L.rbuf] * p.rs_cs[L.rbuf] : 0;
            mbar_wait(d_full, ph_d);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t row_b = (uint32_t)L.IW * ksr_b;
#pragma unroll 1
            for (int t = 0; t < ntiles; ++t) {
              const int pix = pix0 + t * L.OW, oyt = oy + t;
              const uint32_t dcol = tm_lane + (paired ? (uint32_t)t * 128u : 0u);
              uint32_t res_a = zero_a;
              if (act && L.res == 1) res_a = rsrc_a + (uint32_t)pix * ksr_b;
              if (act && L.res == 2) res_a = rsrc_a + (uint32_t)((2 * oyt) * L.IW + 2 * ox) * ksr_b;
              const float* gres = gbase;
              if (act && L.res == 3) gres = gbase + (size_t)pix * gcs;
              if (act && L.res == 4) gres = gbase + ((size_t)(2 * oyt) * p.rs_w[L.rbuf] + 2 * ox) * gcs;
              const uint32_t out_a = dst_a + (uint32_t)(act ? pix : 0) * ksd_b;
              float* o1 = nullptr; float* o2 = nullptr;
              int o1_q = 0;
              if (L.o1 >= 0) {
                o1 = p.outs[L.o1] + (size_t)img * p.out_istride[L.o1] + (size_t)(act ? pix : 0) * p.out_pix[L.o1];
                o1_q = p.out_pix[L.o1] >> 2;
                if (L.o2 >= 0) o2 = p.outs[L.o2] + (size_t)img * p.out_istride[L.o2] + (size_t)(act ? pix : 0) * p.out_pix[L.o2];
              }
              for (int c16 = g; c16 < (ncols >> 4); c16 += kTG) {
                uint32_t u[16];
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(dcol + 16u * (uint32_t)c16));
                float4 bv[4], rv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int cq = (n0 >> 2) + 4 * c16 + j;              // channel quad of the layer
                  const uint32_t qo = 16u * (uint32_t)cq;
                  bv[j] = lds4(bias_a + qo);
                  rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                  if (L.res == 1) {
                    if (cq < res_q) rv[j] = lds4(res_a + qo);
                  } else if (L.res == 2) {
                    if (cq < res_q) rv[j] = max4(max4(lds4(res_a + qo), lds4(res_a + ksr_b + qo)), max4(lds4(res_a + row_b + qo), lds4(res_a + row_b + ksr_b + qo)));
                  } else if (L.res == 3) {
                    if (act && cq < res_q) rv[j] = __ldg(reinterpret_cast<const float4*>(gres + 4 * cq));
                  } else if (L.res == 4) {
                    if (act && cq < res_q)
                      rv[j] = max4(max4(__ldg(reinterpret_cast<const float4*>(gres + 4 * cq)), __ldg(reinterpret_cast<const float4*>(gres + gcs + 4 * cq))),
                                   max4(__ldg(reinterpret_cast<const float4*>(gres + grow + 4 * cq)), __ldg(reinterpret_cast<const float4*>(gres + grow + gcs + 4 * cq))));
                  }
                }
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int cq = (n0 >> 2) + 4 * c16 + j;
                  float4 v = make_float4(fmaf(__uint_as_float(u[4 * j]), L.wscale, bv[j].x), fmaf(__uint_as_float(u[4 * j + 1]), L.wscale, bv[j].y),
                                         fmaf(__uint_as_float(u[4 * j + 2]), L.wscale, bv[j].z), fmaf(__uint_as_float(u[4 * j + 3]), L.wscale, bv[j].w));
                  if (L.kind != 1) {
                    v.x += rv[j].x; v.y += rv[j].y; v.z += rv[j].z; v.w += rv[j].w;
                    if (L.act == 1) {
                      v = max4(v, make_float4(0.f, 0.f, 0.f, 0.f));
                    } else if (L.act == 2) {
                      const float4 al = lds4(alpha_a + 16u * (uint32_t)cq);
                      v = make_float4(fmaf(al.x, fminf(v.x, 0.f), fmaxf(v.x, 0.f)), fmaf(al.y, fminf(v.y, 0.f), fmaxf(v.y, 0.f)),
                                      fmaf(al.z, fminf(v.z, 0.f), fmaxf(v.z, 0.f)), fmaf(al.w, fminf(v.w, 0.f), fmaxf(v.w, 0.f)));
                    }
                    if (act && cq < dst_q) sts4(out_a + 16u * (uint32_t)cq, v);
                    if (act && cq < o1_q) *reinterpret_cast<float4*>(o1 + 4 * cq) = v;
                  } else if (act) {
                    const int c = 4 * cq;
                    if (c + 4 <= L.c1) {
                      *reinterpret_cast<float4*>(o1 + c) = v;
                    } else {
                      const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                      for (int e = 0; e < 4; ++e)
                        if (c + e >= L.c1 && c + e - L.c1 < L.c2) o2[c + e - L.c1] = vv[e];
                    }
                  }
                }
              }
            }
          }
          ph_d ^= 1u;
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");       // the accumulator reads precede the next group's MMAs (ordered by a_full)
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kTComputeThreads) : "memory");
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (l == p.last_a_layer && lane == 0) mbar_arrive(a_free);
      }
    }
  } else if (warp == kTC) {
    if (lane == 0) {
      // =============================== control lane: weight-block ring, depthwise records, MMA issue ========================
      const int total = my_images * nl;
      const int D = p.wdepth;
      const int total_blks = my_images * p.nblks;
      int next_load = 0, next_use = 0;
      auto top_up = [&]() {
        while (next_load < total_blks && next_load < next_use + D) {
          const TailBlk bk = sBlk[next_load % p.nblks];
          const uint32_t slot = (uint32_t)(next_load % D);
          if (next_load >= D) mbar_wait(w_free + 8u * slot, (uint32_t)((next_load / D - 1) & 1));
          mbar_expect_tx(w_full + 8u * slot, (uint32_t)bk.bytes);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(wring_a + slot * (uint32_t)p.wbuf_bytes), "l"(p.blob + (size_t)bk.off), "r"((uint32_t)bk.bytes), "r"(w_full + 8u * slot) : "memory");
          ++next_load;
        }
      };
      auto load_taps = [&](int glayer) {
        const TailLayerD& L = sL[glayer % nl];
        if (!L.tap_bytes) return;
        const uint32_t slot = (uint32_t)(glayer & 1), bar = t_full + 8u * slot;
        mbar_expect_tx(bar, (uint32_t)L.tap_bytes);
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(tring_a + slot * (uint32_t)p.tbuf_bytes), "l"(p.blob + (size_t)L.tap_off), "r"((uint32_t)L.tap_bytes), "r"(bar) : "memory");
      };
      if (my_images > 0) {
        top_up();
        for (int i = 0; i < 2 && i < total; ++i) load_taps(i);
      }
      uint32_t ph_a = 0u, ph_d = 0u;
      int gl = 0;
      for (int img = blockIdx.x; img < B; img += gridDim.x) {
        for (int l = 0; l < nl; ++l, ++gl) {
          const TailLayerD& L = sL[l];
          if (L.kind != 3) {
            const bool paired = L.OH * L.OW > 128;
            const int ntiles = paired ? 2 : 1;
            const uint32_t abase = (!paired && L.Npad > 128) ? 384u : 256u;
            const int ng = L.ng & 0xff, nbg = L.ng >> 8;
            for (int grp = 0; grp < ng; ++grp) {
              for (int kc = 0; kc < L.nk; ++kc) {
                top_up();
                const int kw = min(128, L.K16 - 128 * kc);
                const int ksteps = kw >> 4, khalf = kw >> 5;
                const uint32_t sbo = (uint32_t)(kw >> 3) * 128u;
                const uint32_t b_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
                for (int half = 0; half < 2; ++half) {
                  mbar_wait(a_full + 8u * (uint32_t)half, ph_a);
                  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                  const int k0 = half ? khalf : 0, k1 = half ? ksteps : khalf;
                  for (int nb = 0; nb < nbg; ++nb) {
                    const int gb = next_use + nb;
                    const uint32_t slot = (uint32_t)(gb % D);
                    if (half == 0) mbar_wait(w_full + 8u * slot, (uint32_t)((gb / D) & 1));
                    const TailBlk bk = sBlk[gb % p.nblks];
                    const uint32_t idesc = (1u << 4) | ((uint32_t)(bk.ncols >> 3) << 17) | ((128u >> 4) << 24);
                    const uint32_t wb_a = wring_a + slot * (uint32_t)p.wbuf_bytes;
                    const uint32_t b_lo0 = ((wb_a & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
                    const uint32_t b_lo1 = (((wb_a + (uint32_t)(bk.ncols * kw) * 2u) & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
                    for (int t = 0; t < ntiles; ++t) {
                      const uint32_t dcol = tmem_base + (paired ? (uint32_t)t * 128u : (uint32_t)bk.n0);
                      const uint32_t acol = tmem_base + abase + (uint32_t)t * 128u;
#pragma unroll 1
                      for (int ks = k0; ks < k1; ++ks) {
                        mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, (kc || ks) ? 1u : 0u);
                        mma_ts_f16(dcol, acol + 64u + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
                        if (L.w_parts == 2) mma_ts_f16(dcol, acol + 8u * (uint32_t)ks, b_lo1 + 16u * (uint32_t)ks, b_hi, idesc, 1u);
                      }
                    }
                  }
                }
                ph_a ^= 1u;
                for (int nb = 0; nb < nbg; ++nb)
                  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(w_free + 8u * (uint32_t)((next_use + nb) % D)) : "memory");
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(a_done) : "memory");
                next_use += nbg;
              }
              asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(d_full) : "memory");
              mbar_wait(d_full, ph_d);
              ph_d ^= 1u;
            }
          } else {
            mbar_wait(a_full, ph_a);
            mbar_wait(a_full + 8u, ph_a);
            ph_a ^= 1u;
          }
          if (gl + 2 < total) load_taps(gl + 2);
        }
      }
    }
    __syncwarp();
  } else {
    // =============================== loader warp: the next image's input map -> buffer 0, one bulk copy per pixel ============
    const uint32_t row_bytes = (uint32_t)p.CinS * 4u, ks0_b = 4u * (uint32_t)p.buf_ks[0];
    int it = 0;
    for (int img = blockIdx.x; img < B; img += gridDim.x, ++it) {
      if (it > 0) mbar_wait(a_free, (uint32_t)((it - 1) & 1));
      if (lane == 0) mbar_expect_tx(in_full, (uint32_t)p.in_px * row_bytes);
      __syncwarp();
      const float* src = p.in + (size_t)img * p.in_istride;
      for (int px = lane; px < p.in_px; px += 32)
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(act_a + (uint32_t)px * ks0_b), "l"(src + (size_t)px * p.CinS), "r"(row_bytes), "r"(in_full) : "memory");
    }
  }
  __syncwarp();
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kTC) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}

// f32 [cap][H][W][CinS], box {buf_ks[0], W, H, 1}: channels >= CinS are zero-filled (the residual's zero channel pad)
bool tail_tensor_map(const TailP& p, int cap, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, long long> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  Key key(p.in, cap, p.H, p.W, p.CinS, p.buf_ks[0], p.in_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.CinS, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.CinS * 4, (cuuint64_t)p.W * p.CinS * 4, (cuuint64_t)p.in_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.buf_ks[0], (cuuint32_t)p.W, (cuuint32_t)p.H, 1u};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUtensorMap tm;
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(p.in), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return false;
  if (cache.size() > 256) cache.clear();
  cache[key] = tm;
  *out = tm;
  return true;
}

}  // namespace

size_t tail_smem_bytes(int act_floats, int wbuf_bytes, int wdepth, int tbuf_bytes) {
  size_t head = (size_t)kTailMaxLayers * sizeof(TailLayerD) + 2 * (size_t)kTailMaxLayers * 128 * 4 + 16 * 8 + 16 * 4 + 128;
  size_t bufs = ((size_t)act_floats + 128) * 4 + 128;
  return head + bufs + 2 * (size_t)tbuf_bytes + (size_t)wdepth * wbuf_bytes;
}

size_t wide_smem_bytes(int act_floats, int wbuf_bytes, int wdepth, int tbuf_bytes, int bias_floats) {
  size_t head = (size_t)kTailMaxLayers * sizeof(TailLayerD) + (size_t)kTailMaxBlks * sizeof(TailBlk) + 2 * (size_t)bias_floats * 4 + 32 * 8 + 16 * 4 + 128;
  size_t bufs = ((size_t)act_floats + kWideZero) * 4 + 128;
  return head + bufs + 2 * (size_t)tbuf_bytes + (size_t)wdepth * wbuf_bytes;
}

static bool launch_chain_wide(const TailP& p, int B, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      if (cudaFuncSetAttribute(k_chain_wide, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess) return false;
      c = p.smem_bytes;
    }
  }
  if (p.wdepth < 1 || p.wdepth > 4 || p.nblks > kTailMaxBlks || p.CinS % 4 != 0 || p.in_istride % 4 != 0) return false;
  k_chain_wide<<<std::min(B, sms), kWThreads, p.smem_bytes, s>>>(p, B);
  return true;
}

bool launch_tail_ws(const TailP& p, int B, int cap, cudaStream_t s) {
  if (B <= 0) return true;
  if (p.generic == 2) return launch_chain_wide(p, B, s);
  CUtensorMap tm;
  if (!tail_tensor_map(p, cap, &tm)) return false;
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      if (cudaFuncSetAttribute(k_tail_ws<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess ||
          cudaFuncSetAttribute(k_tail_ws<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess) return false;
      c = p.smem_bytes;
    }
  }
  const int grid = std::min(B, sms);
  if (p.generic) k_tail_ws<true><<<grid, kTThreads, p.smem_bytes, s>>>(tm, p, B);
  else k_tail_ws<false><<<grid, kTThreads, p.smem_bytes, s>>>(tm, p, B);
  return true;
}

}  // namespace fdt
