// k_block_ts — BlazeBlock (depthwise 3x3 -> pointwise 1x1 -> + residual -> ReLU) for the high-resolution layers of the
// fp16-weight detectors, with the pointwise GEMM's A operand written straight into TENSOR MEMORY.
//
// Why a second block kernel: ncu shows k_block_ws limited by the L1 / shared-memory pipe (l1tex 69-81 %, DRAM 32 %).
// Per output element it moves through shared memory: the TMA fill, the depthwise taps, the TF32 hi / lo operand STORES,
// the tensor core's operand READS, the residual.  Here the operand never touches shared memory (tcgen05.st -> TMEM,
// tcgen05.mma reads A from TMEM), it is split into two fp16 terms instead of two TF32 terms (the detectors' weights are
// fp16-origin, i.e. exact: same ~22-bit products at HALF the MMA count), and the two 128-pixel M-tiles of a 16x16 output
// tile are the tile's even and odd rows, so that every thread (= one TMEM lane of both M-tiles) owns two vertically
// adjacent pixels and feeds both from ONE 4 x 3 window of loads (6 instead of 9 LDS.128 per output quad).
//
//   warps 0..11  compute : thread = (TMEM lane = pixel column x / row pair yy, one third of the channel quads);
//                          per tile: depthwise -> fp16 hi / lo -> TMEM operand of slot i & 1 -> arrive a_full;
//                          then the epilogue of the PREVIOUS tile (its MMAs ran meanwhile): tcgen05.ld -> + bias
//                          + residual from the still-staged input tile -> ReLU -> 256-bit stores to HBM.
//   warp 12 lane 0 producer: TMA of input tile + halo into a ring of `ns` stages (out-of-bounds zero fill = SAME padding
//                          and channel pad).
//   warp 13 lane 0 MMA   : tcgen05.mma.kind::f16, A from TMEM, W from shared memory; two passes (hi, lo); one commit per tile.
// No __syncthreads and no named barrier in the steady state: tiles are independent, all hand-offs are mbarriers.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>
#include <type_traits>

#include "kernels.h"
#include "tile_walk.h"

namespace fdt {
namespace {

constexpr int kC = 12;                        // compute warps
constexpr int kThreads = (kC + 2) * 32;
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
// The single-lane roles (TMA producer, MMA issuer) wait with a suspend-time hint: with the bare try_wait loop the two lanes re-issued
// YIELD / TRYWAIT / BRA every ~16 cycles - ncu counted 7.1 M spin iterations per 1024 frames on short-range block 1, 13.5 % of ALL
// issued instructions, on the two schedulers that also host compute warps.
__device__ __forceinline__ void mbar_wait_idle(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok)
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                 : "=r"(ok) : "r"(bar), "r"(parity), "r"(0x989680u) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
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
__device__ __forceinline__ uint32_t pack_h2(float e0, float e1) {
  uint32_t d;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(e1), "f"(e0));
  return d;
}
// a = hi + lo (fp16 each) -> two 2-column TMEM stores; lo_off = column distance of the lo half
__device__ __forceinline__ void split_store_tmem(uint32_t taddr_hi, uint32_t lo_off, const float4& a) {
  const uint32_t h0 = pack_h2(a.x, a.y), h1 = pack_h2(a.z, a.w);
  const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&h0)), f1 = __half22float2(*reinterpret_cast<const __half2*>(&h1));
  const uint32_t l0 = pack_h2(a.x - f0.x, a.y - f0.y), l1 = pack_h2(a.z - f1.x, a.w - f1.y);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi), "r"(h0), "r"(h1) : "memory");
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr_hi + lo_off), "r"(l0), "r"(l1) : "memory");
}
__device__ __forceinline__ void mma_ts_f16(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(tmem_d), "r"(tmem_a), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc));
}
__device__ __forceinline__ void stg8(float* p, const float4& a, const float4& b) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
}

// TMEM columns: accumulators at (slot * 2 + t) * 64, operands at 256 + (slot * 2 + t) * 64 (hi halves +0, lo halves +32):
// K16 <= 64 and Npad <= 64 (plan_ts), two tile slots in flight.
__device__ __forceinline__ uint32_t col_d(int slot, int t) { return (uint32_t)(slot * 2 + t) * 64u; }
__device__ __forceinline__ uint32_t col_a(int slot, int t) { return 256u + (uint32_t)(slot * 2 + t) * 64u; }

// Shared-memory carve-up (must match ts_smem_bytes): [W fp16 Npad x K16 | dww 9 x K16 | dwb K16] [bias Npad] [barriers 32 x 8 B]
//   | 128-byte aligned: [input ring ns x stage_bytes]
// S = depthwise stride.  S == 1: output tile 16 x 16 (M-tiles = even / odd rows), staged 18 x 18.  S == 2: output tile 8 x 16
// (one M-tile), staged 17 rows x 33 columns as TWO planes of 17 x 17 pixels - the even and the odd input columns, each fetched by
// its own TMA with an element stride of 2 along x.  Neighbouring lanes (output columns x, x + 1) then read pixels that are ONE
// record apart (KS = an odd number of 16-byte quads: conflict-free LDS.128) instead of two records apart (every load 2-way
// bank-conflicted: ncu counted 12.7 M conflict cycles per 1024 frames on short-range block 3).
// KSC = staged pixel stride in floats as a COMPILE-TIME value (28 / 36 / 44 / 52: the 24..48-channel layers; 0 = read p.KS): the twelve
// window loads of a quad then carry immediate offsets instead of ~20 address multiply-adds per quad.
template <int S, int KSC>
__global__ void __launch_bounds__(kThreads, 1) k_block_ts(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ BlockTsP p, int B) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ uint32_t tmem_base_s;
  constexpr int NT = S == 1 ? 2 : 1;                         // M-tiles per output tile
  constexpr int TH = S == 1 ? 16 : 8, TW = 16;
  constexpr int IW = S == 1 ? 18 : 17;                       // staged row length (pixels; S == 2: of one column-parity plane)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* sRec = reinterpret_cast<float*>(smem_raw);
  float* sBias = sRec + (p.rec_bytes >> 2);
  uint64_t* bars = reinterpret_cast<uint64_t*>(sBias + p.Npad);
  unsigned char* ring = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(bars + 32) + 127) & ~(uintptr_t)127);
  const uint32_t bar0 = smem_u32(bars);
  const uint32_t in_full = bar0, in_empty = bar0 + 64u, a_full = bar0 + 128u, d_full = bar0 + 144u;   // [8] [8] [2] [2]
  const int NS = p.ns;
  const int tiles_x = p.OW / TW, tiles_y = p.OH / TH, tpi = tiles_x * tiles_y;
  const int ntiles = B * tpi;
  const int n_my = (int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  // ---- prologue ---------------------------------------------------------------------------------------------------
  for (int i = tid; i < (p.rec_bytes >> 4); i += kThreads) reinterpret_cast<uint4*>(sRec)[i] = reinterpret_cast<const uint4*>(p.rec)[i];
  for (int i = tid; i < p.Npad; i += kThreads) sBias[i] = p.bias[i];
  if (tid < 2) bars[24 + tid] = 0ull;                           // 16 zero bytes: what a column quad without residual reads
  if (tid == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(in_full + 8u * i, 1); mbar_init(in_empty + 8u * i, kC); }
    for (int i = 0; i < 2; ++i) { mbar_init(a_full + 8u * i, kC); mbar_init(d_full + 8u * i, 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap) : "memory");
  }
  if (warp == kC + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // W was written by the generic proxy, the MMA reads it through the async proxy
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  const int KS = KSC ? KSC : p.KS;
  const uint32_t ring_a = smem_u32(ring), ks_b = (uint32_t)KS * 4u, row_b = (uint32_t)IW * ks_b;
  const uint32_t stage_b = (uint32_t)p.stage_bytes;
  const uint32_t plane_b = (17u * 17u * ks_b + 127u) & ~127u;   // S == 2: the odd-column plane follows the even-column plane (TMA destinations are 128-byte aligned)
  // S == 2: byte offset of window column kx = 0, 1, 2 (input columns 2x, 2x + 1, 2x + 2)
  const uint32_t col2_b[3] = {0u, plane_b, ks_b};

  if (warp < kC) {
    // =============================== compute warps ==============================================================
    const int lq = warp & 3, g = warp >> 2;
    const int L = lq * 32 + lane;                            // TMEM lane
    const int yy = L >> 4, x = L & 15;                       // S == 1: row pair / column of the 16 x 16 tile; S == 2: row / column of the 8 x 16 tile
    const uint32_t tm_lane = tmem_base + ((uint32_t)(lq * 32) << 16);
    // top-left of this thread's input window inside a stage: S == 1 rows 2yy .. 2yy + 3, cols x .. x + 2 (halo included);
    // S == 2 rows 2yy .. 2yy + 2, cols 2x .. 2x + 2 = entries x (even plane), x (odd plane), x + 1 (even plane)
    const uint32_t win = (uint32_t)((2 * yy) * IW + x) * ks_b;
    const int nq = p.K16 >> 2, nq_real = (p.Cin + 3) >> 2, ks_q = KS >> 2;
    const uint32_t bias_a = smem_u32(sBias), zero_a = bar0 + 192u;
    const int couts = p.CoutS, nc8 = (couts + 7) >> 3;             // 8-column groups that hold real channels
    const int gw = (nc8 + 2) / 3, g0 = gw * g, gn = max(0, min(gw, nc8 - g0));   // this warp's share of them: groups g0 .. g0 + gn - 1 (gn <= 3)

    // S == 2: the pooled residual of the tile whose epilogue is still to come, pulled out of the input stage right after the tile's
    // depthwise so that the stage goes back to the producer one iteration earlier (with two stages the next tile's TMA otherwise
    // starts only when the previous epilogue has finished: its whole latency was exposed, ~2.5 of 5.9 us per tile on block 6)
    float4 rvp[3][2];
    auto epilogue = [&](int i, int stage, int b, int ty, int tx) {
      const int slot = i & 1;
      mbar_wait(d_full + 8u * (uint32_t)slot, (uint32_t)((i >> 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t st_a = ring_a + (uint32_t)stage * stage_b + win;
      // the thread's first output pixel (one 64-bit multiply per tile; an image is far below 2^31 floats); t = 1 is the row below
      float* const orow0 = p.out + (size_t)b * p.out_istride + (uint32_t)(((ty * TH + (S == 1 ? 2 * yy : yy)) * p.OW + tx * TW + x) * couts);
      const uint32_t orow_step = (uint32_t)(p.OW * couts);
#pragma unroll
      for (int t = 0; t < NT; ++t) {
        float* orow = orow0 + (t ? orow_step : 0u);
        // residual: S == 1 the centre of the window of output row t; S == 2 the 2x2 max-pool at the window's top-left
        const uint32_t res_a = S == 1 ? st_a + (uint32_t)(t + 1) * row_b + ks_b : st_a;
        const uint32_t dcol = tm_lane + col_d(slot, t);
        // The warp's share of the 8-column groups is contiguous: gw = ceil(nc8 / 3) groups from gw * g.  Two groups go through ONE
        // 16-column TMEM round trip (with the deal g, g + 3, ... every group cost its own tcgen05.ld / wait, and four groups - 25..32
        // output channels - gave one warp of three double work); lg = index of the group inside the warp's share.
        auto unit = [&](auto ngc, auto lgc, int c8) {
          constexpr int NG = decltype(ngc)::value;             // groups in this unit: 1 (8 columns) or 2 (16 columns)
          constexpr int lg = decltype(lgc)::value;             // (compile-time: rvp must stay in registers)
          uint32_t u[8 * NG];
          if (NG == 2)
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                  "=r"(u[8 * (NG - 1)]), "=r"(u[8 * (NG - 1) + 1]), "=r"(u[8 * (NG - 1) + 2]), "=r"(u[8 * (NG - 1) + 3]),
                  "=r"(u[8 * (NG - 1) + 4]), "=r"(u[8 * (NG - 1) + 5]), "=r"(u[8 * (NG - 1) + 6]), "=r"(u[8 * (NG - 1) + 7])
                : "r"(dcol + 8u * (uint32_t)c8));
          else
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                         : "r"(dcol + 8u * (uint32_t)c8));
          float4 bv[2 * NG], rv[2 * NG];
#pragma unroll
          for (int j = 0; j < 2 * NG; ++j) {
            const int cq = 2 * c8 + j;
            const uint32_t qo = 16u * (uint32_t)cq;
            bv[j] = lds4(bias_a + qo);
            rv[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (S == 2) {
              rv[j] = rvp[lg + (j >> 1)][j & 1];
            } else if (S == 1) {                    // (stride 1: residual = the window's centre pixel or none; no branch: a quad without one reads zeros)
              rv[j] = lds4(p.res == 1 && cq < ks_q ? res_a + qo : zero_a);
            } else if (p.res == 2) {
              if (cq < ks_q) rv[j] = max4(max4(lds4(res_a + qo), lds4(res_a + plane_b + qo)), max4(lds4(res_a + row_b + qo), lds4(res_a + row_b + plane_b + qo)));
            }
          }
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          float4 v[2 * NG];
#pragma unroll
          for (int j = 0; j < 2 * NG; ++j) {
            v[j] = make_float4(__uint_as_float(u[4 * j]) + bv[j].x + rv[j].x, __uint_as_float(u[4 * j + 1]) + bv[j].y + rv[j].y,
                               __uint_as_float(u[4 * j + 2]) + bv[j].z + rv[j].z, __uint_as_float(u[4 * j + 3]) + bv[j].w + rv[j].w);
            if (p.relu) v[j] = max4(v[j], make_float4(0.f, 0.f, 0.f, 0.f));
          }
#pragma unroll
          for (int k = 0; k < NG; ++k) {
            const int c = 8 * (c8 + k);                          // columns >= Cout inside CoutS are exact zeros
            if (c + 8 <= couts) {
              if (p.wide) stg8(orow + c, v[2 * k], v[2 * k + 1]);
              else { *reinterpret_cast<float4*>(orow + c) = v[2 * k]; *reinterpret_cast<float4*>(orow + c + 4) = v[2 * k + 1]; }
            } else if (c + 4 <= couts) {
              *reinterpret_cast<float4*>(orow + c) = v[2 * k];
            }
          }
        };
        if (gn >= 2) unit(std::integral_constant<int, 2>(), std::integral_constant<int, 0>(), g0);
        if (gn == 1) unit(std::integral_constant<int, 1>(), std::integral_constant<int, 0>(), g0);
        if (gn == 3) unit(std::integral_constant<int, 1>(), std::integral_constant<int, 2>(), g0 + 2);
      }
      // this tile's accumulators and (S == 1) input stage are consumed
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (S == 1 && lane == 0) mbar_arrive(in_empty + 8u * (uint32_t)stage);
    };

    // The tile walk is INCREMENTAL: tile i of this CTA is blockIdx.x + i * gridDim.x; its (image, tile row, tile column) and its ring
    // stage advance by constant steps with a carry.  (With tile / tpi, trem / tiles_x and i % NS spelled as divisions every compute
    // warp spent ~170 of its ~680 instructions per tile on integer division - ncu source counters, short-range block 1.)
    const TileStep step = tile_step((int)gridDim.x, tpi, tiles_x);
    TileAt at = tile_at((int)blockIdx.x, tpi, tiles_x), at_prev = at;
    int stage = 0, stage_prev = 0;
    uint32_t in_phase = 0;
    for (int i = 0; i < n_my; ++i) {
      const int slot = i & 1;
      mbar_wait(in_full + 8u * (uint32_t)stage, in_phase);
      const uint32_t st_a = ring_a + (uint32_t)stage * stage_b + win;
      const uint32_t acol0 = tm_lane + col_a(slot, 0), acol1 = tm_lane + col_a(slot, 1);
      for (int q = g; q < nq; q += 3) {
        // a pad quad (channels beyond Cin, up to K16) is all zeros: written by the first tile of each operand slot, then left alone
        if (q >= nq_real && i >= 2) break;
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0;
        if (q < nq_real) {
          const uint32_t qo = 16u * (uint32_t)q;
          // taps and bias of the quad from the kernel parameters (constant bank: the address is the same for every lane)
          const float4* cw = reinterpret_cast<const float4*>(p.dw) + 10 * q;      // quad-major: the ten float4 of a quad are contiguous
          float4 w[9];
#pragma unroll
          for (int k = 0; k < 9; ++k) w[k] = cw[k];
          const float4 bias = cw[9];
          const uint32_t pa = st_a + qo;
          if (S == 1) {
            // rows 0..3 of the window feed output rows 0 (rows 0-2) and 1 (rows 1-3); each output: bias, then taps in (ky, kx) order
            float4 v[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = lds4(pa + (uint32_t)c * ks_b);
            a0 = bias;
#pragma unroll
            for (int c = 0; c < 3; ++c) fma4(a0, v[c], w[c]);
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = lds4(pa + row_b + (uint32_t)c * ks_b);
            a1 = bias;
#pragma unroll
            for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[3 + c]); fma4(a1, v[c], w[c]); }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = lds4(pa + 2u * row_b + (uint32_t)c * ks_b);
#pragma unroll
            for (int c = 0; c < 3; ++c) { fma4(a0, v[c], w[6 + c]); fma4(a1, v[c], w[3 + c]); }
#pragma unroll
            for (int c = 0; c < 3; ++c) v[c] = lds4(pa + 3u * row_b + (uint32_t)c * ks_b);
#pragma unroll
            for (int c = 0; c < 3; ++c) fma4(a1, v[c], w[6 + c]);
          } else {
            a0 = bias;
#pragma unroll
            for (int k = 0; k < 9; ++k) fma4(a0, lds4(pa + (uint32_t)(k / 3) * row_b + col2_b[k % 3]), w[k]);
          }
        }
        split_store_tmem(acol0 + 2u * (uint32_t)q, 32u, a0);
        if (S == 1) split_store_tmem(acol1 + 2u * (uint32_t)q, 32u, a1);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full + 8u * (uint32_t)slot);
      if (i > 0) epilogue(i - 1, stage_prev, at_prev.b, at_prev.ty, at_prev.tx);      // the previous tile's MMAs ran during this tile's depthwise
      if (S == 2) {
        // residual of THIS tile (2x2 max-pool at the window's top-left; zero when the block has none) -> registers; stage released
#pragma unroll
        for (int c8i = 0; c8i < 3; ++c8i)
#pragma unroll
          for (int j = 0; j < 2; ++j) {
            const int cq = 2 * (g0 + c8i) + j;
            const uint32_t qo = 16u * (uint32_t)cq;
            rvp[c8i][j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.res == 2 && c8i < gn && cq < ks_q)
              rvp[c8i][j] = max4(max4(lds4(st_a + qo), lds4(st_a + plane_b + qo)), max4(lds4(st_a + row_b + qo), lds4(st_a + row_b + plane_b + qo)));
          }
        __syncwarp();
        if (lane == 0) mbar_arrive(in_empty + 8u * (uint32_t)stage);
      }
      at_prev = at; stage_prev = stage;
      tile_advance(at, step, tiles_x, tiles_y);
      if (++stage == NS) { stage = 0; in_phase ^= 1u; }
    }
    if (n_my > 0) epilogue(n_my - 1, stage_prev, at_prev.b, at_prev.ty, at_prev.tx);
  } else if (warp == kC) {
    // =============================== TMA producer ===============================================================
    if (lane == 0) {
      const TileStep step = tile_step((int)gridDim.x, tpi, tiles_x);
      TileAt at = tile_at((int)blockIdx.x, tpi, tiles_x);
      int stage = 0;
      uint32_t empty_phase = 1;                                  // parity of the in_empty phase that frees the stage: ((i / NS) - 1) & 1
      for (int i = 0; i < n_my; ++i) {
        const int b = at.b, ty = at.ty, tx = at.tx;
        if (i >= NS) mbar_wait_idle(in_empty + 8u * (uint32_t)stage, empty_phase);
        const uint32_t bar = in_full + 8u * (uint32_t)stage;
        mbar_expect_tx(bar, (uint32_t)((S == 1 ? 18 * 18 : 2 * 17 * 17) * KS) * 4u);
        const int ix0 = S == 1 ? tx * TW - 1 : tx * TW * 2, iy0 = S == 1 ? ty * TH - 1 : ty * TH * 2;
        asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                     ::"r"(ring_a + (uint32_t)stage * stage_b), "l"(&tmap), "r"(0), "r"(ix0), "r"(iy0), "r"(b), "r"(bar) : "memory");
        if (S == 2)        // the odd input columns (the map walks x with an element stride of 2)
          asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                       ::"r"(ring_a + (uint32_t)stage * stage_b + plane_b), "l"(&tmap), "r"(0), "r"(ix0 + 1), "r"(iy0), "r"(b), "r"(bar) : "memory");
        tile_advance(at, step, tiles_x, tiles_y);
        if (++stage == NS) { stage = 0; empty_phase ^= 1u; }
      }
    }
    __syncwarp();
  } else {
    // =============================== MMA issuer =================================================================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(p.Npad >> 3) << 17) | ((128u >> 4) << 24);       // D f32, A / B f16, K-major
      const uint32_t sbo = (uint32_t)(p.K16 >> 3) * 128u;
      const uint32_t b_hi = ((sbo >> 4) & 0x3FFFu) | (1u << 14);
      const uint32_t b_lo0 = ((smem_u32(sRec) & 0x3FFFFu) >> 4) | ((kLBO >> 4) << 16);
      const int ksteps = p.K16 >> 4;
      for (int i = 0; i < n_my; ++i) {
        const int slot = i & 1;
        mbar_wait_idle(a_full + 8u * (uint32_t)slot, (uint32_t)((i >> 1) & 1));
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
        for (int t = 0; t < NT; ++t) {
          const uint32_t dcol = tmem_base + col_d(slot, t), acol = tmem_base + col_a(slot, t);
          uint32_t acc = 0u;
#pragma unroll 1
          for (int part = 0; part < 2; ++part)
#pragma unroll 1
            for (int ks = 0; ks < ksteps; ++ks) {
              mma_ts_f16(dcol, acol + 32u * (uint32_t)part + 8u * (uint32_t)ks, b_lo0 + 16u * (uint32_t)ks, b_hi, idesc, acc);
              acc = 1u;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(d_full + 8u * (uint32_t)slot) : "memory");
      }
    }
    __syncwarp();
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == kC + 1) {
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

bool ts_tensor_map(const BlockTsP& p, int cap, CUtensorMap* out) {
  typedef std::tuple<const void*, int, int, int, int, int, int, long long> Key;
  static std::mutex mu;
  static std::map<Key, CUtensorMap> cache;
  Key key(p.in, cap, p.H, p.W, p.CinS, p.KS, p.stride, p.in_istride);
  std::lock_guard<std::mutex> g(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return true; }
  EncodeTiledFn enc = encode_fn();
  if (!enc) return false;
  cuuint64_t gdim[4] = {(cuuint64_t)p.CinS, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)cap};
  cuuint64_t gstr[3] = {(cuuint64_t)p.CinS * 4, (cuuint64_t)p.W * p.CinS * 4, (cuuint64_t)p.in_istride * 4};
  cuuint32_t box[4] = {(cuuint32_t)p.KS, p.stride == 1 ? 18u : 33u, p.stride == 1 ? 18u : 17u, 1u};
  cuuint32_t estr[4] = {1, p.stride == 1 ? 1u : 2u, 1, 1};     // stride 2: every other column (33 traversed -> 17 delivered)
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

template <int S, int KSC>
bool launch_ts(const CUtensorMap& tm, const BlockTsP& p, int B, cudaStream_t s) {
  static std::mutex mu;
  static std::map<int, size_t> cur;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  {
    std::lock_guard<std::mutex> g(mu);
    size_t& c = cur[dev];
    if (p.smem_bytes > c) {
      if (cudaFuncSetAttribute(k_block_ts<S, KSC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem_bytes) != cudaSuccess) return false;
      c = p.smem_bytes;
    }
  }
  const int ntiles = B * (p.OW / 16) * (p.OH / (S == 1 ? 16 : 8));
  const int grid = std::max(1, std::min(ntiles, sms));
  k_block_ts<S, KSC><<<grid, kThreads, p.smem_bytes, s>>>(tm, p, B);
  return true;
}

}  // namespace

size_t ts_smem_bytes(int rec_bytes, int Npad, int ns, int stage_bytes) {
  return (size_t)rec_bytes + (size_t)Npad * 4 + 32 * 8 + 128 + (size_t)ns * stage_bytes + 128;
}

bool launch_block_ts(const BlockTsP& p, int B, int cap, cudaStream_t s) {
  if (B <= 0) return true;
  CUtensorMap tm;
  if (!ts_tensor_map(p, cap, &tm)) return false;
  if (p.stride == 1) {
    switch (p.KS) {
      case 28: return launch_ts<1, 28>(tm, p, B, s);
      case 36: return launch_ts<1, 36>(tm, p, B, s);
      case 44: return launch_ts<1, 44>(tm, p, B, s);
      case 52: return launch_ts<1, 52>(tm, p, B, s);
      default: return launch_ts<1, 0>(tm, p, B, s);
    }
  }
  switch (p.KS) {
    case 28: return launch_ts<2, 28>(tm, p, B, s);
    case 36: return launch_ts<2, 36>(tm, p, B, s);
    case 44: return launch_ts<2, 44>(tm, p, B, s);
    case 52: return launch_ts<2, 52>(tm, p, B, s);
    default: return launch_ts<2, 0>(tm, p, B, s);
  }
}

}  // namespace fdt
