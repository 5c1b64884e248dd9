// Warp-specialised band kernel of the B200-native FovVideoVDP core (sm_100a) -- the video hot path (temporal windows of up
// to 8 taps, no debug outputs).  Same arithmetic and data layout as fused::band_kernel (fvvdp_fused.cuh: pyramid once per
// luminance frame, (test, reference) pairs, temporal rings on chip), restructured around the measured limiter of that
// kernel -- latency at 16 warps per SM behind three CTA-wide barriers per frame:
//
//   * ONE CTA per SM owns a 32x64 tile (tile halo 1.41x instead of 1.69x) and runs 28 warps (24 on the pyramid levels) in two roles,
//       producer warps (12 at level 0, 8 on the pyramid levels): wait for the TMA tile -> display EOTF -> luminance tile -> 5-tap reduce (rows, columns;
//                            fvvdp_lpyr_dec.py:183-207) -> next pyramid level out -> ring of reduced tiles -> temporal
//                            filters of the reduced tile (fvvdp.py:294-300),
//       consumer warps (16): register ring of the tile's own pixels (one 2x2 quad per thread) -> temporal filters ->
//                            expand (fvvdp_lpyr_dec.py:219-235) -> contrast, CSF, masking, sum D^beta
//                            (fvvdp_lpyr_dec.py:259-269, fvvdp.py:520-537, 574-607),
//     decoupled by mbarriers (full / empty per luminance buffer): no CTA-wide barrier in the frame loop, the producers run
//     up to two frames ahead, MUFU-heavy (EOTF, masking) and FMA-heavy (filters, stencils) warps are co-resident by
//     construction.  `setmaxnreg` moves registers from the producers (40 / 48) to the consumers (96).
//   * 7-position rings: the newest frame's sustained tap is 1e-36 (t = 0 in fvvdp.py:609-620) and the oldest frame's
//     transient tap is exactly 0 (:626), so a window of fl taps needs fl-1 stored frames; the newest frame enters the
//     transient filter from the registers it was just loaded into.
//   * a frame stays at ring position (index in the clip) mod 7 and the step's position is a compile-time constant of one
//     of seven code versions of the filter, so every filter weight is a uniform-register operand of its FFMA2 (no vector
//     registers, no loads) and the summation order does not depend on how the clip is cut into blocks or ranks.
#pragma once
#include "fvvdp_fused.cuh"
#include "fvvdp_ws_geometry.h"

#ifndef WS_TAPS16
#define WS_TAPS16 0
#endif
#if WS_TAPS16
#define WS_NS ws16
#else
#define WS_NS ws
#endif

namespace fvvdp {
namespace WS_NS {

using fused::BandParams;
using fused::u64;
using fused::pk; using fused::lo_of; using fused::hi_of; using fused::ffma2; using fused::fmul2; using fused::fadd2; using fused::tap5;
using fused::lds128; using fused::sts128; using fused::smem_u32; using fused::mbar_init; using fused::mbar_expect_tx; using fused::mbar_wait;
using fused::tma_load_3d; using fused::eotf8; using fused::eotf_checks_range; using fused::locate_direct;
using fused::IN_PYRAMID_TMA; using fused::IN_LEVEL0_TMA; using fused::locate_smem;

constexpr int PXT = TH / 8;                     // pixels per consumer thread: a 2x2 quad (32-row tiles) or one row of it (16-row tiles)
constexpr int NH = TH / 2 + 2, NW = TW / 2 + 2; // reduced tile with 1-px halo: origin (jy0-1, jx0-1)
constexpr int NE = NH * NW;                     // 612
#ifndef WS_BACKOFF
#define WS_BACKOFF 0    // ns a consumer warp sleeps between two looks at a barrier that is not ready (0: hardware-suspended try_wait only)
#endif
// Warp split and registers per thread after setmaxnreg (the CTA's pool is NT x (65536 / NT rounded down to 8)).  Each
// translation unit instantiates ONE input kind, so the split is chosen per kind:
//   level 0 (EOTF pass in the producers): 12 producer warps, 512 * 96 + 384 * 40 = 896 * 72   (the consumers are the bound)
//   pyramid levels (no EOTF pass):         8 producer warps, 512 * 96 + 256 * 48 = 768 * 80
#if defined(WS_KIND) && WS_KIND == 2 && defined(WS2_NPW)   // experiment builds: override for the pyramid-level unit only
#define WS_NPW WS2_NPW
#define WS_CREGS WS2_CREGS
#define WS_PREGS WS2_PREGS
#endif
#ifndef WS_NPW
#if defined(WS_KIND) && WS_KIND == 2
#define WS_NPW 8
#define WS_CREGS 96
#define WS_PREGS 48
#else
#define WS_NPW 12
#define WS_CREGS 96
#define WS_PREGS 40
#endif
#endif
constexpr int NCW = 16, NPW = WS_NPW;           // consumer / producer warps
constexpr int NCT = NCW * 32, NPT = NPW * 32, NT = NCT + NPT;
// registers per thread after setmaxnreg; the CTA's pool is NT x (65536 / NT rounded down to 8):
//   8 producer warps: 512 * 96 + 256 * 48 = 768 * 80;  12 producer warps: 512 * 96 + 384 * 40 = 896 * 72
constexpr int CONSUMER_REGS = WS_CREGS, PRODUCER_REGS = WS_PREGS;
constexpr int MAXCHUNK = fused::MAXCHUNK;
constexpr int LV4 = LW / 4;                     // 4-pixel chunks per staged row
constexpr int NPC = LH * LV4;                   // 720
constexpr int NLD = (NPC + NPT - 1) / NPT;      // 3
constexpr int NCOL = (NE + NPT - 1) / NPT;      // 3
constexpr int PLANE = LH * LW;                  // one stream of a landing buffer (floats)
constexpr int TILE_FLOATS = 2 * PLANE;          // one staged tile, both streams (23040 bytes)
constexpr int ROW_CP = LW / 2;                  // row pass: column pairs (36) x segments of ROW_SEG reduced rows
constexpr int ROW_SEG = TH == 32 ? (NPW >= 12 ? 2 : (NPW >= 8 ? 3 : 6)) : (NPW >= 12 ? 1 : 2);   // NH / ROW_SEG segments x 36 column pairs <= NPT threads
constexpr int ROW_THREADS = ROW_CP * (NH / ROW_SEG);  // 216 / 324

template <int KIND, bool FOV>
struct Layout {
  static constexpr bool LANDING = KIND == IN_LEVEL0_TMA;     // raw planes land first, the EOTF pass interleaves them
  // luminance / filtered-reduced-tile buffers (frames in flight between the roles).  The pyramid levels stage with TMA straight
  // into these buffers, AHEAD = NLB - 2 frames ahead: the buffer that is refilled was released two iterations ago, so the
  // thread that issues the copies never waits for the consumers
#ifndef WS_FOV_NLB
#define WS_FOV_NLB 3   // measured: a third buffer (+8 %) beats the larger L1 that two buffers would leave for the CSF records
#endif
  static constexpr int NLB = LANDING ? (FOV ? WS_FOV_NLB : 3) : (FOV ? 3 : 4);
  static constexpr int AHEAD = LANDING ? 2 : NLB - 2;
  static constexpr int oL = 0;                               // [NLB][LH][LW][2]
  static constexpr int oRaw = oL + NLB * TILE_FLOATS;        // LANDING: [2][2 streams][LH][LW]
  static constexpr int NV = 2;                               // row-reduced tiles: the column pass of frame i-1 overlaps the row pass of frame i
  static constexpr int oV = oRaw + (LANDING ? 2 * TILE_FLOATS : 0);  // [NV][NH][LW][2] row-reduced
  static constexpr int oNr = oV + NV * 2 * NH * LW;          // [RP][NE][2] ring of reduced tiles
  static constexpr int oNc = oNr + RP * 2 * NE;              // [NLB][2][NE][2] temporally filtered reduced tiles
  static constexpr int oTab = oNc + NLB * 4 * NE;            // [32][8]
  static constexpr int oFov = oTab + 256;                    // FOV: float2 [PXT][NCT] (rho fraction, rho cell), then the view
  static constexpr int oView = oFov + (FOV ? 2 * PXT * NCT : 0);  //   direction of the tile's columns [TW] and rows [TH] (deg)
  static constexpr int total = oView + (FOV ? TW + TH : 0);
  static constexpr size_t bytes = sizeof(float) * (size_t)total;
  static_assert(bytes + 1024 <= 227 * 1024, "shared memory of one CTA");
};

__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// wait with back-off: a warp that finds the barrier not ready yet sleeps instead of competing for issue slots
__device__ __forceinline__ void mbar_wait_backoff(unsigned bar, unsigned parity) {
#if WS_BACKOFF > 0
  unsigned done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  while (!done) {
    __nanosleep(WS_BACKOFF);
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  }
#else
  mbar_wait(bar, parity);
#endif
}
// switch over the ring position with one code version per position (J = compile-time position)
#if WS_TAPS16
#define FVVDP_RING_SWITCH(rp_, STMT)                                   \
  switch (rp_) {                                                       \
    case 0: { constexpr int J = 0; STMT; } break;                      \
    case 1: { constexpr int J = 1; STMT; } break;                      \
    case 2: { constexpr int J = 2; STMT; } break;                      \
    case 3: { constexpr int J = 3; STMT; } break;                      \
    case 4: { constexpr int J = 4; STMT; } break;                      \
    case 5: { constexpr int J = 5; STMT; } break;                      \
    case 6: { constexpr int J = 6; STMT; } break;                      \
    case 7: { constexpr int J = 7; STMT; } break;                      \
    case 8: { constexpr int J = 8; STMT; } break;                      \
    case 9: { constexpr int J = 9; STMT; } break;                      \
    case 10: { constexpr int J = 10; STMT; } break;                    \
    case 11: { constexpr int J = 11; STMT; } break;                    \
    case 12: { constexpr int J = 12; STMT; } break;                    \
    case 13: { constexpr int J = 13; STMT; } break;                    \
    default: { constexpr int J = 14; STMT; } break;                    \
  }
#else
#define FVVDP_RING_SWITCH(rp_, STMT)                                   \
  switch (rp_) {                                                       \
    case 0: { constexpr int J = 0; STMT; } break;                      \
    case 1: { constexpr int J = 1; STMT; } break;                      \
    case 2: { constexpr int J = 2; STMT; } break;                      \
    case 3: { constexpr int J = 3; STMT; } break;                      \
    case 4: { constexpr int J = 4; STMT; } break;                      \
    case 5: { constexpr int J = 5; STMT; } break;                      \
    default: { constexpr int J = 6; STMT; } break;                     \
  }
#endif
template <int ID, int COUNT>
__device__ __forceinline__ void named_bar_sync() { asm volatile("bar.sync %0, %1;" ::"n"(ID), "n"(COUNT) : "memory"); }

// One 4-pixel position chunk: landing buffer (two planes) -> EOTF -> luminance tile ((test, ref) interleaved).
// The range of the raw samples ("Pixel outside the valid range 0-1", fvvdp_display_model.py:149-151) is tracked with
// three-input min / max (FMNMX3): one instruction per sample.
template <int EOTF, bool INSIDE>
__device__ __forceinline__ void eotf_chunk(unsigned raw, unsigned lum, bool inside, const BandParams& p, float& vmin, float& vmax) {
  const float4 a = lds128(raw), b = lds128(raw + PLANE * 4);
  float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  if (eotf_checks_range(EOTF)) {
#ifdef WS_NO_RANGE
#else
#pragma unroll
    for (int j = 0; j < 8; j += 2) {
      vmin = fminf(fminf(vmin, x[j]), x[j + 1]);
      vmax = fmaxf(fmaxf(vmax, x[j]), x[j + 1]);
    }
#endif
  }
  eotf8<EOTF>(x, p);
  if (!INSIDE && !inside) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.0f;
  }
  sts128(lum, x[0], x[4], x[1], x[5]);
  sts128(lum + 16, x[2], x[6], x[3], x[7]);
}

// Temporal filters of one quad with the step's ring position K as a compile-time constant: position j holds the frame of
// age ((K - j) mod 7), position K the oldest one (age 7), X is the newest frame (age 0).
//   sustained: ages 1..7 (all seven positions; the newest frame's tap is 1e-36 of the sum), transient: X and ages 1..6.
// Then X replaces the oldest frame.
template <int K, bool EMIT>
__device__ __forceinline__ void ring_step(u64 (&ring)[RP][PXT], const u64 (&X)[PXT], const BandParams& p, u64 (&R)[2][PXT]) {
  if (EMIT) {
#pragma unroll
    for (int e = 0; e < PXT; ++e) {
      u64 a0 = 0ull, a1 = fmul2(X[e], p.wext[1][0]);
#pragma unroll
      for (int j = 0; j < RP; ++j) {
        const int age = ((K - j + RP) % RP) == 0 ? RP : (K - j + RP) % RP;
        a0 = j == 0 ? fmul2(ring[j][e], p.wext[0][age]) : ffma2(ring[j][e], p.wext[0][age], a0);
        if (j != K) a1 = ffma2(ring[j][e], p.wext[1][age], a1);
      }
      R[0][e] = a0;
      R[1][e] = a1;
    }
  }
#pragma unroll
  for (int e = 0; e < PXT; ++e) ring[K][e] = X[e];
}

// the producers' filter of one reduced-tile element: ring positions in shared memory (this thread's own element), v = newest
template <int K>
__device__ __forceinline__ void coarse_step(const float* __restrict__ ring_o, u64 v, const BandParams& p, u64& r0, u64& r1) {
  u64 a0 = 0ull, a1 = fmul2(v, p.wext[1][0]);
#pragma unroll
  for (int j = 0; j < RP; ++j) {
    const int age = ((K - j + RP) % RP) == 0 ? RP : (K - j + RP) % RP;
    const u64 x = *reinterpret_cast<const u64*>(ring_o + j * (2 * NE));
    a0 = j == 0 ? fmul2(x, p.wext[0][age]) : ffma2(x, p.wext[0][age], a0);
    if (j != K) a1 = ffma2(x, p.wext[1][age], a1);
  }
  r0 = a0;
  r1 = a1;
}

template <int KIND, bool FOV>
__global__ void __launch_bounds__(NT, 1) band_ws_kernel(const __grid_constant__ BandParams p) {
  using LY = Layout<KIND, FOV>;
  constexpr bool LANDING = LY::LANDING;
  constexpr int NLB = LY::NLB;
  extern __shared__ __align__(128) float smem[];
  float* sL = smem + LY::oL;
  float* sV0 = smem + LY::oV;
  float* sNr = smem + LY::oNr;
  float* sNc = smem + LY::oNc;
  float* sTab = smem + LY::oTab;
  float2* sFov = reinterpret_cast<float2*>(smem + LY::oFov);
  float* sView = smem + LY::oView;
  __shared__ __align__(8) u64 bars[4 + 2 * 4 + 4];  // [0..3] tile landed (TMA), [4..7] full (producers -> consumers), [8..11] empty,
                                                    // [12..13] EOTF pass done, [14..15] row pass done (among the producer warps)

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bx, by, bz;  // read once through volatile asm (see fused::band_kernel)
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
  asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
  const int tx0 = bx * TW, ty0 = by * TH;
  const int jx0 = tx0 >> 1, jy0 = ty0 >> 1;
  const int h = p.h, w = p.w, h2 = p.h2, w2 = p.w2;
  const int f_lo = bz * p.chunk, f_hi = min(f_lo + p.chunk, p.n_frames);
  const int s_lo = f_lo, s_hi = f_hi + p.fl - 1;  // slots walked by this CTA
  // Replicate padding repeats the first frame through the warm-up slots (fvvdp.py:259-260): iteration 0 of the CTAs that
  // start at slot 0 stages and reduces it ONCE and fills every ring position (and the pyramid slots of the repeats) with
  // it; iteration 1 continues at the first slot that differs.
  const int dup = (s_lo == 0) ? min(p.dup_prefix, p.fl - 2) : 0;
  const int n_iter = s_hi - s_lo - dup;
  unsigned bar0 = smem_u32(&bars[0]), sL_u32 = smem_u32(sL), sRaw_u32 = smem_u32(smem + LY::oRaw);
  asm volatile("" : "+r"(bar0), "+r"(sL_u32), "+r"(sRaw_u32));
  const unsigned bar_full = bar0 + 32, bar_empty = bar0 + 64, bar_conv = bar0 + 96, bar_row = bar0 + 112;
  // ring position of the first slot (the position follows the frame's index in the clip), advanced as the slots go by
  int rp_first = (s_lo + p.ring_phase_ws) % RP;

  auto slot_of = [&](int i) { return s_lo + i + (i > 0 ? dup : 0); };
  // start staging the tile of iteration i (lb = i mod NLB)
  auto issue_load = [&](int i, int lb) {
    const int slot = slot_of(i);
    if (LANDING) {
      const int rb = i & 1;
      const unsigned bar = bar0 + 8 * rb, dst = sRaw_u32 + rb * (TILE_FLOATS * 4);
      mbar_expect_tx(bar, TILE_FLOATS * 4);
      tma_load_3d(dst, &p.tmap_ws[0], bar, tx0 - 4, ty0 - 4, (int)p.slot_frame[0][slot]);
      tma_load_3d(dst + PLANE * 4, &p.tmap_ws[1], bar, tx0 - 4, ty0 - 4, (int)p.slot_frame[1][slot]);
    } else {
      const unsigned bar = bar0 + 8 * lb;
      mbar_expect_tx(bar, TILE_FLOATS * 4);
      tma_load_3d(sL_u32 + lb * (TILE_FLOATS * 4), &p.tmap_ws[0], bar, 2 * (tx0 - 4), ty0 - 4, slot);
    }
  };

  if (tid == 0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mbar_init(bar0 + 8 * i, 1);
      mbar_init(bar_full + 8 * i, NPW);
      mbar_init(bar_empty + 8 * i, NCW);
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      mbar_init(bar_conv + 8 * i, NPW);
      mbar_init(bar_row + 8 * i, NPW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // the first tiles are requested at once, before the tables below are loaded and the roles split: their latency is the
    // first thing every CTA waits for
    for (int i = 0; i < LY::AHEAD && i < n_iter; ++i) issue_load(i, i);
  }
  if (FOV && tid < 64) {
    // foveated: the Y and eccentricity axes of the CSF table as (x[j], 1 / (x[j+1] - x[j] + 1e-6)) pairs (interp.py:11-20)
    const int ax = 1 + (tid >> 5), j = tid & 31;
    reinterpret_cast<float2*>(sTab)[tid] = make_float2(__ldg(p.ax.x[ax] + j), j < 31 ? __ldg(p.ax.inv[ax] + j + 1) : 0.0f);
  }
  if (!FOV && tid < 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid + 1);
    // p.cell holds {Y_log, 1/step, t0, dt0, t1, dt1, 0, 0}; here the two channels' entries and steps sit side by side
    reinterpret_cast<float4*>(sTab)[2 * tid] = make_float4(a.x, a.y, 0.0f, 0.0f);
    reinterpret_cast<float4*>(sTab)[2 * tid + 1] = make_float4(a.z, b.x, a.w, b.y);
  }
  for (int i = tid; i < RP * 2 * NE; i += NT) sNr[i] = 0.0f;  // window positions that are never loaded must hold finite values
  __syncthreads();

  if (warp >= NCW) {
    // =========================================================================================== producer warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(PRODUCER_REGS));
    const int ptid = tid - NCT;
    const float K0 = 0.05f, K1 = 0.25f, K3 = 0.25f, K4 = 0.05f;
    const bool rows_interior = (jy0 - 1 >= 1) && (jy0 + TH / 2 <= h2 - 2);
    const bool cols_interior = (jx0 - 1 >= 1) && (jx0 + TW / 2 <= w2 - 2);
    float vmin = 0.0f, vmax = 1.0f;  // range of the raw level-0 samples this thread converted
    // LANDING: which of this thread's 4-pixel position chunks lie in the image (outside it the tile holds the zero padding)
    unsigned inside_mask = 0u;
    if (LANDING) {
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        const int pc = ptid + i * NPT;
        if (pc < NPC) {
          const int r = pc / LV4, c4 = pc % LV4;
          const int y = ty0 - 4 + r, x = tx0 - 4 + 4 * c4;
          if (y >= 0 && y < h && x >= 0 && x < w) inside_mask |= 1u << i;
        }
      }
    }
    // column-pass outputs of this thread: source offset in sV (floats) | edge flags << 28, global offset in Pn or -1
    int cl_src[NCOL], cl_g[NCOL];
#pragma unroll
    for (int i = 0; i < NCOL; ++i) {
      const int o = ptid + i * NPT;
      cl_src[i] = -1; cl_g[i] = -1;
      if (o < NE) {
        const int a = o / NW, b = o % NW;
        const int ic = min(max(jx0 - 1 + b, 0), w2 - 1);   // expand clamps the coarse index
        const int flags = (ic == 0 ? 1 : 0) | (ic == w2 - 1 ? 2 : 0);
        cl_src[i] = (2 * (a * LW + 2 * (ic - jx0) + 2)) | (flags << 28);
        const int j = jy0 - 1 + a, ii = jx0 - 1 + b;
        if (a >= 1 && a <= TH / 2 && b >= 1 && b <= TW / 2 && j < h2 && ii < w2) cl_g[i] = j * p.pitch2 + 2 * ii;
      }
    }
    // row pass: column pair and first reduced row of this thread (ptid < ROW_THREADS)
    const int rw_cp = ptid % ROW_CP, rw_a0 = ROW_SEG * (ptid / ROW_CP);

    const bool halo_inside = (ty0 >= 4) && (ty0 + TH + 4 <= h) && (tx0 >= 4) && (tx0 + TW + 4 <= w);

    // ---- stage C of iteration j: reduce, columns -> ring position rp (+ next level out), then the temporal filters of the
    //      same elements (this thread's own: no barrier in between) -> sNc[lb]; hand the frame to the consumers
    auto stage_c = [&](int j, int lb, unsigned lpar, int rp) {
      const int s = slot_of(j);
      const bool first_dup = (j == 0) && dup > 0;
      const bool emit = s >= f_lo + p.fl - 1;
      mbar_wait(bar_row + 8 * (j & 1), (j >> 1) & 1);                // every warp has finished the row pass of iteration j
      if (!LANDING) mbar_wait(bar_empty + 8 * lb, lpar ^ 1);          // sNc[lb] is free (LANDING: waited for before the EOTF pass)
      const float* sV = sV0 + (j & 1) * (2 * NH * LW);
      float* gout = (p.Pn != nullptr && s >= s_lo + ((bz > 0) ? p.fl - 1 : 0)) ? p.Pn + (long long)s * p.Pn_slot_stride : nullptr;
      float* nc = sNc + lb * (4 * NE);
#pragma unroll
      for (int k = 0; k < NCOL; ++k) {
        if (k < NCOL - 1 || cl_src[k] >= 0) {  // only the last round is partial
          const int o = ptid + k * NPT;
          const float* v = sV + (cl_src[k] & 0xFFFFFFF);
          const ulonglong2 v01 = *reinterpret_cast<const ulonglong2*>(v), v23 = *reinterpret_cast<const ulonglong2*>(v + 4);
          const u64 v4 = *reinterpret_cast<const u64*>(v + 8);
          u64 ov = tap5(v01.x, v01.y, v23.x, v23.y, v4);
          if (!cols_interior) {
            float ot = lo_of(ov), orf = hi_of(ov);
            if (cl_src[k] & (1 << 28)) { ot += K1 * v[4] + K0 * v[6]; orf += K1 * v[5] + K0 * v[7]; }
            if (cl_src[k] & (2 << 28)) {
              const float* e = sV + ((cl_src[k] & 0xFFFFFFF) / (2 * LW)) * (2 * LW) + 2 * (w - 1 - tx0 + 4);  // y[w-1] of this row
              // keyed on the ROW count, fvvdp_lpyr_dec.py:202
              ot += p.h_odd ? (K3 * e[0] + K4 * e[-2]) : K4 * e[0];
              orf += p.h_odd ? (K3 * e[1] + K4 * e[-1]) : K4 * e[1];
            }
            ov = pk(ot, orf);
          }
          if (gout != nullptr && cl_g[k] >= 0) *reinterpret_cast<u64*>(gout + cl_g[k]) = ov;
          float* ring_o = sNr + 2 * o;
          if (first_dup) {
#pragma unroll
            for (int q = 0; q < RP; ++q) *reinterpret_cast<u64*>(ring_o + q * (2 * NE)) = ov;
            for (int d = 1; d <= dup; ++d)
              if (gout != nullptr && cl_g[k] >= 0) *reinterpret_cast<u64*>(gout + d * p.Pn_slot_stride + cl_g[k]) = ov;
          } else {
            if (emit) {
              u64 r0, r1;
              FVVDP_RING_SWITCH(rp, (coarse_step<J>(ring_o, ov, p, r0, r1)))
              *reinterpret_cast<u64*>(nc + 2 * o) = r0;
              *reinterpret_cast<u64*>(nc + 2 * NE + 2 * o) = r1;
            }
            *reinterpret_cast<u64*>(ring_o + rp * (2 * NE)) = ov;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_full + 8 * lb);
    };

    // The producers' own stages are software-pipelined: the column pass and filters of frame i-1 sit between the EOTF pass of
    // frame i and the row pass of frame i, and the two hand-offs between the stages (EOTF -> rows, rows -> columns) are split
    // arrive / wait pairs on mbarriers, so a warp that is early at one of them has the other stage's work to do meanwhile.
    int lb = 0, rp = rp_first;          // luminance buffer / ring position of iteration i
    unsigned lpar = 0;                  // phase parity of the full / empty barriers of buffer lb in iteration i
    int plb = 0, prp = 0;               // the same of iteration i - 1
    unsigned plpar = 0;
    for (int i = 0; i < n_iter; ++i) {
      const float* sLb = sL + lb * TILE_FLOATS;
      float* sV = sV0 + (i & 1) * (2 * NH * LW);
      if (i == 1) rp = (slot_of(1) + p.ring_phase_ws) % RP;
      // ---- stage A: display EOTF, landing buffer -> luminance tile of iteration i
      if (LANDING) {
        mbar_wait(bar_empty + 8 * lb, lpar ^ 1);   // the consumers are done with the frame that used this buffer
        mbar_wait(bar0 + 8 * (i & 1), (i >> 1) & 1);
        const unsigned raw = sRaw_u32 + (i & 1) * (TILE_FLOATS * 4) + 16 * ptid, lum = sL_u32 + lb * (TILE_FLOATS * 4) + 32 * ptid;
#define FVVDP_EOTF_PASS(E)                                                                                                     \
  if (halo_inside) {                                                                                                           \
    _Pragma("unroll") for (int k = 0; k < NLD; ++k)                                                                            \
      if (k < NLD - 1 || ptid + k * NPT < NPC) eotf_chunk<E, true>(raw + k * (NPT * 16), lum + k * (NPT * 32), true, p, vmin, vmax); \
  } else {                                                                                                                     \
    _Pragma("unroll") for (int k = 0; k < NLD; ++k)                                                                            \
      if (k < NLD - 1 || ptid + k * NPT < NPC)                                                                                 \
        eotf_chunk<E, false>(raw + k * (NPT * 16), lum + k * (NPT * 32), (inside_mask >> k) & 1u, p, vmin, vmax);              \
  }
        switch (p.eotf) {  // uniform; one specialised conversion loop per EOTF
          case FVVDP_B200_EOTF_NONE: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_NONE) break;
          case FVVDP_B200_EOTF_SRGB: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_SRGB) break;
          case FVVDP_B200_EOTF_GAMMA: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_GAMMA) break;
          case FVVDP_B200_EOTF_PQ: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_PQ) break;
          case FVVDP_B200_EOTF_LINEAR: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_LINEAR) break;
          default: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_ABSOLUTE) break;
        }
#undef FVVDP_EOTF_PASS
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_conv + 8 * (i & 1));
      }
      // ---- stage C of the previous iteration
      if (i > 0) stage_c(i - 1, plb, plpar, prp);
      // ---- stage B: reduce, rows: sV[a][c] = sum_k K[k] L[2a+k][c] (zero padding + edge terms)
      if (LANDING) {
        mbar_wait(bar_conv + 8 * (i & 1), (i >> 1) & 1);   // luminance tile complete; the landing buffer may be refilled
        if (ptid == 0 && i + 2 < n_iter) issue_load(i + 2, 0);
      } else {
        mbar_wait(bar0 + 8 * lb, lpar);
      }
      if (ptid < ROW_THREADS) {
        const float* col = sLb + 4 * rw_cp;   // two adjacent columns, (test, ref) pairs
        float* out = sV + 4 * rw_cp;
        if (rows_interior) {  // no clamped coarse rows, no edge terms: sliding 5-row window over two columns
          const float* g = col + (2 * rw_a0) * (2 * LW);
          ulonglong2 g0 = *reinterpret_cast<const ulonglong2*>(g), g1 = *reinterpret_cast<const ulonglong2*>(g + 2 * LW),
                     g2 = *reinterpret_cast<const ulonglong2*>(g + 4 * LW);
#pragma unroll
          for (int j = 0; j < ROW_SEG; ++j) {
            const ulonglong2 g3 = *reinterpret_cast<const ulonglong2*>(g + (2 * j + 3) * (2 * LW)),
                             g4 = *reinterpret_cast<const ulonglong2*>(g + (2 * j + 4) * (2 * LW));
            ulonglong2 o;
            o.x = tap5(g0.x, g1.x, g2.x, g3.x, g4.x);
            o.y = tap5(g0.y, g1.y, g2.y, g3.y, g4.y);
            *reinterpret_cast<ulonglong2*>(out + (rw_a0 + j) * (2 * LW)) = o;
            g0 = g2; g1 = g3; g2 = g4;
          }
        } else {
          for (int j = 0; j < ROW_SEG; ++j) {
            const int a = rw_a0 + j;
            const int jc = min(max(jy0 - 1 + a, 0), h2 - 1);  // expand clamps the coarse index
            const float* g = col + (2 * (jc - jy0) + 2) * (2 * LW);
#pragma unroll
            for (int st = 0; st < 4; ++st) {  // two columns x two streams
              float v = fmaf(K0, g[st] + g[8 * LW + st], fmaf(K1, g[2 * LW + st] + g[6 * LW + st], 0.4f * g[4 * LW + st]));
              if (jc == 0) v += K1 * g[4 * LW + st] + K0 * g[6 * LW + st];     // x[0], x[1]   (fvvdp_lpyr_dec.py:191)
              if (jc == h2 - 1) {
                const float* e = col + (h - 1 - ty0 + 4) * (2 * LW) + st;      // x[h-1]
                v += (h & 1) ? (K3 * e[0] + K4 * e[-2 * LW]) : K4 * e[0];      // (:192-195)
              }
              out[a * (2 * LW) + st] = v;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_row + 8 * (i & 1));
      if (!LANDING && ptid == 0 && i + LY::AHEAD < n_iter) {
        // the tile of iteration i + AHEAD goes into the buffer of iteration i - 2: every warp is past the row pass of that
        // iteration (stage C above waited for the one after it), and the consumers released it long ago (the wait is a formality)
        const int tb = (lb + LY::AHEAD) % NLB;
        if (i >= 2) mbar_wait(bar_empty + 8 * tb, (lb + LY::AHEAD >= NLB) ? lpar : (lpar ^ 1u));
        issue_load(i + LY::AHEAD, tb);
      }
      plb = lb; plpar = lpar; prp = rp;
      if (++lb == NLB) { lb = 0; lpar ^= 1u; }
      rp = (rp + 1 == RP) ? 0 : rp + 1;
    }
    stage_c(n_iter - 1, plb, plpar, prp);
    if (KIND != IN_PYRAMID_TMA && eotf_checks_range(p.eotf) && (vmin < 0.0f || vmax > 1.0f) && p.flags) atomicOr(p.flags, 1u);
    return;
  }

  // ============================================================================================= consumer warps
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(CONSUMER_REGS));
  const int tile = by * gridDim.x + bx;
  // the quad of this thread
  // (32-row tiles: warp = quad row; 16-row tiles: two warps per quad row, `half` selects the pixel row of the quad)
  const int qa = PXT == 4 ? warp : (warp >> 1), qb = lane, half = PXT == 4 ? 0 : (warp & 1);
  const int qy = ty0 + 2 * qa + half, qx = tx0 + 2 * qb;
  const int coff = 2 * ((4 + 2 * qa + half) * LW + 4 + 2 * qb);   // first pixel in the luminance tile (floats)
  const int noff = 2 * (qa * NW + qb);                     // top-left of its 3x3 coarse neighbourhood
  const bool tile_full = (ty0 + TH <= h) && (tx0 + TW <= w);
  bool valid[PXT];
#pragma unroll
  for (int e = 0; e < PXT; ++e) valid[e] = (qy + (PXT == 4 ? (e >> 1) : 0) < h) && (qx + (e & 1) < w);

  if (FOV) {
    // per-pixel constants of the time walk: view direction and the rho cell / fraction of the CSF look-up
    // (rho = rho_band * resolution magnification, fvvdp.py:436-438)
#pragma unroll
    for (int e = 0; e < PXT; ++e) {
      const int x = min(qx + (e & 1), w - 1), y = min(qy + (PXT == 4 ? (e >> 1) : 0), h - 1);
      // stock display geometry only (the view direction separates into a column and a row term, fvvdp_display_model.py:498-510);
      // the maps of a fvvdp_display_geometry subclass go through the fused kernel
      const float vx = __ldg(p.vx + x), vy = __ldg(p.vy + y);
      const float va = fminf(sqrtf(vx * vx + vy * vy), 89.9f) * 0.017453292519943295f;
      const float res_mag = p.res_k0 / (__cosf(va) * __cosf(va + p.res_delta_rad));
      const float rq = fast_log2(fminf(fmaxf(p.rho_band * res_mag, p.ax.lo[0]), p.ax.hi[0]));
      int ii;
      float fr;
      locate_direct(rq, p.ax.x[0], p.ax.inv[0], p.ax.x0[0], p.ax.inv_dx[0], ii, fr);
      sFov[e * NCT + tid] = make_float2(fr, __int_as_float(ii * 1024));
      if (warp == 0 || (PXT == 2 && warp == 1)) sView[2 * qb + (e & 1)] = vx;              // the tile's columns (any quad row has them all)
      if (lane == 0) sView[TW + 2 * qa + half + (PXT == 4 ? (e >> 1) : 0)] = vy;          // the tile's rows
    }
    asm volatile("bar.sync 2, %0;" ::"n"(NCT) : "memory");  // the consumers' view-direction tables are complete
  }

  u64 ring[RP][PXT];
#pragma unroll
  for (int k = 0; k < RP; ++k)
#pragma unroll
    for (int e = 0; e < PXT; ++e) ring[k][e] = 0ull;

  // per-frame partial sums: one warp-shuffle butterfly per frame, written per warp ([frame][channel][tile][warp]; final_kernel adds
  // them in double precision).  The two channel sums share the butterfly: after the first exchange the lower half-warp carries
  // channel 0, the upper half channel 1.
  u64 pend2 = 0ull;
  int pend_row = -1;
  auto warp_sums = [&](u64 a2, int row) {
    const bool up = lane >= 16;
    float v = (up ? hi_of(a2) : lo_of(a2)) + __shfl_xor_sync(0xffffffffu, up ? lo_of(a2) : hi_of(a2), 16);
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((lane & 15) == 0) p.partial[((long long)((f_lo + row) * 2 + (up ? 1 : 0)) * p.ntiles + tile) * NCW + warp] = v;
  };
  int rp = rp_first, lb = 0;
  unsigned lpar = 0;
  for (int i = 0; i < n_iter; ++i, lpar ^= (lb + 1 == NLB ? 1u : 0u), lb = (lb + 1 == NLB ? 0 : lb + 1)) {
    const int s = s_lo + i + (i > 0 ? dup : 0);
    if (i == 1) rp = (s + p.ring_phase_ws) % RP;
    if (!LANDING) mbar_wait(bar0 + 8 * lb, lpar);   // the tile itself was written by TMA
    mbar_wait_backoff(bar_full + 8 * lb, lpar);
    const float* sLb = sL + lb * TILE_FLOATS;
    u64 X[PXT];
    {
      const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(sLb + coff);
      X[0] = r0.x; X[1] = r0.y;
      if (PXT == 4) {
        const ulonglong2 r1 = *reinterpret_cast<const ulonglong2*>(sLb + coff + 2 * LW);
        X[PXT - 2] = r1.x; X[PXT - 1] = r1.y;
      }
    }
    const bool emit = s >= f_lo + p.fl - 1;
    u64 R[2][PXT];
    if (!emit) {
      if (i == 0 && dup > 0) {  // every ring position starts as the first frame
#pragma unroll
        for (int k = 0; k < RP; ++k)
#pragma unroll
          for (int e = 0; e < PXT; ++e) ring[k][e] = X[e];
      } else {
#pragma unroll
        for (int k = 0; k < RP; ++k)
#pragma unroll
          for (int e = 0; e < PXT; ++e) ring[k][e] = (k == rp) ? X[e] : ring[k][e];
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_empty + 8 * lb);
      rp = (rp + 1 == RP) ? 0 : rp + 1;
      continue;
    }
    FVVDP_RING_SWITCH(rp, (ring_step<J, true>(ring, X, p, R)))
    rp = (rp + 1 == RP) ? 0 : rp + 1;
    const int fi = s - (p.fl - 1);  // output frame
    if (pend_row >= 0) warp_sums(pend2, pend_row);  // the previous frame's sums (independent of everything around it)

    // ---- expand of the filtered coarse tile (both temporal channels) -> bands; then the buffers go back to the producers
    const u64 c01 = pk(0.1f, 0.1f), c08 = pk(0.8f, 0.8f), c05 = pk(0.5f, 0.5f), cm1 = pk(-1.0f, -1.0f);
    u64 Bp[2][PXT];  // band (G_l - E), (test, reference)
    float Lb[PXT];
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
      const float* n = sNc + lb * (4 * NE) + cc * (2 * NE) + noff;
      u64 ve[3], vo[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const u64 n0 = *reinterpret_cast<const u64*>(n + 2 * c), n1 = *reinterpret_cast<const u64*>(n + 2 * (NW + c)),
                  n2 = *reinterpret_cast<const u64*>(n + 2 * (2 * NW + c));
        if (PXT == 4 || half == 0) ve[c] = ffma2(c08, n1, fmul2(c01, fadd2(n0, n2)));  // even row: taps 2K[0], 2K[2], 2K[4]
        if (PXT == 4 || half == 1) vo[c] = fmul2(c05, fadd2(n1, n2));                  // odd row:  taps 2K[1], 2K[3]
      }
      u64 E[PXT];
      if (PXT == 4) {
        E[0] = ffma2(c08, ve[1], fmul2(c01, fadd2(ve[0], ve[2])));
        E[1] = fmul2(c05, fadd2(ve[1], ve[2]));
        E[PXT - 2] = ffma2(c08, vo[1], fmul2(c01, fadd2(vo[0], vo[2])));
        E[PXT - 1] = fmul2(c05, fadd2(vo[1], vo[2]));
      } else {
        if (half) { ve[0] = vo[0]; ve[1] = vo[1]; ve[2] = vo[2]; }
        E[0] = ffma2(c08, ve[1], fmul2(c01, fadd2(ve[0], ve[2])));
        E[1] = fmul2(c05, fadd2(ve[1], ve[2]));
      }
#pragma unroll
      for (int e = 0; e < PXT; ++e) {
        Bp[cc][e] = ffma2(E[e], cm1, R[cc][e]);  // R - E, rounded once like the scalar subtraction
        if (cc == 0) Lb[e] = fmaxf(hi_of(E[e]), 0.1f);  // L_bkg = expanded sustained reference (:264-266)
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_empty + 8 * lb);
    if (!tile_full) {  // pixels outside the image: a zero band gives log2|T' - R'| = -inf and D^beta = 0
#pragma unroll
      for (int e = 0; e < PXT; ++e)
        if (!valid[e]) { Bp[0][e] = 0ull; Bp[1][e] = 0ull; }
    }

    // ---- contrast, CSF, masking, pooling.  Everything that is the same formula for the two temporal channels runs on
    //      (sustained, transient) pairs; with S' = S * sensitivity_correction, m = band multiplier, c = 10^mask_c:
    //        base   = log2 S' + log2 m - log2 L_bkg
    //        log2 |T' - R'| = log2 |bT - bR| + base,   log2 M = log2 min(|bT|, |bR|) + base + log2 c      (fvvdp.py:583-588)
    //        beta log2 D = beta p log2|T' - R'| - beta log2(1 + M^q), capped at beta log2 1e4              (:593-595)
    u64 acc2 = 0ull;
#pragma unroll
    for (int e = 0; e < PXT; ++e) {
      const float lgL = fast_log2(Lb[e]);
      const float yq = fminf(lgL, p.lg_y_hi);
      u64 lS2;  // log2 S' of the two channels
      if (!FOV) {
        const int cj = min((int)((yq - p.y0) * p.inv_dy), 30) * 8;  // L_bkg >= 0.1 lies above the first axis point: no lower clamp
        const float2 xi = *reinterpret_cast<const float2*>(sTab + cj);
        const float fj = (yq - xi.x) * xi.y;
        const ulonglong2 td = *reinterpret_cast<const ulonglong2*>(sTab + cj + 4);  // (t0, t1), (dt0, dt1)
        lS2 = ffma2(pk(fj, fj), td.y, td.x);
      } else {
        const float2 fc = sFov[e * NCT + tid];
        int jj, kk;
        float fy, fe;
        locate_smem(yq, sTab, p.ax.x0[1], p.ax.inv_dx[1], jj, fy);
        const float ex = sView[2 * qb + (e & 1)] - p.gaze[fi][0], ey = sView[TW + 2 * qa + half + (PXT == 4 ? (e >> 1) : 0)] - p.gaze[fi][1];
        const float ecc = fast_sqrt(fmaf(ex, ex, ey * ey));  // eccentricity [deg] (fvvdp.py:432)
        const float eq = fast_sqrt(fminf(fmaxf(ecc, p.ax.lo[2]), p.ax.hi[2]));
        locate_smem(eq, sTab + 64, p.ax.x0[2], p.ax.inv_dx[2], kk, fe);
        // trilinear look-up of both temporal channels: 4 (rho, ecc) corners, each record holds the Y entry and its step
        const float4* v = p.lut4 + __float_as_int(fc.y) + kk * 32 + jj;
        const float4 c00 = __ldg(v), c01v = __ldg(v + 32), c10 = __ldg(v + 1024), c11 = __ldg(v + 1056);
        const float fr = fc.x;
        float ls[2];
#pragma unroll
        for (int c2 = 0; c2 < 2; ++c2) {
          const float t00 = c2 ? fmaf(fy, c00.w, c00.z) : fmaf(fy, c00.y, c00.x), t01 = c2 ? fmaf(fy, c01v.w, c01v.z) : fmaf(fy, c01v.y, c01v.x);
          const float t10 = c2 ? fmaf(fy, c10.w, c10.z) : fmaf(fy, c10.y, c10.x), t11 = c2 ? fmaf(fy, c11.w, c11.z) : fmaf(fy, c11.y, c11.x);
          const float lo = fmaf(fr, t10 - t00, t00), hi = fmaf(fr, t11 - t01, t01);  // along rho at ecc cell kk, kk + 1
          ls[c2] = fmaf(fe, hi - lo, lo);
        }
        lS2 = pk(ls[0], ls[1]);
      }
      const float ce = p.log2_m - lgL;
      const u64 base2 = fadd2(lS2, pk(ce, ce));
      // T_f = min(band/L_bkg, 1000) * m  (fvvdp_lpyr_dec.py:268, :57-63); T/N = T_f * S  (fvvdp.py:583-584)
      const float lim = 1000.0f * Lb[e];
      float lgd[2], lgm[2];
#pragma unroll
      for (int cc = 0; cc < 2; ++cc) {
        const float bT = fminf(lo_of(Bp[cc][e]), lim), bR = fminf(hi_of(Bp[cc][e]), lim);
        lgd[cc] = fast_log2(fabsf(bT - bR));
        lgm[cc] = fast_log2(fminf(fabsf(bT), fabsf(bR)));
      }
      const u64 X = ffma2(p.m_bp, pk(lgd[0], lgd[1]), fmul2(p.m_bp, base2));          // beta p log2 |T' - R'|
      const u64 qM = ffma2(p.m_q, pk(lgm[0], lgm[1]), ffma2(p.m_q, base2, p.m_qlmc));   // q log2 M
      const u64 one2 = pk(1.0f, 1.0f);
      const u64 s1 = fadd2(pk(fast_exp2(lo_of(qM)), fast_exp2(hi_of(qM))), one2);       // 1 + M^q
      const u64 lD = ffma2(p.m_nbeta, pk(fast_log2(lo_of(s1)), fast_log2(hi_of(s1))), X);
      u64 D2 = pk(fast_exp2(fminf(lo_of(lD), p.m_cap)), fast_exp2(fminf(hi_of(lD), p.m_cap)));   // D^beta
      acc2 = fadd2(acc2, D2);
    }
    pend2 = acc2;      // reduced over the warp in the next iteration, under the latency of its shared-memory loads
    pend_row = fi - f_lo;
  }
  if (pend_row >= 0) warp_sums(pend2, pend_row);
}

}  // namespace ws / ws16
}  // namespace fvvdp
