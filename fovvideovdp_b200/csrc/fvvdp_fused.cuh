// Fused band kernel of the B200-native FovVideoVDP core (sm_100a): ONE kernel per pyramid level that
//   * stages the level's luminance tile (+4 px halo) of the NEXT frames into shared memory while the current one is
//     processed: TMA (cp.async.bulk.tensor, zero fill outside the image, mbarrier completion) for the pyramid levels and
//     for contiguous float input frames, cp.async for float frames TMA cannot address; every other input format
//     (uint8, uint16, RGB, strided) is converted to luminance planes in the pyramid layout by luminance_kernel first;
//     level 0 applies the display EOTF on the way from the landing buffer to the luminance tile,
//   * reduces the tile to the next Gaussian level (separable 5-tap, stride 2; fvvdp_lpyr_dec.py:183-207) and writes that
//     level out for the next launch,
//   * keeps the last `fl` frames of both streams ON CHIP while it walks through time: the tile's own pixels in a
//     register ring, the reduced tile in a shared-memory ring; a frame stays in the ring position given by its index in
//     the clip and the filter WEIGHTS rotate from step to step (one code version of the filters, and rounding that does
//     not depend on how the clip is cut into blocks),
//   * applies the sustained / transient temporal filters (fvvdp.py:294-300) to both rings, expands the filtered
//     reduced tile (fvvdp_lpyr_dec.py:219-235), forms the contrast bands (:259-269), looks up the CSF
//     (fvvdp.py:520-537), applies the masking model (:574-596) and accumulates sum D^beta (:467,598-607).
//
// The reference filters in time first and then builds one pyramid per temporal channel (4 channels).  Reduce and
// expand are linear, so  pyr(sum_k w_k L_{t-k}) = sum_k w_k pyr(L_{t-k}):  here the pyramid is built ONCE per
// luminance frame and stream and the temporal filter is applied to its levels.
//
// Data layout: everything after the EOTF is a (test, reference) PAIR of floats per pixel -- the luminance tile, the
// row-reduced tile, both rings, the pyramid planes in HBM ([slot][row][2 * column + stream]).  The kernel is bound by
// instruction issue, not by HBM, and both streams go through exactly the same stencils and filters, so
// one packed fp32x2 instruction (fma/add/mul.rn.f32x2 -> FFMA2/FADD2/FMUL2) does the work of two.
//
// Border semantics follow the reference exactly: zero padding + additive edge terms for the reduce (including the
// row-parity quirk of fvvdp_lpyr_dec.py:202), index clamping for the expand.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved at run time)

#include "fvvdp_common.cuh"

#ifndef FVVDP_WAIT_HINT_NS
#define FVVDP_WAIT_HINT_NS 20000
#endif

namespace fvvdp {
namespace fused {

constexpr int TH = 16, TW = 64;                 // output tile of one CTA
constexpr int LH = TH + 8, LW = TW + 8;         // staged luminance tile: origin (ty0-4, tx0-4)
constexpr int NH = TH / 2 + 2, NW = TW / 2 + 2; // reduced tile with 1-px halo: origin (jy0-1, jx0-1)
constexpr int NE = NH * NW;                     // 340
constexpr int RING = 8;                         // temporal window kept on chip by the 256-thread kernels (one 2x2 quad per thread)
constexpr int MAXRING = 16;                     // ... and by the 512-thread kernels (two pixels per thread; 60 fps clips)
constexpr int MAXCHUNK = FVVDP_B200_MAX_BLOCK_FRAMES;  // output frames walked by one CTA
constexpr int LV4 = LW / 4;                     // 4-pixel chunks per staged row
constexpr int NPC = LH * LV4;                   // 4-pixel position chunks of a staged tile (432)
// threads per CTA / pixels per thread: up to 8 taps the thread owns a 2x2 quad (8 x 4 ring registers pairs), beyond that
// one row of the quad (16 x 2), so that the register ring stays at 64 registers
__host__ __device__ constexpr int threads_of(int FL) { return FL > RING ? 512 : 256; }
__host__ __device__ constexpr int pixels_of(int FL) { return FL > RING ? 2 : 4; }
__host__ __device__ constexpr int nld_of(int FL) { return (NPC + threads_of(FL) - 1) / threads_of(FL); }   // position chunks per thread and frame
__host__ __device__ constexpr int ncol_of(int FL) { return (NE + threads_of(FL) - 1) / threads_of(FL); }   // column-pass outputs per thread
constexpr int PLANE = LH * LW;                  // one stream of a landing buffer
constexpr int TILE_FLOATS = 2 * PLANE;          // one staged tile, both streams
constexpr int ROW_SEGS = 3;                     // row pass: the NH coarse rows are split 4 + 3 + 3 over three threads per column
constexpr int ROW_THREADS = ROW_SEGS * LW;      // 216

typedef unsigned long long u64;  // two packed floats (lo = test, hi = reference)

enum InputKind { IN_LEVEL0_CPASYNC = 0, IN_PYRAMID_TMA = 2, IN_LEVEL0_TMA = 3 };  // any other level-0 input: luminance_kernel first

struct BandParams {
  // ---- TMA descriptors: [0] pyramid planes (3-D: 2x+stream, y, slot) or level-0 test frames (3-D: x, y, frame); [1] level-0 reference frames
  CUtensorMap tmap[2];
  CUtensorMap tmap_ws[2];                     // the same tensors with the staged-tile box of the warp-specialised kernel (fvvdp_ws.cuh)
  // ---- input ----
  const void* slot[2][FVVDP_B200_MAX_SLOTS];   // level 0 without TMA: [test|ref][slot] frame base pointers
  unsigned short slot_frame[2][FVVDP_B200_MAX_SLOTS];  // level 0 with TMA: frame coordinate of each slot
  // ---- output ----
  float* Pn;                                  // next level: [slot][h2][pitch2] floats, (test, ref) interleaved, or nullptr (last scored band)
  long long Pn_slot_stride;                   // floats between slots (= h2 * pitch2)
  int pitch2;                                 // row pitch of Pn in floats (multiple of 4, >= 2 * w2)
  float* partial;                             // [n_frames][2][ntiles]
  // ---- geometry / schedule ----
  int h, w, h2, w2, h_odd, ntiles;
  int n_frames, fl, chunk;                    // output frames; filter taps (<= ring length of the kernel); output frames per CTA
  int dup_prefix;                             // slots 1 .. dup_prefix hold the same frame as slot 0 (replicate padding)
  int ring_phase;                             // slot s sits at ring position (s + ring_phase) mod ring length: the position follows
                                              //   the frame's index in the CLIP, so the summation order of the temporal filters (and
                                              //   with it every rounding) does not depend on how the clip is cut into blocks / ranks
  int ring_phase_ws;                          // the same for the 7-position rings of the warp-specialised kernel
  int ch2_slots_in, ch2_slots_out;            // CH2 (see band_kernel): planes of temporal channel 1 start this many slots after channel 0's
  u64 wext[2][2 * MAXRING];                    // [temporal channel][i]: (w, w) packed weight of AGE (i mod ring length), 0 = newest frame;
                                              //   ring position j has age (rp - j) mod ring length when the newest frame sits at position rp: weight wext[rp + RL - j]
  // ---- level-0 input format ----
  long long sC, sH, sW;
  int C, dtype, eotf;
  float Yscale, Y_black, Y_peak, gamma, L_min, L_max;
  float rgb2y[3];
  uint32_t* flags;
  // ---- CSF / masking ----
  const float* cell;                          // [32][8] per band: Y_log[j], 1/(Y_log[j+1]-Y_log[j]+1e-6), t0[j], t0[j+1]-t0[j], t1[j], t1[j+1]-t1[j]
  float y0, inv_dy, lg_y_hi;                  // uniform first guess of the Y cell; log2 of the upper clamp
  float log2_m;                               // log2(band multiplier), fvvdp_lpyr_dec.py:57-63
  float band_mul;
  float mask_p, mask_q[2], log2_mask_c, beta, w_transient;
  // the same constants as (sustained, transient) pairs for the packed masking maths of the warp-specialised kernel:
  u64 m_q, m_qlmc, m_bp, m_nbeta;             // (q0, q1); (q0, q1) * log2 c; (beta p, beta p); (-beta, -beta)
  float m_cap;                                // beta * log2(1e4)
  // foveated
  CsfAxes ax;
  const float4* lut4;                         // [rho][ecc][Y] (t0, dt0, t1, dt1): log2(S * sens_mul) of both temporal channels + step to Y+1
  float log2_sens_mul, rho_band;
  const float* vx;
  const float* vy;
  const float* vmap;                          // custom display geometry: view direction per pixel [2][h][w] (deg), else nullptr
  const float* rqmap;                         // custom display geometry: log2(clamp(rho_band * res_mag)) per pixel [h][w]
  float res_k0, res_delta_rad;
  float gaze[FVVDP_B200_MAX_BLOCK_FRAMES][2];
  // ---- optional outputs (EXTRA) ----
  float* tapR;   // level 0: [F][NCH][h][w]
  float* tapG;   // [F][NCH][h2][w2]  temporally filtered next Gaussian level
  float* tapC;   // [F][NCH][h][w]
  float* tapL;   // [F][h][w]
  float* tapS;   // [F][TC][h][w]
  float* tapD;   // [F][TC][h][w]
  float* dmap;   // [F][h][w]
  float* ctxmap; // level 0: [F][h][w] sustained test frame = context image of the heat-map visualisation (fvvdp.py:475)
};

// ------------------------------------------------------------------------------------------------ packed fp32x2
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ float lo_of(u64 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return lo;
}
__device__ __forceinline__ float hi_of(u64 v) {
  float lo, hi;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
  return hi;
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// separable 5-tap [.05 .25 .4 .25 .05] on pairs, in the operation order of the scalar fmaf(K0, g0+g4, fmaf(K1, g1+g3, K2*g2))
__device__ __forceinline__ u64 tap5(u64 g0, u64 g1, u64 g2, u64 g3, u64 g4) {
  const u64 k0 = pk(0.05f, 0.05f), k1 = pk(0.25f, 0.25f), k2 = pk(0.4f, 0.4f);
  return ffma2(k0, fadd2(g0, g4), ffma2(k1, fadd2(g1, g3), fmul2(k2, g2)));
}

// ------------------------------------------------------------------------------------------------ shared memory / async copies
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ float4 lds128(unsigned a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
__device__ __forceinline__ void sts128(unsigned a, float x, float y, float z, float w) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}

__device__ __forceinline__ void cp_async16(unsigned smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  // try_wait with a suspend-time hint: the warp sleeps in hardware until the phase completes (or the hint expires) instead of
  // spinning through the issue slots of the warps that have work
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "FVVDP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
      "@p bra FVVDP_DONE;\n\t"
      "bra FVVDP_WAIT;\n\t"
      "FVVDP_DONE:\n\t}" ::"r"(bar), "r"(parity), "r"(FVVDP_WAIT_HINT_NS)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(unsigned dst, const CUtensorMap* map, unsigned bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}

// ------------------------------------------------------------------------------------------------ display EOTF
// EOTF -> luminance for one sample (fvvdp_display_model.py:147-165, 203-212).  KIND is a compile-time fvvdp_b200_eotf.
// The callers track the range of the raw values in vmin/vmax ("Pixel outside the valid range 0-1", :149-151).
template <int KIND>
__device__ __forceinline__ float eotf_k(float v, const BandParams& p) {
  if (KIND == FVVDP_B200_EOTF_NONE) return v;
  if (KIND == FVVDP_B200_EOTF_ABSOLUTE) return fminf(fmaxf(v, p.L_min), p.L_max);
  if (KIND == FVVDP_B200_EOTF_LINEAR) return fminf(fmaxf(v, 0.005f), p.Y_peak) + p.Y_black;
  if (KIND == FVVDP_B200_EOTF_SRGB) {
    // clamp(v,0,1) folded into saturating arithmetic: (v + 0.055)/1.055 maps [0,1] into [0.052,1], and values below
    // 0.04045 take the linear branch, so saturating the affine result / the linear product is the same clamp
    const float t = __saturatef(fmaf(v, 1.0f / 1.055f, 0.055f / 1.055f));
    const float lin = (v > 0.04045f) ? fast_exp2(2.4f * fast_log2(t)) : __saturatef(v * (1.0f / 12.92f));
    return fmaf(p.Yscale, lin, p.Y_black);
  }
  v = __saturatef(v);
  if (KIND == FVVDP_B200_EOTF_GAMMA) {
    return fmaf(p.Yscale, fast_pow(v, p.gamma), p.Y_black);
  } else {  // PQ
    const float n_inv = 1.0f / 0.15930175781250000f, m_inv = 1.0f / 78.843750000000000f;
    const float c1 = 0.83593750000000000f, c2 = 18.851562500000000f, c3 = 18.687500000000000f;
    const float t = fast_pow(v, m_inv);
    const float L = 10000.0f * fast_pow(fmaxf(t - c1, 0.0f) / (c2 - c3 * t), n_inv);
    return fminf(fmaxf(L, 0.005f), p.Y_peak) + p.Y_black;
  }
}
// EOTFs that expect display-encoded values in [0,1] and report anything outside ("Pixel outside the valid range 0-1")
__device__ __forceinline__ bool eotf_checks_range(int kind) {
  return kind == FVVDP_B200_EOTF_SRGB || kind == FVVDP_B200_EOTF_GAMMA || kind == FVVDP_B200_EOTF_PQ;
}

// EOTF of 8 raw samples in place (4 pixels of the test stream, 4 of the reference stream)
template <int KIND>
__device__ __forceinline__ void eotf8(float (&x)[8], const BandParams& p) {
  if (KIND == FVVDP_B200_EOTF_NONE) return;
  if (KIND == FVVDP_B200_EOTF_SRGB || KIND == FVVDP_B200_EOTF_GAMMA) {
    // stage-wise with volatile MUFU ops: all lg2 back to back, then all ex2.  Left to itself the compiler predicates
    // the MUFU pair of every sample on its own (v > 0.04045) test and serialises the chains.
    float t[8];
    const float gam = KIND == FVVDP_B200_EOTF_SRGB ? 2.4f : p.gamma;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = KIND == FVVDP_B200_EOTF_SRGB ? __saturatef(fmaf(x[j], 1.0f / 1.055f, 0.055f / 1.055f)) : __saturatef(x[j]);
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(a));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float a = t[j] * gam;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(a));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float lin = t[j];
      if (KIND == FVVDP_B200_EOTF_SRGB) lin = (x[j] > 0.04045f) ? lin : __saturatef(x[j] * (1.0f / 12.92f));
      x[j] = fmaf(p.Yscale, lin, p.Y_black);
    }
  } else if (KIND == FVVDP_B200_EOTF_PQ) {
    // pq2lin (fvvdp_display_model.py:100-112) stage-wise: L = 1e4 (max(t - c1, 0) / (c2 - c3 t))^(1/n), t = V^(1/m); the
    // quotient is taken in the log2 domain (5 MUFU operations per sample, no division)
    const float n_inv = 1.0f / 0.15930175781250000f, m_inv = 1.0f / 78.843750000000000f;
    const float c1 = 0.83593750000000000f, c2 = 18.851562500000000f, c3 = 18.687500000000000f;
    float t[8], u[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(__saturatef(x[j])));
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(t[j] * m_inv));
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(u[j]) : "f"(fmaxf(t[j] - c1, 0.0f)));
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(fmaf(-c3, t[j], c2)));
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(fmaf(u[j] - t[j], n_inv, 13.287712379549449f)));
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = fminf(fmaxf(t[j], 0.005f), p.Y_peak) + p.Y_black;
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = eotf_k<KIND>(x[j], p);
  }
}

// One 4-pixel position chunk: landing buffer (two planes) -> EOTF -> luminance tile ((test, ref) interleaved).
// `inside` = the chunk lies in the image (outside it the tile holds the zero padding of the reduce).
template <int KIND>
__device__ __forceinline__ void eotf_chunk(unsigned raw, unsigned lum, bool inside, const BandParams& p, float& vmin, float& vmax) {
  const float4 a = lds128(raw), b = lds128(raw + PLANE * 4);
  float x[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
  vmin = fminf(vmin, fminf(fminf(fminf(x[0], x[1]), fminf(x[2], x[3])), fminf(fminf(x[4], x[5]), fminf(x[6], x[7]))));
  vmax = fmaxf(vmax, fmaxf(fmaxf(fmaxf(x[0], x[1]), fmaxf(x[2], x[3])), fmaxf(fmaxf(x[4], x[5]), fmaxf(x[6], x[7]))));
  eotf8<KIND>(x, p);
  if (!inside) {
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = 0.0f;
  }
  sts128(lum, x[0], x[4], x[1], x[5]);
  sts128(lum + 16, x[2], x[6], x[3], x[7]);
}

// cell of a 32-point (nearly uniform) axis containing q, and the reference's interpolation fraction
// (get_interpolants_v1, interp.py:11-20; the fraction uses the stored axis values)
__device__ __forceinline__ void locate_direct(float q, const float* __restrict__ x, const float* __restrict__ inv, float x0, float inv_dx, int& j,
                                              float& f) {
  j = min(max((int)((q - x0) * inv_dx), 0), 30);
  f = fmaxf((q - __ldg(x + j)) * __ldg(inv + j + 1), 0.0f);
}

// the same with the axis staged in shared memory as (x[j], 1 / (x[j+1] - x[j] + 1e-6)) pairs
__device__ __forceinline__ void locate_smem(float q, const float* __restrict__ ax, float x0, float inv_dx, int& j, float& f) {
  j = min(max((int)((q - x0) * inv_dx), 0), 30);
  const float2 a = *reinterpret_cast<const float2*>(ax + 2 * j);
  f = fmaxf((q - a.x) * a.y, 0.0f);
}

// ------------------------------------------------------------------------------------------------ temporal rings
template <int FL>
struct Ring {
  u64 v[FL][pixels_of(FL)];  // [ring slot][pixel of the 2x2 quad, or of its row] = (test, reference)
};

template <int FL, int J>
__device__ __forceinline__ void ring_store(Ring<FL>& ring, const float* __restrict__ sLb, int coff) {
  const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(sLb + coff);
  ring.v[J][0] = r0.x; ring.v[J][1] = r0.y;
  if (pixels_of(FL) == 4) {
    const ulonglong2 r1 = *reinterpret_cast<const ulonglong2*>(sLb + coff + 2 * LW);
    ring.v[J][2] = r1.x; ring.v[J][3] = r1.y;
  }
}

// R[cc][e] = sum_j w(age of slot j) * ring[j][e]: the frames stay in their ring slots and the WEIGHTS rotate with the step
// (wp[cc] points at wext[cc][(s mod FL) + FL]; slot j has weight wp[cc][-j]), so the filter needs no per-step code version
template <int FL, int TC>
__device__ __forceinline__ void fir_quad(const Ring<FL>& ring, const u64* const (&wp)[2], u64 (&R)[TC][pixels_of(FL)]) {
#pragma unroll
  for (int cc = 0; cc < TC; ++cc)
#pragma unroll
    for (int e = 0; e < pixels_of(FL); ++e) {
      if (FL == 1) {
        R[cc][e] = ring.v[0][e];
      } else {
        u64 a = fmul2(ring.v[0][e], wp[cc][0]);
#pragma unroll
        for (int j = 1; j < FL; ++j) a = ffma2(ring.v[j][e], wp[cc][-j], a);
        R[cc][e] = a;
      }
    }
}

// temporal filter of this thread's own elements of the reduced-tile ring (written by the same thread, so no barrier is
// needed between the column pass and this): sNc[cc][o] = sum_j w(age of slot j) sNr[j][o]
template <int FL, int TC>
__device__ __forceinline__ void fir_coarse(const float* __restrict__ sNr, float* __restrict__ sNc, const u64* const (&wp)[2], int tid) {
  constexpr int NT = threads_of(FL), NCOL = ncol_of(FL);
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    const int o = tid + i * NT;
    if (i < NCOL - 1 || o < NE) {
      u64 a[TC];
#pragma unroll
      for (int j = 0; j < FL; ++j) {
        const u64 v = *reinterpret_cast<const u64*>(sNr + j * (2 * NE) + 2 * o);
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) a[cc] = j == 0 ? fmul2(v, wp[cc][0]) : ffma2(v, wp[cc][-j], a[cc]);
      }
#pragma unroll
      for (int cc = 0; cc < TC; ++cc) *reinterpret_cast<u64*>(sNc + cc * (2 * NE) + 2 * o) = a[cc];
    }
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int KIND, int FL, int TC, bool FOV, bool EXTRA>
__global__ void __launch_bounds__(threads_of(FL), FL > RING ? 1 : 2) band_kernel(const __grid_constant__ BandParams p) {
  constexpr int NT = threads_of(FL), PXT = pixels_of(FL), NLD = nld_of(FL), NCOL = ncol_of(FL);
  constexpr bool LEVEL0 = KIND != IN_PYRAMID_TMA;
  constexpr bool TMA = KIND == IN_PYRAMID_TMA || KIND == IN_LEVEL0_TMA;
  constexpr bool LANDING = KIND == IN_LEVEL0_TMA || KIND == IN_LEVEL0_CPASYNC;  // raw planes land first, the EOTF pass interleaves them
  constexpr int NLUM = KIND == IN_PYRAMID_TMA ? 2 : 1;  // the pyramid planes are interleaved in HBM: TMA lands them as luminance tiles
  // CH2: temporal windows too long for the on-chip rings (17..32 taps).  The temporal filters are applied to the frames by a
  // register-ring walk of their own (front_kernel, pairs output) and every slot arrives as TWO (test, reference) planes, the
  // sustained and the transient channel; the kernel stages, reduces and writes out both (the reference's pyramid per temporal
  // channel, fvvdp_lpyr_dec.py:248-273) and needs no ring.
  constexpr bool CH2 = FL == 1 && TC == 2;
  static_assert(!CH2 || KIND == IN_PYRAMID_TMA, "two-channel slots are staged from pyramid-layout planes");
  constexpr int CHN = CH2 ? 2 : 1;
  constexpr int NCH = 2 * TC;
  extern __shared__ __align__(128) float smem[];
  float* sL = smem;                                  // [NLUM][CHN][LH][LW][2]   luminance tile, (test, ref) interleaved
  float* sRaw = sL + NLUM * CHN * TILE_FLOATS;       // LANDING: [2 buffers][2 streams][LH][LW]
  float* sV = sRaw + (LANDING ? 2 * TILE_FLOATS : 0);  // [CHN][NH][LW][2]   row-reduced
  float* sNr = sV + CHN * 2 * NH * LW;               // [FL][NE][2]   ring of reduced tiles (CH2: the two channels' reduced tiles)
  float* sNc = (FL == 1) ? sNr : sNr + FL * 2 * NE;  // [TC][NE][2]   temporally filtered reduced tiles
  float* sTab = sNr + (CH2 ? 2 : FL) * 2 * NE + (FL == 1 ? 0 : NCH * NE);  // [32][8]
  float* sRed = sTab + 256;                          // [MAXCHUNK][2][NT/32]
  float4* sFov = reinterpret_cast<float4*>(sRed + MAXCHUNK * 2 * (NT / 32));  // FOV: [PXT pixels][NT] (view x, view y, rho fraction, rho cell)
  __shared__ __align__(8) u64 bars[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // the CTA coordinates are read ONCE through volatile asm: left to itself the compiler re-reads %ctaid (S2UR, long
  // latency) inside the time loop instead of keeping the derived values
  int bx, by, bz;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(bx));
  asm volatile("mov.u32 %0, %%ctaid.y;" : "=r"(by));
  asm volatile("mov.u32 %0, %%ctaid.z;" : "=r"(bz));
  const int tx0 = bx * TW, ty0 = by * TH;
  const int jx0 = tx0 >> 1, jy0 = ty0 >> 1;
  const int h = p.h, w = p.w, h2 = p.h2, w2 = p.w2;
  const int f_lo = bz * p.chunk, f_hi = min(f_lo + p.chunk, p.n_frames);
  const int s_lo = f_lo, s_hi = f_hi + p.fl - 1;  // slots walked by this CTA
  const int tile = by * gridDim.x + bx;
  float vmin = 0.0f, vmax = 1.0f;  // range of the raw level-0 samples this thread converted
  // shared-window addresses, made opaque so that they are computed once (the conversion reads %cluster_ctaid: S2UR)
  unsigned bar0 = smem_u32(&bars[0]), sL_u32 = smem_u32(sL), sRaw_u32 = smem_u32(sRaw);
  asm volatile("" : "+r"(bar0), "+r"(sL_u32), "+r"(sRaw_u32));

  // ---------------- one-time set-up (all index arithmetic lives here, outside the time loop) ----------------
  if (TMA && tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (FOV && tid < 64) {
    // foveated: the Y and eccentricity axes of the CSF table as (x[j], 1 / (x[j+1] - x[j] + 1e-6)) pairs (interp.py:11-20)
    const int ax = 1 + (tid >> 5), j = tid & 31;
    reinterpret_cast<float2*>(sTab)[tid] = make_float2(__ldg(p.ax.x[ax] + j), j < 31 ? __ldg(p.ax.inv[ax] + j + 1) : 0.0f);
  }
  if (!FOV && tid < 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid + 1);
    reinterpret_cast<float4*>(sTab)[2 * tid] = a;
    reinterpret_cast<float4*>(sTab)[2 * tid + 1] = b;
  }
  for (int i = tid; i < (CH2 ? 2 : FL) * 2 * NE; i += NT) sNr[i] = 0.0f;  // window positions that are never loaded must hold finite values

  // LANDING: the 4-pixel position chunks of this thread are chunk tid (and tid + NT): 16 * chunk bytes into a landing plane,
  // 32 * chunk bytes into the luminance tile.  ld_goff = element offset inside a frame, or -1 = outside the image.
  const bool halo_inside = (ty0 >= 4) && (ty0 + TH + 4 <= h) && (tx0 >= 4) && (tx0 + TW + 4 <= w);
  int ld_goff[NLD];
  if (LANDING) {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int pc = tid + i * NT;
      ld_goff[i] = -1;
      if (pc < NPC) {
        const int r = pc / LV4, c4 = pc % LV4;
        const int y = ty0 - 4 + r, x = tx0 - 4 + 4 * c4;
        if (y >= 0 && y < h && x >= 0 && x < w) ld_goff[i] = y * (int)p.sH + x;
      }
    }
  }
  // column-pass outputs of this thread: source offset in sV (floats) | edge flags << 28, global offset in Pn or -1
  int cl_src[NCOL], cl_g[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    const int o = tid + i * NT;
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
  // row pass: column and first coarse row of this thread (tid < ROW_THREADS)
  const int rw_c = tid % LW, rw_seg = tid / LW;
  const int rw_a0 = rw_seg == 0 ? 0 : (rw_seg == 1 ? 4 : 7), rw_n = rw_seg == 0 ? 4 : 3;
  // the quad of this thread (PXT == 4), or the row `half` of the quad (PXT == 2; warp-uniform)
  const int qa = (tid >> 5) & 7, qb = lane, half = PXT == 2 ? (tid >> 8) : 0;
  const int qy = ty0 + 2 * qa + half, qx = tx0 + 2 * qb;  // first pixel of the thread
#define FVVDP_EY(e) (PXT == 4 ? ((e) >> 1) : 0)
#define FVVDP_EX(e) ((e) & 1)
  const int coff = 2 * ((4 + 2 * qa + half) * LW + 4 + 2 * qb);   // first pixel in the luminance tile (floats)
  const int noff = 2 * (qa * NW + qb);                     // top-left of its 3x3 coarse neighbourhood
  bool valid[PXT];
#pragma unroll
  for (int e = 0; e < PXT; ++e) valid[e] = (qy + FVVDP_EY(e) < h) && (qx + FVVDP_EX(e) < w);

  if (FOV) {
    // per-pixel constants of the time walk: view direction and the rho cell / fraction of the CSF look-up
    // (rho = rho_band * resolution magnification, fvvdp.py:436-438)
#pragma unroll
    for (int e = 0; e < PXT; ++e) {
      const int x = min(qx + FVVDP_EX(e), w - 1), y = min(qy + FVVDP_EY(e), h - 1);
      float vx, vy, rq;
      if (p.vmap != nullptr) {  // maps computed by a fvvdp_display_geometry subclass
        const long long po = (long long)y * w + x;
        vx = __ldg(p.vmap + po); vy = __ldg(p.vmap + (long long)h * w + po); rq = __ldg(p.rqmap + po);
      } else {
        vx = __ldg(p.vx + x); vy = __ldg(p.vy + y);
        const float va = fminf(sqrtf(vx * vx + vy * vy), 89.9f) * 0.017453292519943295f;
        const float res_mag = p.res_k0 / (__cosf(va) * __cosf(va + p.res_delta_rad));
        rq = fast_log2(fminf(fmaxf(p.rho_band * res_mag, p.ax.lo[0]), p.ax.hi[0]));
      }
      int ii;
      float fr;
      locate_direct(rq, p.ax.x[0], p.ax.inv[0], p.ax.x0[0], p.ax.inv_dx[0], ii, fr);
      sFov[e * NT + tid] = make_float4(vx, vy, fr, __int_as_float(ii * 1024));
    }
  }

  Ring<FL> ring;
#pragma unroll
  for (int k = 0; k < FL; ++k)
#pragma unroll
    for (int e = 0; e < PXT; ++e) ring.v[k][e] = 0ull;

  const float K0 = 0.05f, K1 = 0.25f, K3 = 0.25f, K4 = 0.05f;
  const bool rows_interior = (jy0 - 1 >= 1) && (jy0 + TH / 2 <= h2 - 2);
  const bool cols_interior = (jx0 - 1 >= 1) && (jx0 + TW / 2 <= w2 - 2);
  const bool tile_full = (ty0 + TH <= h) && (tx0 + TW <= w);

  // start staging the tile of `slot` into buffer `buf` (landing buffer, or luminance buffer for the pyramid levels)
  auto issue_load = [&](int slot, int buf) {
    if (TMA) {
      if (tid == 0) {
        mbar_expect_tx(bar0 + 8 * buf, CHN * TILE_FLOATS * 4);
        if (KIND == IN_PYRAMID_TMA) {
          tma_load_3d(sL_u32 + buf * (CHN * TILE_FLOATS * 4), &p.tmap[0], bar0 + 8 * buf, 2 * (tx0 - 4), ty0 - 4, slot);
          if (CH2) tma_load_3d(sL_u32 + (buf * CHN + 1) * (TILE_FLOATS * 4), &p.tmap[0], bar0 + 8 * buf, 2 * (tx0 - 4), ty0 - 4, slot + p.ch2_slots_in);
        } else {
          const unsigned dst = sRaw_u32 + buf * (TILE_FLOATS * 4);
          tma_load_3d(dst, &p.tmap[0], bar0 + 8 * buf, tx0 - 4, ty0 - 4, (int)p.slot_frame[0][slot]);
          tma_load_3d(dst + PLANE * 4, &p.tmap[1], bar0 + 8 * buf, tx0 - 4, ty0 - 4, (int)p.slot_frame[1][slot]);
        }
      }
    } else if (KIND == IN_LEVEL0_CPASYNC) {
      const float* b0 = reinterpret_cast<const float*>(p.slot[0][slot]);
      const float* b1 = reinterpret_cast<const float*>(p.slot[1][slot]);
      const unsigned dst = sRaw_u32 + buf * (TILE_FLOATS * 4) + 16 * tid;
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        if (i < NLD - 1 || tid + i * NT < NPC) {
          const int g = ld_goff[i];
          cp_async16(dst + i * (NT * 16), g >= 0 ? (const void*)(b0 + g) : (const void*)b0, g >= 0 ? 16 : 0);
          cp_async16(dst + i * (NT * 16) + PLANE * 4, g >= 0 ? (const void*)(b1 + g) : (const void*)b1, g >= 0 ? 16 : 0);
        }
      }
      cp_async_commit();
    }
  };
  // level 0: landing buffer -> display EOTF -> luminance tile
  auto convert = [&](int buf) {
    const unsigned raw = sRaw_u32 + buf * (TILE_FLOATS * 4) + 16 * tid, lum = sL_u32 + 32 * tid;
#define FVVDP_EOTF_PASS(K)                                                                                                      \
  _Pragma("unroll") for (int i = 0; i < NLD; ++i)                                                                               \
    if (tid + i * NT < NPC) eotf_chunk<K>(raw + i * (NT * 16), lum + i * (NT * 32), halo_inside || ld_goff[i] >= 0, p, vmin, vmax);
    switch (p.eotf) {  // uniform; one specialised conversion loop per EOTF
      case FVVDP_B200_EOTF_NONE: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_NONE) break;
      case FVVDP_B200_EOTF_SRGB: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_SRGB) break;
      case FVVDP_B200_EOTF_GAMMA: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_GAMMA) break;
      case FVVDP_B200_EOTF_PQ: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_PQ) break;
      case FVVDP_B200_EOTF_LINEAR: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_LINEAR) break;
      default: FVVDP_EOTF_PASS(FVVDP_B200_EOTF_ABSOLUTE) break;
    }
#undef FVVDP_EOTF_PASS
  };

  // ---- stages of one slot (used by the time loop and by the peeled first slot below) ----
  // stage A: luminance tile of `slot` complete in its buffer (this thread's part of it; a barrier follows)
  auto stage_tile = [&](int buf, int phase, bool later_in_flight) {
    if (TMA) mbar_wait(bar0 + 8 * buf, phase);
    else if (later_in_flight) cp_async_wait<1>();
    else cp_async_wait<0>();
    if (LANDING) {
      if (!TMA) __syncthreads();  // cp.async: the chunks of other threads
      convert(buf);
    }
  };
  // stage B: reduce, rows: sV[a][c] = sum_k K[k] L[2j-2+k][c],  j = clamp(jy0-1+a)  (zero padding + edge terms)
  auto rows_pass = [&](const float* sLb0, int ch) {
    const float* sLb = sLb0 + ch * TILE_FLOATS;
    float* sVc = sV + ch * (2 * NH * LW);
    if (tid < ROW_THREADS) {  // one thread walks down a third of one staged column, (test, ref) pairs
      const float* col = sLb + 2 * rw_c;
      float* out = sVc + 2 * rw_c;
      if (rows_interior) {  // no clamped coarse rows, no edge terms: sliding 5-row window
        const float* g = col + (2 * rw_a0) * (2 * LW);
        u64 g0 = *reinterpret_cast<const u64*>(g), g1 = *reinterpret_cast<const u64*>(g + 2 * LW), g2 = *reinterpret_cast<const u64*>(g + 4 * LW);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < 3 || rw_n == 4) {
            const u64 g3 = *reinterpret_cast<const u64*>(g + (2 * j + 3) * (2 * LW)), g4 = *reinterpret_cast<const u64*>(g + (2 * j + 4) * (2 * LW));
            *reinterpret_cast<u64*>(out + (rw_a0 + j) * (2 * LW)) = tap5(g0, g1, g2, g3, g4);
            g0 = g2; g1 = g3; g2 = g4;
          }
        }
      } else {
        for (int j = 0; j < rw_n; ++j) {
          const int a = rw_a0 + j;
          const int jc = min(max(jy0 - 1 + a, 0), h2 - 1);  // expand clamps the coarse index
          const float* g = col + (2 * (jc - jy0) + 2) * (2 * LW);
#pragma unroll
          for (int st = 0; st < 2; ++st) {
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
  };
  // stage C: reduce, columns -> ring position rp (+ next level out); dup > 0: every ring position and `dup` further
  // pyramid slots get the same tile (repeats of the first frame)
  auto cols_pass = [&](int s, int rp, int dup, int ch) {
    {
      const float* sVc = sV + ch * (2 * NH * LW);
      float* ring_s = sNr + (CH2 ? ch : rp) * (2 * NE);
      float* gout = (p.Pn != nullptr && s >= s_lo + ((bz > 0) ? p.fl - 1 : 0)) ? p.Pn + (long long)(s + ch * p.ch2_slots_out) * p.Pn_slot_stride : nullptr;
#pragma unroll
      for (int i = 0; i < NCOL; ++i) {
        if (i < NCOL - 1 || cl_src[i] >= 0) {  // only the last round is partial
          const float* v = sVc + (cl_src[i] & 0xFFFFFFF);
          const ulonglong2 v01 = *reinterpret_cast<const ulonglong2*>(v), v23 = *reinterpret_cast<const ulonglong2*>(v + 4);
          const u64 v4 = *reinterpret_cast<const u64*>(v + 8);
          u64 o = tap5(v01.x, v01.y, v23.x, v23.y, v4);
          if (!cols_interior) {
            float ot = lo_of(o), orf = hi_of(o);
            if (cl_src[i] & (1 << 28)) { ot += K1 * v[4] + K0 * v[6]; orf += K1 * v[5] + K0 * v[7]; }
            if (cl_src[i] & (2 << 28)) {
              const float* e = sVc + ((cl_src[i] & 0xFFFFFFF) / (2 * LW)) * (2 * LW) + 2 * (w - 1 - tx0 + 4);  // y[w-1] of this row
              // keyed on the ROW count, fvvdp_lpyr_dec.py:202
              ot += p.h_odd ? (K3 * e[0] + K4 * e[-2]) : K4 * e[0];
              orf += p.h_odd ? (K3 * e[1] + K4 * e[-1]) : K4 * e[1];
            }
            o = pk(ot, orf);
          }
          *reinterpret_cast<u64*>(ring_s + 2 * (tid + i * NT)) = o;
          if (gout != nullptr && cl_g[i] >= 0) *reinterpret_cast<u64*>(gout + cl_g[i]) = o;
          if (dup > 0) {
            for (int j = 0; j < FL; ++j) *reinterpret_cast<u64*>(sNr + j * (2 * NE) + 2 * (tid + i * NT)) = o;
            for (int d = 1; d <= dup; ++d)
              if (gout != nullptr && cl_g[i] >= 0) *reinterpret_cast<u64*>(gout + d * p.Pn_slot_stride + cl_g[i]) = o;
          }
        }
      }
    }

  };

  if (TMA) __syncthreads();  // barrier initialisation visible before the first wait
  // Replicate padding repeats the first frame through the warm-up slots (fvvdp.py:259-260): the CTAs that start at slot 0
  // stage and reduce it ONCE, copy the result into the ring positions and pyramid slots of the repeats, and start the time
  // loop at the first slot that differs.
  int s_begin = s_lo;
  {
    const int dup = (s_lo == 0) ? min(p.dup_prefix, p.fl - 2) : 0;
    if (dup > 0) {
      issue_load(0, 0);
      stage_tile(0, 0, false);
      __syncthreads();
      rows_pass(sL, 0);
      __syncthreads();
      cols_pass(0, 0, dup, 0);
      // every ring position starts as the first frame: the repeats sit where the time loop would have put them, the
      // positions beyond the window carry zero weights
      ring_store<FL, 0>(ring, sL, coff);
#pragma unroll
      for (int k = 1; k < FL; ++k)
#pragma unroll
        for (int e = 0; e < PXT; ++e) ring.v[k][e] = ring.v[0][e];
      __syncthreads();
      if (TMA) {  // fresh barrier phases for the time loop
        if (tid == 0) {
          asm volatile("mbarrier.inval.shared.b64 [%0];" ::"r"(bar0) : "memory");
          asm volatile("mbarrier.inval.shared.b64 [%0];" ::"r"(bar0 + 8) : "memory");
          mbar_init(bar0, 1);
          mbar_init(bar0 + 8, 1);
          asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
      }
      s_begin = dup + 1;
    }
  }
  issue_load(s_begin, 0);
  if (LANDING && s_begin + 1 < s_hi) issue_load(s_begin + 1, 1);

  for (int s = s_begin; s < s_hi; ++s) {
    const int buf = (s - s_begin) & 1;
    const float* sLb = sL + (NLUM == 2 ? buf * CHN * TILE_FLOATS : 0);
    stage_tile(buf, ((s - s_begin) >> 1) & 1, s + 1 < s_hi);
    __syncthreads();  // (1) luminance tile of slot s complete; every reader of the buffers refilled below is done
    if (LANDING) { if (s + 2 < s_hi) issue_load(s + 2, buf); }
    else if (KIND == IN_PYRAMID_TMA) { if (s + 1 < s_hi) issue_load(s + 1, buf ^ 1); }

    rows_pass(sLb, 0);
    if (CH2) rows_pass(sLb, 1);
    __syncthreads();  // (2)
    const int rp = (s + p.ring_phase) % FL;  // ring position of slot s
    cols_pass(s, rp, 0, 0);
    if (CH2) cols_pass(s, rp, 0, 1);

    const bool emit = s >= f_lo + p.fl - 1;
    const int fi = s - (p.fl - 1);  // output frame
    // ---- temporal filter of the thread's own coarse elements; its pixels into the register ring (the only step-dependent
    //      register index: one small switch); the filter of the pixels follows the barrier, next to the masking maths
    const u64* const wp[2] = {p.wext[0] + rp + FL, p.wext[1] + rp + FL};
    if (emit && FL > 1) fir_coarse<FL, TC>(sNr, sNc, wp, tid);
    switch (rp) {
#define FVVDP_CASE(J) case J: ring_store<FL, (J) % FL>(ring, sLb, coff); break;
      FVVDP_CASE(0) FVVDP_CASE(1) FVVDP_CASE(2) FVVDP_CASE(3) FVVDP_CASE(4) FVVDP_CASE(5) FVVDP_CASE(6) FVVDP_CASE(7)
#if FUSED_MAXRING_CASES
      FVVDP_CASE(8) FVVDP_CASE(9) FVVDP_CASE(10) FVVDP_CASE(11) FVVDP_CASE(12) FVVDP_CASE(13) FVVDP_CASE(14) FVVDP_CASE(15)
#endif
#undef FVVDP_CASE
    }
    u64 ring2[PXT];  // CH2: the thread's pixels of the second temporal channel
    if (CH2) {
      const ulonglong2 r0 = *reinterpret_cast<const ulonglong2*>(sLb + TILE_FLOATS + coff);
      ring2[0] = r0.x; ring2[1] = r0.y;
      if (PXT == 4) {
        const ulonglong2 r1 = *reinterpret_cast<const ulonglong2*>(sLb + TILE_FLOATS + coff + 2 * LW);
        ring2[PXT - 2] = r1.x; ring2[PXT - 1] = r1.y;
      }
    }
    if (NLUM == 1 || emit) __syncthreads();  // (3) filtered coarse tiles visible; the single luminance tile may be rewritten
    if (!emit) continue;

    // ---- temporal filter of the own pixels, expand of the filtered coarse tile, contrast, CSF, masking, pooling ----
    u64 R[TC][PXT];
    fir_quad<FL, TC>(ring, wp, R);
    if (CH2) {
#pragma unroll
      for (int e = 0; e < PXT; ++e) R[TC - 1][e] = ring2[e];
    }
    float acc[2] = {0.0f, 0.0f};
    float Lb[PXT], lgL[PXT], fj[PXT];
    int cj[PXT];
    float lsf[TC][PXT];  // FOV: log2 S per temporal channel
    float Dsum[PXT];
#pragma unroll
    for (int e = 0; e < PXT; ++e) Dsum[e] = 0.0f;
    const u64 c01 = pk(0.1f, 0.1f), c08 = pk(0.8f, 0.8f), c05 = pk(0.5f, 0.5f), cm1 = pk(-1.0f, -1.0f);
#pragma unroll
    for (int cc = 0; cc < TC; ++cc) {
      // expand both streams at once: every value below is a (test, reference) pair
      const float* n = sNc + cc * (2 * NE) + noff;
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
      float B[2][PXT];  // band (G_l - E) of the test / reference channel
#pragma unroll
      for (int e = 0; e < PXT; ++e) {
        const u64 b = ffma2(E[e], cm1, R[cc][e]);  // R - E, rounded once like the scalar subtraction
        B[0][e] = lo_of(b);
        B[1][e] = hi_of(b);
      }
      if (cc == 0) {
#pragma unroll
        for (int e = 0; e < PXT; ++e) {
          Lb[e] = fmaxf(hi_of(E[e]), 0.1f);  // L_bkg = expanded sustained reference (:264-266)
          lgL[e] = fast_log2(Lb[e]);
          const float yq = fminf(lgL[e], p.lg_y_hi);
          if (!FOV) {
            cj[e] = min((int)((yq - p.y0) * p.inv_dy), 30) * 8;  // L_bkg >= 0.1 lies above the first axis point: no lower clamp
            const float2 xi = *reinterpret_cast<const float2*>(sTab + cj[e]);
            fj[e] = (yq - xi.x) * xi.y;
          } else {
            const float4 fc = sFov[e * NT + tid];
            int jj, kk;
            float fy, fe;
            locate_smem(yq, sTab, p.ax.x0[1], p.ax.inv_dx[1], jj, fy);
            const float ex = fc.x - p.gaze[fi][0], ey = fc.y - p.gaze[fi][1];
            const float ecc = fast_sqrt(fmaf(ex, ex, ey * ey));  // eccentricity [deg] (fvvdp.py:432)
            const float eq = fast_sqrt(fminf(fmaxf(ecc, p.ax.lo[2]), p.ax.hi[2]));
            locate_smem(eq, sTab + 64, p.ax.x0[2], p.ax.inv_dx[2], kk, fe);
            // trilinear look-up of both temporal channels: 4 (rho, ecc) corners, each record holds the Y entry and its step
            const float4* v = p.lut4 + __float_as_int(fc.w) + kk * 32 + jj;
            const float4 c00 = __ldg(v), c01 = __ldg(v + 32), c10 = __ldg(v + 1024), c11 = __ldg(v + 1056);
            const float fr = fc.z;
#pragma unroll
            for (int c2 = 0; c2 < TC; ++c2) {
              const float t00 = c2 ? fmaf(fy, c00.w, c00.z) : fmaf(fy, c00.y, c00.x), t01 = c2 ? fmaf(fy, c01.w, c01.z) : fmaf(fy, c01.y, c01.x);
              const float t10 = c2 ? fmaf(fy, c10.w, c10.z) : fmaf(fy, c10.y, c10.x), t11 = c2 ? fmaf(fy, c11.w, c11.z) : fmaf(fy, c11.y, c11.x);
              const float lo = fmaf(fr, t10 - t00, t00), hi = fmaf(fr, t11 - t01, t01);  // along rho at ecc cell kk, kk + 1
              lsf[c2][e] = fmaf(fe, hi - lo, lo);
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < PXT; ++e) {
        float lS;  // log2 of (sensitivity x sensitivity_correction)
        if (!FOV) {
          const float2 td = *reinterpret_cast<const float2*>(sTab + cj[e] + 2 + 2 * cc);
          lS = fmaf(fj[e], td.y, td.x);
        } else {
          lS = lsf[cc][e];
        }
        // T_f = min(band/L_bkg, 1000) * m  (:268, :57-63); T/N = T_f * S  (fvvdp.py:583-584)
        const float lim = 1000.0f * Lb[e];
        const float bT = fminf(B[0][e], lim), bR = fminf(B[1][e], lim);
        const float lSL = lS + (p.log2_m - lgL[e]);
        const float ld = fast_log2((tile_full || valid[e]) ? fabsf(bT - bR) : 0.0f) + lSL;  // log2 |T' - R'|
        const float lM = fast_log2(fminf(fabsf(bT), fabsf(bR))) + (lSL + p.log2_mask_c);  // log2 M  (:588)
        const float Mq = fast_exp2(p.mask_q[cc] * lM);
        const float lD = fminf(fmaf(p.mask_p, ld, -fast_log2(1.0f + Mq)), 13.287712379549449f);  // D <= 1e4 (:593-595)
        acc[cc] += fast_exp2(p.beta * lD);
        if (EXTRA && valid[e]) {
          const long long plane = (long long)h * w;
          const long long pofs = (long long)(qy + FVVDP_EY(e)) * w + qx + FVVDP_EX(e);
          const float invL = 1.0f / Lb[e];
          if (p.tapC) {
            p.tapC[((long long)fi * NCH + cc * 2 + 0) * plane + pofs] = bT * invL * p.band_mul;
            p.tapC[((long long)fi * NCH + cc * 2 + 1) * plane + pofs] = bR * invL * p.band_mul;
          }
          if (p.tapS) p.tapS[((long long)fi * TC + cc) * plane + pofs] = fast_exp2(lS);
          const float D = fast_exp2(lD);
          if (p.tapD) p.tapD[((long long)fi * TC + cc) * plane + pofs] = D;
          Dsum[e] += (cc == 0 ? 1.0f : p.w_transient) * D;
          if (cc == 0 && p.tapL) p.tapL[(long long)fi * plane + pofs] = Lb[e];
          if (p.tapR) {
            p.tapR[((long long)fi * NCH + cc * 2 + 0) * plane + pofs] = lo_of(R[cc][e]);
            p.tapR[((long long)fi * NCH + cc * 2 + 1) * plane + pofs] = hi_of(R[cc][e]);
          }
          if (cc == TC - 1 && p.dmap) p.dmap[(long long)fi * plane + pofs] = Dsum[e] / p.band_mul;
          if (cc == 0 && p.ctxmap) p.ctxmap[(long long)fi * plane + pofs] = lo_of(R[0][e]);
        }
      }
    }
    if (EXTRA && p.tapG) {  // temporally filtered next Gaussian level (interior of the coarse tile)
      for (int o = tid; o < NCH * NE; o += NT) {
        const int cc = o / (2 * NE), rem = o % (2 * NE), el = rem >> 1, st = rem & 1, a = el / NW, b = el % NW;
        const int j = jy0 - 1 + a, ii = jx0 - 1 + b;
        if (a >= 1 && a <= TH / 2 && b >= 1 && b <= TW / 2 && j < h2 && ii < w2)
          p.tapG[(((long long)fi * NCH + cc * 2 + st) * h2 + j) * w2 + ii] = sNc[o];
      }
    }
    // ---- per-frame partial sums: warp shuffle now, one pass over the warps at the end.  The two channel sums share
    //      the butterfly: after the first exchange the lower half-warp carries channel 0, the upper half channel 1.
    {
      const bool up = lane >= 16;
      float v = (up ? acc[1] : acc[0]) + __shfl_xor_sync(0xffffffffu, up ? acc[0] : acc[1], 16);
#pragma unroll
      for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if ((lane & 15) == 0) sRed[((fi - f_lo) * 2 + (up ? 1 : 0)) * (NT / 32) + warp] = v;
    }
  }
  __syncthreads();
  for (int i = tid; i < (f_hi - f_lo) * 2; i += NT) {
    float v = 0.0f;
#pragma unroll
    for (int k = 0; k < NT / 32; ++k) v += sRed[i * (NT / 32) + k];
    const int fi = f_lo + (i >> 1), cc = i & 1;
    p.partial[((long long)fi * 2 + cc) * p.ntiles + tile] = v;
  }
  if (LEVEL0 && eotf_checks_range(p.eotf) && (vmin < 0.0f || vmax > 1.0f) && p.flags) atomicOr(p.flags, 1u);
#undef FVVDP_EY
#undef FVVDP_EX
}

template <int KIND, int FL, int TC, bool FOV>
constexpr size_t band_smem_bytes() {
  constexpr int NT = threads_of(FL);
  constexpr int CHN = (FL == 1 && TC == 2) ? 2 : 1;  // CH2: two planes per slot
  return (FOV ? sizeof(float4) * pixels_of(FL) * NT : 0) + sizeof(float) * (size_t)((KIND == IN_PYRAMID_TMA ? 2 : 1) * CHN * TILE_FLOATS + ((KIND == IN_LEVEL0_TMA || KIND == IN_LEVEL0_CPASYNC) ? 2 * TILE_FLOATS : 0) +
                                  CHN * 2 * NH * LW + (CHN == 2 ? 2 : FL) * 2 * NE + (FL == 1 ? 0 : 2 * TC * NE) + 256 + MAXCHUNK * 2 * (NT / 32));
}

}  // namespace fused
}  // namespace fvvdp
