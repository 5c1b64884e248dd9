// Fused band kernel of the B200-native FovVideoVDP core (sm_100a): ONE kernel per pyramid level that
//   * stages the level's luminance tile (+4 px halo) of the NEXT frame into shared memory while the current one is
//     processed: TMA (cp.async.bulk.tensor, zero fill outside the image, mbarrier completion) for the pyramid levels and
//     for contiguous float input frames, cp.async / plain loads for everything else (uint8, RGB, strided);
//     level 0 applies the display EOTF in shared memory,
//   * reduces the tile to the next Gaussian level (separable 5-tap, stride 2; fvvdp_lpyr_dec.py:183-207) and writes that
//     level out for the next launch,
//   * keeps the last `fl` frames of both streams ON CHIP while it walks through time: the tile's own pixels in a
//     register ring, the reduced tile in a shared-memory ring,
//   * applies the sustained / transient temporal filters (fvvdp.py:294-300) to both rings, expands the filtered
//     reduced tile (fvvdp_lpyr_dec.py:219-235), forms the contrast bands (:259-269), looks up the CSF
//     (fvvdp.py:520-537), applies the masking model (:574-596) and accumulates sum D^beta (:467,598-607).
//
// The reference filters in time first and then builds one pyramid per temporal channel (4 channels).  Reduce and
// expand are linear, so  pyr(sum_k w_k L_{t-k}) = sum_k w_k pyr(L_{t-k}):  here the pyramid is built ONCE per
// luminance frame and stream (2 planes) and the temporal filter is applied to its levels.  Per frame pair this moves
// 2 input planes + 2 planes per coarser level through HBM instead of the 4-channel R tensor and 4-channel levels.
//
// The kernel is bound by instruction issue, not by HBM, so the arithmetic uses Blackwell's packed
// fp32x2 instructions (fma/add/mul.rn.f32x2 -> FFMA2/FADD2/FMUL2) wherever two lanes share an operation: pixel pairs
// in the temporal filter, (test, reference) pairs in the expand.
//
// Border semantics follow the reference exactly: zero padding + additive edge terms for the reduce (including the
// row-parity quirk of fvvdp_lpyr_dec.py:202), index clamping for the expand.
#pragma once
#include <cuda.h>  // CUtensorMap (type only; the driver entry point is resolved at run time)

#include "fvvdp_common.cuh"

namespace fvvdp {
namespace fused {

constexpr int TH = 16, TW = 64;                 // output tile of one CTA
constexpr int LH = TH + 8, LW = TW + 8;         // staged luminance tile: origin (ty0-4, tx0-4)
constexpr int NH = TH / 2 + 2, NW = TW / 2 + 2; // reduced tile with 1-px halo: origin (jy0-1, jx0-1)
constexpr int NE = NH * NW;                     // 340
constexpr int NT = 256;                         // threads: one 2x2 quad each
constexpr int RING = 8;                         // temporal window kept on chip
constexpr int MAXCHUNK = 64;                    // output frames walked by one CTA
constexpr int LV4 = LW / 4;                     // 16-byte chunks per staged row
constexpr int NLD = (2 * LH * LV4 + NT - 1) / NT;  // 16-byte chunks per thread and frame (4)
constexpr int NCOL = (2 * NE + NT - 1) / NT;       // column-pass outputs per thread (3)
constexpr int TILE_FLOATS = 2 * LH * LW;           // one staged buffer: [stream][LH][LW]

typedef unsigned long long u64;  // two packed floats (lo, hi)

enum InputKind { IN_LEVEL0_CPASYNC = 0, IN_LEVEL0_GENERIC = 1, IN_PYRAMID_TMA = 2, IN_LEVEL0_TMA = 3 };

struct BandParams {
  // ---- TMA descriptors: [0] pyramid planes (4-D: x, y, stream, slot) or level-0 test frames (3-D: x, y, frame); [1] level-0 reference frames
  CUtensorMap tmap[2];
  // ---- input ----
  const void* slot[2][FVVDP_B200_MAX_SLOTS];   // level 0 without TMA: [test|ref][slot] frame base pointers
  unsigned short slot_frame[2][FVVDP_B200_MAX_SLOTS];  // level 0 with TMA: frame coordinate of each slot
  const float* P;                             // level >= 1: luminance pyramid planes [slot][2][h][pitch]
  long long P_slot_stride;                    // floats between slots (= 2 * h * pitch)
  int pitch;                                  // row pitch of P in floats (multiple of 4)
  // ---- output ----
  float* Pn;                                  // [slot][2][h2][pitch2] or nullptr (last scored band)
  long long Pn_slot_stride;
  int pitch2;
  float* partial;                             // [n_frames][2][ntiles]
  // ---- geometry / schedule ----
  int h, w, h2, w2, h_odd, ntiles;
  int n_frames, fl, chunk;                    // output frames; filter taps (<= RING); output frames per CTA
  u64 wgt2[2][RING];                          // [temporal channel][ring window position, 0 = oldest]: (w, w) packed
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
  // foveated
  CsfAxes ax;
  const float* lut3d;
  float log2_sens_mul, rho_band;
  const float* vx;
  const float* vy;
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

// ------------------------------------------------------------------------------------------------ async copies
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void cp_async16(float* smem_dst, const void* gsrc, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(u64* bar, int count) {
  asm volatile("mbarrier.init.shared.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.release.cta.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(u64* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "FVVDP_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra FVVDP_DONE;\n\t"
      "bra FVVDP_WAIT;\n\t"
      "FVVDP_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_4d(float* dst, const CUtensorMap* map, u64* bar, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(float* dst, const CUtensorMap* map, u64* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

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

__device__ __forceinline__ float eotf_one(float v, const BandParams& p, float& vmin, float& vmax) {
  vmin = fminf(vmin, v);
  vmax = fmaxf(vmax, v);
  switch (p.eotf) {
    case FVVDP_B200_EOTF_NONE: return v;
    case FVVDP_B200_EOTF_ABSOLUTE: return eotf_k<FVVDP_B200_EOTF_ABSOLUTE>(v, p);
    case FVVDP_B200_EOTF_LINEAR: return eotf_k<FVVDP_B200_EOTF_LINEAR>(v, p);
    case FVVDP_B200_EOTF_SRGB: return eotf_k<FVVDP_B200_EOTF_SRGB>(v, p);
    case FVVDP_B200_EOTF_GAMMA: return eotf_k<FVVDP_B200_EOTF_GAMMA>(v, p);
    default: return eotf_k<FVVDP_B200_EOTF_PQ>(v, p);
  }
}

__device__ __forceinline__ float lum_generic(const BandParams& p, const void* base, int y, int x, float& vmin, float& vmax) {
  const long long off = (long long)y * p.sH + (long long)x * p.sW;
  if (p.C == 3) {
    const float r = eotf_one(load_sample(base, off, p.dtype), p, vmin, vmax);
    const float g = eotf_one(load_sample(base, off + p.sC, p.dtype), p, vmin, vmax);
    const float b = eotf_one(load_sample(base, off + 2 * p.sC, p.dtype), p, vmin, vmax);
    return r * p.rgb2y[0] + g * p.rgb2y[1] + b * p.rgb2y[2];
  }
  return eotf_one(load_sample(base, off, p.dtype), p, vmin, vmax);
}

// in-place EOTF of this thread's 16-byte chunks of the staged tile (level 0, contiguous float input)
template <int KIND>
__device__ __forceinline__ void eotf_chunks(float* dst, const int (&ld_soff)[NLD], const int (&ld_goff)[NLD], const BandParams& p, float& vmin,
                                            float& vmax) {
  // all loads first, then the arithmetic of all 16 samples, then all stores: written as one load-convert-store per
  // chunk the stores could alias the next load and the chunks would serialise
  float4 v[NLD];
  bool ok[NLD];
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    ok[i] = ld_soff[i] >= 0 && ld_goff[i] >= 0;
    if (ok[i]) v[i] = *reinterpret_cast<const float4*>(dst + (ld_soff[i] & 0xFFFFFF));
    else v[i] = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
  }
#pragma unroll
  for (int i = 0; i < NLD; ++i) {
    vmin = fminf(vmin, fminf(fminf(v[i].x, v[i].y), fminf(v[i].z, v[i].w)));  // 3-input min/max on sm_100
    vmax = fmaxf(vmax, fmaxf(fmaxf(v[i].x, v[i].y), fmaxf(v[i].z, v[i].w)));
  }
  if (KIND == FVVDP_B200_EOTF_SRGB || KIND == FVVDP_B200_EOTF_GAMMA) {
    // stage-wise over the 16 samples with volatile MUFU ops: all lg2 back to back, then all ex2.  Left to itself the
    // compiler predicates the MUFU pair of every sample on its own (v > 0.04045) test and serialises the 16 chains.
    float* x = reinterpret_cast<float*>(v);
    float t[4 * NLD];
    const float gam = KIND == FVVDP_B200_EOTF_SRGB ? 2.4f : p.gamma;
#pragma unroll
    for (int j = 0; j < 4 * NLD; ++j) {
      const float a = KIND == FVVDP_B200_EOTF_SRGB ? __saturatef(fmaf(x[j], 1.0f / 1.055f, 0.055f / 1.055f)) : __saturatef(x[j]);
      asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(a));
    }
#pragma unroll
    for (int j = 0; j < 4 * NLD; ++j) {
      const float a = t[j] * gam;
      asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(t[j]) : "f"(a));
    }
#pragma unroll
    for (int j = 0; j < 4 * NLD; ++j) {
      float lin = t[j];
      if (KIND == FVVDP_B200_EOTF_SRGB) lin = (x[j] > 0.04045f) ? lin : __saturatef(x[j] * (1.0f / 12.92f));
      x[j] = fmaf(p.Yscale, lin, p.Y_black);
    }
  } else {
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      v[i].x = eotf_k<KIND>(v[i].x, p); v[i].y = eotf_k<KIND>(v[i].y, p); v[i].z = eotf_k<KIND>(v[i].z, p); v[i].w = eotf_k<KIND>(v[i].w, p);
    }
  }
#pragma unroll
  for (int i = 0; i < NLD; ++i)
    if (ok[i]) *reinterpret_cast<float4*>(dst + (ld_soff[i] & 0xFFFFFF)) = v[i];
}

// cell of a 32-point (nearly uniform) axis containing q, and the reference's interpolation fraction
// (get_interpolants_v1, interp.py:11-20; the fraction uses the stored axis values)
__device__ __forceinline__ void locate_direct(float q, const float* __restrict__ x, const float* __restrict__ inv, float x0, float inv_dx, int& j,
                                              float& f) {
  j = min(max((int)((q - x0) * inv_dx), 0), 30);
  f = fmaxf((q - __ldg(x + j)) * __ldg(inv + j + 1), 0.0f);
}

// ------------------------------------------------------------------------------------------------ temporal rings
template <int FL>
struct Ring {
  u64 v[2][FL][2];  // [stream][ring slot][row of the quad] = (left, right) pixel
};

template <int FL, int J>
__device__ __forceinline__ void ring_store(Ring<FL>& ring, const float* __restrict__ sLb, int coff) {
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    ring.v[s][J][0] = *reinterpret_cast<const u64*>(sLb + s * LH * LW + coff);
    ring.v[s][J][1] = *reinterpret_cast<const u64*>(sLb + s * LH * LW + coff + LW);
  }
}

// R[cc*2+s][row] = sum_k wgt[cc][k] * ring[s][(J+1+k) % FL][row]   (window position k = 0 is the oldest frame)
template <int FL, int TC, int J>
__device__ __forceinline__ void fir_quad(const Ring<FL>& ring, const BandParams& p, u64 (&R)[2 * TC][2]) {
#pragma unroll
  for (int cc = 0; cc < TC; ++cc)
#pragma unroll
    for (int s = 0; s < 2; ++s)
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        if (FL == 1) {
          R[cc * 2 + s][r] = ring.v[s][0][r];
        } else {
          u64 a = fmul2(ring.v[s][(J + 1) % FL][r], p.wgt2[cc][0]);
#pragma unroll
          for (int k = 1; k < FL; ++k) a = ffma2(ring.v[s][(J + 1 + k) % FL][r], p.wgt2[cc][k], a);
          R[cc * 2 + s][r] = a;
        }
      }
}

// temporal filter of the reduced tiles: sNc[cc][i] = sum_k wgt[cc][k] sNr[(J+1+k) % FL][i], i over [NE][stream]
template <int FL, int TC, int J>
__device__ __forceinline__ void fir_coarse(const float* __restrict__ sNr, float* __restrict__ sNc, const BandParams& p, int tid) {
  if (tid < 2 * NE / 4) {  // four consecutive floats per thread, 16-byte shared accesses
    const float* base = sNr + 4 * tid;
    u64 a[2][2];
#pragma unroll
    for (int k = 0; k < FL; ++k) {
      const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(base + ((J + 1 + k) % FL) * (2 * NE));
#pragma unroll
      for (int cc = 0; cc < TC; ++cc) {
        a[cc][0] = k == 0 ? fmul2(v.x, p.wgt2[cc][0]) : ffma2(v.x, p.wgt2[cc][k], a[cc][0]);
        a[cc][1] = k == 0 ? fmul2(v.y, p.wgt2[cc][0]) : ffma2(v.y, p.wgt2[cc][k], a[cc][1]);
      }
    }
#pragma unroll
    for (int cc = 0; cc < TC; ++cc) *reinterpret_cast<ulonglong2*>(sNc + cc * (2 * NE) + 4 * tid) = make_ulonglong2(a[cc][0], a[cc][1]);
  }
}

// ------------------------------------------------------------------------------------------------ the kernel
template <int KIND, int FL, int TC, bool FOV, bool EXTRA>
__global__ void __launch_bounds__(NT, 2) band_kernel(const __grid_constant__ BandParams p) {
  constexpr bool LEVEL0 = KIND != IN_PYRAMID_TMA;
  constexpr bool TMA = KIND == IN_PYRAMID_TMA || KIND == IN_LEVEL0_TMA;
  constexpr bool CHUNKED = KIND != IN_LEVEL0_GENERIC;  // staged as 16-byte chunks of contiguous float rows
  constexpr int NCH = 2 * TC;
  extern __shared__ __align__(128) float smem[];
  float* sL = smem;                                  // [2 buffers][2 streams][LH][LW]
  float* sV = sL + 2 * TILE_FLOATS;                  // [2][NH][LW]   row-reduced
  float* sNr = sV + 2 * NH * LW;                     // [FL][NE][2]   ring of reduced tiles, (test, ref) interleaved
  float* sNc = (FL == 1) ? sNr : sNr + FL * 2 * NE;  // [TC][NE][2]   temporally filtered reduced tiles
  float* sTab = sNr + FL * 2 * NE + (FL == 1 ? 0 : NCH * NE);  // [32][8]
  float* sRed = sTab + 256;                          // [MAXCHUNK][2][NT/32]
  __shared__ __align__(8) u64 bars[2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int jx0 = tx0 >> 1, jy0 = ty0 >> 1;
  const int h = p.h, w = p.w, h2 = p.h2, w2 = p.w2;
  const int f_lo = blockIdx.z * p.chunk, f_hi = min(f_lo + p.chunk, p.n_frames);
  const int s_lo = f_lo, s_hi = f_hi + p.fl - 1;  // slots walked by this CTA
  const int tile = blockIdx.y * gridDim.x + blockIdx.x;
  float vmin = 0.0f, vmax = 1.0f;  // range of the raw level-0 samples this thread converted

  // ---------------- one-time set-up (all index arithmetic lives here, outside the time loop) ----------------
  if (TMA && tid == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (tid < 32) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid);
    const float4 b = __ldg(reinterpret_cast<const float4*>(p.cell) + 2 * tid + 1);
    reinterpret_cast<float4*>(sTab)[2 * tid] = a;
    reinterpret_cast<float4*>(sTab)[2 * tid + 1] = b;
  }
  for (int i = tid; i < FL * 2 * NE; i += NT) sNr[i] = 0.0f;  // window positions that are never loaded must hold finite values

  // 16-byte chunks of the staged tile owned by this thread: shared offset | stream << 30 (or -1), element offset
  // inside a frame (or -1 = outside the image: zero fill, no EOTF)
  int ld_soff[NLD], ld_goff[NLD];
  if (CHUNKED) {
    const int pitch = LEVEL0 ? (int)p.sH : p.pitch;
#pragma unroll
    for (int i = 0; i < NLD; ++i) {
      const int item = tid + i * NT;
      ld_soff[i] = -1;
      ld_goff[i] = -1;
      if (item < 2 * LH * LV4) {
        const int s = item / (LH * LV4), rem = item % (LH * LV4), r = rem / LV4, c4 = rem % LV4;
        const int y = ty0 - 4 + r, x = tx0 - 4 + 4 * c4;
        ld_soff[i] = ((s * LH + r) * LW + 4 * c4) | (s << 30);
        if (y >= 0 && y < h && x >= 0 && x < w) ld_goff[i] = y * pitch + x;
      }
    }
  }
  // column-pass outputs of this thread
  int cl_src[NCOL], cl_dst[NCOL], cl_g[NCOL];
#pragma unroll
  for (int i = 0; i < NCOL; ++i) {
    const int o = tid + i * NT;
    cl_src[i] = -1; cl_dst[i] = 0; cl_g[i] = -1;
    if (o < 2 * NE) {
      const int s = o / NE, rem = o % NE, a = rem / NW, b = rem % NW;
      const int ic = min(max(jx0 - 1 + b, 0), w2 - 1);   // expand clamps the coarse index
      const int flags = (ic == 0 ? 1 : 0) | (ic == w2 - 1 ? 2 : 0);
      cl_src[i] = ((s * NH + a) * LW + 2 * (ic - jx0) + 2) | (flags << 28);
      cl_dst[i] = 2 * rem + s;
      const int j = jy0 - 1 + a, ii = jx0 - 1 + b;
      if (a >= 1 && a <= TH / 2 && b >= 1 && b <= TW / 2 && j < h2 && ii < w2) cl_g[i] = (s * h2 + j) * p.pitch2 + ii;
    }
  }
  // the quad of this thread
  const int qa = tid >> 5, qb = lane;
  const int qy = ty0 + 2 * qa, qx = tx0 + 2 * qb;
  const int coff = (4 + 2 * qa) * LW + 4 + 2 * qb;   // quad's top-left pixel in the staged tile
  const int noff = 2 * (qa * NW + qb);               // top-left of its 3x3 coarse neighbourhood ((test, ref) interleaved)
  bool valid[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) valid[e] = (qy + (e >> 1) < h) && (qx + (e & 1) < w);

  Ring<FL> ring;
#pragma unroll
  for (int s = 0; s < 2; ++s)
#pragma unroll
    for (int k = 0; k < FL; ++k) ring.v[s][k][0] = ring.v[s][k][1] = 0ull;

  const float K0 = 0.05f, K1 = 0.25f, K2 = 0.4f, K3 = 0.25f, K4 = 0.05f;
  const bool rows_interior = (jy0 - 1 >= 1) && (jy0 + TH / 2 <= h2 - 2);
  const bool cols_interior = (jx0 - 1 >= 1) && (jx0 + TW / 2 <= w2 - 2);
  const bool tile_full = (ty0 + TH <= h) && (tx0 + TW <= w);

  // stage the tile of `slot` into buffer `buf`
  auto issue_load = [&](int slot, int buf) {
    float* dst = sL + buf * TILE_FLOATS;
    if (TMA) {
      if (tid == 0) {
        mbar_expect_tx(&bars[buf], TILE_FLOATS * 4);
        if (KIND == IN_PYRAMID_TMA) {
          tma_load_4d(dst, &p.tmap[0], &bars[buf], tx0 - 4, ty0 - 4, 0, slot);
        } else {
          tma_load_3d(dst, &p.tmap[0], &bars[buf], tx0 - 4, ty0 - 4, (int)p.slot_frame[0][slot]);
          tma_load_3d(dst + LH * LW, &p.tmap[1], &bars[buf], tx0 - 4, ty0 - 4, (int)p.slot_frame[1][slot]);
        }
      }
    } else if (CHUNKED) {
      const float* b0 = reinterpret_cast<const float*>(p.slot[0][slot]);
      const float* b1 = reinterpret_cast<const float*>(p.slot[1][slot]);
#pragma unroll
      for (int i = 0; i < NLD; ++i) {
        if (ld_soff[i] >= 0) {
          const float* b = (ld_soff[i] >> 30) ? b1 : b0;
          const int g = ld_goff[i];
          cp_async16(dst + (ld_soff[i] & 0xFFFFFF), g >= 0 ? (const void*)(b + g) : (const void*)b, g >= 0 ? 16 : 0);
        }
      }
    } else {
      for (int item = tid; item < TILE_FLOATS; item += NT) {
        const int s = item / (LH * LW), rem = item % (LH * LW), r = rem / LW, c = rem % LW;
        const int y = ty0 - 4 + r, x = tx0 - 4 + c;
        float v = 0.0f;
        if (y >= 0 && y < h && x >= 0 && x < w) v = lum_generic(p, p.slot[s][slot], y, x, vmin, vmax);
        dst[item] = v;
      }
    }
  };
  // wait for the staged tile; level 0 with raw float input: display EOTF in place on this thread's chunks
  auto finish_load = [&](int buf, unsigned parity) {
    if (TMA) mbar_wait(&bars[buf], parity);
    else if (CHUNKED) cp_async_commit_wait_all();
    if (LEVEL0 && CHUNKED) {
      float* dst = sL + buf * TILE_FLOATS;
      switch (p.eotf) {  // uniform; one specialised conversion loop per EOTF
        case FVVDP_B200_EOTF_NONE: break;
        case FVVDP_B200_EOTF_SRGB: eotf_chunks<FVVDP_B200_EOTF_SRGB>(dst, ld_soff, ld_goff, p, vmin, vmax); break;
        case FVVDP_B200_EOTF_GAMMA: eotf_chunks<FVVDP_B200_EOTF_GAMMA>(dst, ld_soff, ld_goff, p, vmin, vmax); break;
        case FVVDP_B200_EOTF_PQ: eotf_chunks<FVVDP_B200_EOTF_PQ>(dst, ld_soff, ld_goff, p, vmin, vmax); break;
        case FVVDP_B200_EOTF_LINEAR: eotf_chunks<FVVDP_B200_EOTF_LINEAR>(dst, ld_soff, ld_goff, p, vmin, vmax); break;
        default: eotf_chunks<FVVDP_B200_EOTF_ABSOLUTE>(dst, ld_soff, ld_goff, p, vmin, vmax); break;
      }
      if (TMA) fence_proxy_async_smem();  // these generic-proxy writes precede the next TMA write into this buffer
    }
  };

  if (TMA) __syncthreads();  // barrier initialisation visible before the first wait
  issue_load(s_lo, 0);

  for (int s = s_lo; s < s_hi; ++s) {
    const int buf = (s - s_lo) & 1;
    const float* sLb = sL + buf * TILE_FLOATS;
    finish_load(buf, ((s - s_lo) >> 1) & 1);
    __syncthreads();  // (1) tile of slot s staged; every reader of the other buffer is done
    if (s + 1 < s_hi) issue_load(s + 1, buf ^ 1);

    // ---- reduce, rows: sV[st][a][c] = sum_k K[k] L[2j-2+k][c],  j = clamp(jy0-1+a)  (zero padding + edge terms) ----
    if (tid < 2 * LW) {  // one thread walks down one staged column
      const int st = tid >= LW ? 1 : 0, c = tid - st * LW;
      const float* col = sLb + st * LH * LW + c;
      float* out = sV + st * NH * LW + c;
      if (rows_interior) {  // no clamped coarse rows, no edge terms: sliding 5-row window
        float g0 = col[0], g1 = col[LW], g2 = col[2 * LW];
#pragma unroll
        for (int a = 0; a < NH; ++a) {
          const float g3 = col[(2 * a + 3) * LW], g4 = col[(2 * a + 4) * LW];
          out[a * LW] = fmaf(K0, g0 + g4, fmaf(K1, g1 + g3, K2 * g2));
          g0 = g2; g1 = g3; g2 = g4;
        }
      } else {
#pragma unroll
        for (int a = 0; a < NH; ++a) {
          const int jc = min(max(jy0 - 1 + a, 0), h2 - 1);  // expand clamps the coarse index
          const float* g = col + (2 * (jc - jy0) + 2) * LW;
          float v = fmaf(K0, g[0] + g[4 * LW], fmaf(K1, g[LW] + g[3 * LW], K2 * g[2 * LW]));
          if (jc == 0) v += K1 * g[2 * LW] + K0 * g[3 * LW];     // x[0], x[1]   (fvvdp_lpyr_dec.py:191)
          if (jc == h2 - 1) {
            const float* e = col + (h - 1 - ty0 + 4) * LW;       // x[h-1]
            v += (h & 1) ? (K3 * e[0] + K4 * e[-LW]) : K4 * e[0];  // (:192-195)
          }
          out[a * LW] = v;
        }
      }
    }
    __syncthreads();  // (2)
    // ---- reduce, columns -> ring slot s % FL (+ next level out) ----
    {
      float* ring_s = sNr + (s % FL) * (2 * NE);
      float* gout = (p.Pn != nullptr && s >= s_lo + ((blockIdx.z > 0) ? p.fl - 1 : 0)) ? p.Pn + (long long)s * p.Pn_slot_stride : nullptr;
#pragma unroll
      for (int i = 0; i < NCOL; ++i) {
        if (i < NCOL - 1 || cl_src[i] >= 0) {  // only the last round is partial
          const float* v = sV + (cl_src[i] & 0xFFFFFFF);
          float o = fmaf(K0, v[0] + v[4], fmaf(K1, v[1] + v[3], K2 * v[2]));
          if (!cols_interior) {
            if (cl_src[i] & (1 << 28)) o += K1 * v[2] + K0 * v[3];
            if (cl_src[i] & (2 << 28)) {
              const float* e = sV + ((cl_src[i] & 0xFFFFFFF) / LW) * LW + (w - 1 - tx0 + 4);  // y[w-1] of this row
              o += p.h_odd ? (K3 * e[0] + K4 * e[-1]) : K4 * e[0];  // keyed on the ROW count, fvvdp_lpyr_dec.py:202
            }
          }
          ring_s[cl_dst[i]] = o;
          if (gout != nullptr && cl_g[i] >= 0) gout[cl_g[i]] = o;
        }
      }
    }
    __syncthreads();  // (3)

    const bool emit = s >= f_lo + p.fl - 1;
    const int fi = s - (p.fl - 1);  // output frame
    // ---- this thread's pixels into the register ring; temporal filters of the coarse tiles and of the pixels.
    //      The ring position is a compile-time constant inside each case: no address arithmetic, no register moves.
    u64 R[NCH][2];
    switch (s % FL) {
#define FVVDP_CASE(J)                                          \
  case J:                                                      \
    ring_store<FL, (J) % FL>(ring, sLb, coff);                 \
    if (emit) {                                                \
      if (FL > 1) fir_coarse<FL, TC, (J) % FL>(sNr, sNc, p, tid); \
      fir_quad<FL, TC, (J) % FL>(ring, p, R);                  \
    }                                                          \
    break;
      FVVDP_CASE(0) FVVDP_CASE(1) FVVDP_CASE(2) FVVDP_CASE(3) FVVDP_CASE(4) FVVDP_CASE(5) FVVDP_CASE(6) FVVDP_CASE(7)
#undef FVVDP_CASE
    }
    if (!emit) continue;
    if (FL > 1) __syncthreads();  // (4) filtered coarse tiles visible

    // ---- expand the filtered coarse tile, contrast, CSF, masking, pooling: one 2x2 quad per thread ----
    float acc[2] = {0.0f, 0.0f};
    float Lb[4], lgL[4], fj[4];
    int cj[4];
    float lsf[TC][4];  // FOV: log2 S per temporal channel
    float Dsum[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    const u64 c01 = pk(0.1f, 0.1f), c08 = pk(0.8f, 0.8f), c05 = pk(0.5f, 0.5f);
#pragma unroll
    for (int cc = 0; cc < TC; ++cc) {
      // expand both streams at once: every value below is a (test, reference) pair
      const float* n = sNc + cc * (2 * NE) + noff;
      u64 ve[3], vo[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const u64 n0 = *reinterpret_cast<const u64*>(n + 2 * c), n1 = *reinterpret_cast<const u64*>(n + 2 * (NW + c)),
                  n2 = *reinterpret_cast<const u64*>(n + 2 * (2 * NW + c));
        ve[c] = ffma2(c08, n1, fmul2(c01, fadd2(n0, n2)));  // even row: taps 2K[0], 2K[2], 2K[4]
        vo[c] = fmul2(c05, fadd2(n1, n2));                  // odd row:  taps 2K[1], 2K[3]
      }
      u64 E[4];
      E[0] = ffma2(c08, ve[1], fmul2(c01, fadd2(ve[0], ve[2])));
      E[1] = fmul2(c05, fadd2(ve[1], ve[2]));
      E[2] = ffma2(c08, vo[1], fmul2(c01, fadd2(vo[0], vo[2])));
      E[3] = fmul2(c05, fadd2(vo[1], vo[2]));
      float B[2][4];  // band (G_l - E) of the test / reference channel
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const u64 rt = R[cc * 2 + 0][e >> 1], rr = R[cc * 2 + 1][e >> 1];
        B[0][e] = ((e & 1) ? hi_of(rt) : lo_of(rt)) - lo_of(E[e]);
        B[1][e] = ((e & 1) ? hi_of(rr) : lo_of(rr)) - hi_of(E[e]);
      }
      if (cc == 0) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          Lb[e] = fmaxf(hi_of(E[e]), 0.1f);  // L_bkg = expanded sustained reference (:264-266)
          lgL[e] = fast_log2(Lb[e]);
          const float yq = fminf(lgL[e], p.lg_y_hi);
          if (!FOV) {
            cj[e] = min((int)((yq - p.y0) * p.inv_dy), 30) * 8;  // L_bkg >= 0.1 lies above the first axis point: no lower clamp
            const float2 xi = *reinterpret_cast<const float2*>(sTab + cj[e]);
            fj[e] = (yq - xi.x) * xi.y;
          } else {
            const int x = qx + (e & 1), y = qy + (e >> 1);
            int jj, ii, kk;
            float fy, fr, fe;
            locate_direct(yq, p.ax.x[1], p.ax.inv[1], p.ax.x0[1], p.ax.inv_dx[1], jj, fy);
            const float vx = __ldg(p.vx + min(x, w - 1)), vy = __ldg(p.vy + min(y, h - 1));
            const float ex = vx - p.gaze[fi][0], ey = vy - p.gaze[fi][1];
            const float ecc = sqrtf(ex * ex + ey * ey);
            const float va = fminf(sqrtf(vx * vx + vy * vy), 89.9f) * 0.017453292519943295f;
            const float res_mag = p.res_k0 / (__cosf(va) * __cosf(va + p.res_delta_rad));
            const float rq = fast_log2(fminf(fmaxf(p.rho_band * res_mag, p.ax.lo[0]), p.ax.hi[0]));
            const float eq = sqrtf(fminf(fmaxf(ecc, p.ax.lo[2]), p.ax.hi[2]));
            locate_direct(rq, p.ax.x[0], p.ax.inv[0], p.ax.x0[0], p.ax.inv_dx[0], ii, fr);
            locate_direct(eq, p.ax.x[2], p.ax.inv[2], p.ax.x0[2], p.ax.inv_dx[2], kk, fe);
#pragma unroll
            for (int c2 = 0; c2 < TC; ++c2) {
              const float* v = p.lut3d + c2 * 32768 + (jj * 32 + ii) * 32 + kk;
              const float a00 = __ldg(v), a01 = __ldg(v + 32), a10 = __ldg(v + 1024), a11 = __ldg(v + 1056);
              const float b00 = __ldg(v + 1), b01 = __ldg(v + 33), b10 = __ldg(v + 1025), b11 = __ldg(v + 1057);
              const float lo = (a00 * (1.0f - fr) + a01 * fr) * (1.0f - fy) + (a10 * (1.0f - fr) + a11 * fr) * fy;
              const float hi = (b00 * (1.0f - fr) + b01 * fr) * (1.0f - fy) + (b10 * (1.0f - fr) + b11 * fr) * fy;
              lsf[c2][e] = lo * (1.0f - fe) + hi * fe + p.log2_sens_mul;
            }
          }
        }
      }
#pragma unroll
      for (int e = 0; e < 4; ++e) {
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
          const long long pofs = (long long)(qy + (e >> 1)) * w + qx + (e & 1);
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
          if (LEVEL0 && p.tapR) {
            const u64 rt = R[cc * 2 + 0][e >> 1], rr = R[cc * 2 + 1][e >> 1];
            p.tapR[((long long)fi * NCH + cc * 2 + 0) * plane + pofs] = (e & 1) ? hi_of(rt) : lo_of(rt);
            p.tapR[((long long)fi * NCH + cc * 2 + 1) * plane + pofs] = (e & 1) ? hi_of(rr) : lo_of(rr);
          }
          if (cc == TC - 1 && p.dmap) p.dmap[(long long)fi * plane + pofs] = Dsum[e] / p.band_mul;
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
}

template <int FL, int TC>
constexpr size_t band_smem_bytes() {
  return sizeof(float) * (size_t)(2 * TILE_FLOATS + 2 * NH * LW + FL * 2 * NE + (FL == 1 ? 0 : 2 * TC * NE) + 256 + MAXCHUNK * 2 * (NT / 32));
}

}  // namespace fused
}  // namespace fvvdp
