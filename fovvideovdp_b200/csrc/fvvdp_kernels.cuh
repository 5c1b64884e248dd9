// Device kernels of the B200-native FovVideoVDP core (sm_100a).  See DESIGN.md for the data layout and
// the roofline of each kernel.  Reference semantics are cited per function (pyfvvdp file:line).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>

#include "fvvdp_common.cuh"

namespace fvvdp {

// ------------------------------------------------------------------------------------------------
// K_front: display EOTF -> luminance, sliding temporal window, sustained + transient FIR
//   reference: video_source.py:180-208 (_get_frame), fvvdp_display_model.py:147-165 (EOTF),
//              fvvdp.py:258-300 (window + FIR), channel order [T_sust, R_sust, T_trans, R_trans]
// ------------------------------------------------------------------------------------------------
struct FrontParams {
  const void* slot[2][FVVDP_B200_MAX_SLOTS];  // [test|ref][window slot] frame base pointers (device)
  float wgt[2][FVVDP_B200_MAX_FILTER_LEN];    // [channel][window position], position 0 = OLDEST frame
  float* R;                                   // [n_frames][nch][H][W]; pairs output: [2 channels][pairs_slots][H][pitch], (test, ref) interleaved
  int pitch, pairs_slots;                     // pairs output only (the planes the fused band kernel stages by TMA)
  float wage[2][32];                          // front_pairs_kernel: weight of the frame of AGE a (0 = newest), 0 beyond the window
  int ring_phase;                             //   slot s sits at ring position (s + ring_phase) mod 32: the position follows the index in the clip
  uint32_t* flags;
  long long sC, sH, sW;                       // element strides of the input frames
  int H, W, n_frames, fl, nch, C, dtype, eotf;
  float Yscale, Y_black, Y_peak, gamma, L_min, L_max;
  float rgb2y[3];
};

__device__ __forceinline__ float eotf_apply(float v, const FrontParams& p, bool& oor) {
  switch (p.eotf) {
    case FVVDP_B200_EOTF_NONE:
      return v;
    case FVVDP_B200_EOTF_ABSOLUTE:  // fvvdp_display_model.py:206
      return fminf(fmaxf(v, p.L_min), p.L_max);
    case FVVDP_B200_EOTF_LINEAR:    // :163
      return fminf(fmaxf(v, 0.005f), p.Y_peak) + p.Y_black;
    default:
      break;
  }
  oor |= (v > 1.0f) | (v < 0.0f);   // :149-151 (warn + clamp)
  v = fminf(fmaxf(v, 0.0f), 1.0f);
  if (p.eotf == FVVDP_B200_EOTF_SRGB) {  // :17-19
    float lin = (v > 0.04045f) ? fast_pow((v + 0.055f) / 1.055f, 2.4f) : v / 12.92f;
    return p.Yscale * lin + p.Y_black;
  } else if (p.eotf == FVVDP_B200_EOTF_GAMMA) {
    return p.Yscale * fast_pow(v, p.gamma) + p.Y_black;
  } else {  // PQ :100-112,161
    const float n_inv = 1.0f / 0.15930175781250000f, m_inv = 1.0f / 78.843750000000000f;
    const float c1 = 0.83593750000000000f, c2 = 18.851562500000000f, c3 = 18.687500000000000f;
    float t = fast_pow(v, m_inv);
    float L = 10000.0f * fast_pow(fmaxf(t - c1, 0.0f) / (c2 - c3 * t), n_inv);
    return fminf(fmaxf(L, 0.005f), p.Y_peak) + p.Y_black;
  }
}

// luminance of PX consecutive pixels of one frame
template <int PX, bool CONTIG>
__device__ __forceinline__ void load_lum(const FrontParams& p, const void* base, int y, int x, float (&out)[PX], bool& oor) {
  if (CONTIG) {  // float32, 1 channel, unit column stride, 16-byte aligned rows
    const float* src = reinterpret_cast<const float*>(base) + (long long)y * p.sH + x;
    if (PX == 4) {
      float4 v = __ldg(reinterpret_cast<const float4*>(src));
      out[0] = v.x; out[1 % PX] = v.y; out[2 % PX] = v.z; out[3 % PX] = v.w;
    } else if (PX == 2) {
      float2 v = __ldg(reinterpret_cast<const float2*>(src));
      out[0] = v.x; out[1 % PX] = v.y;
    } else {
      out[0] = __ldg(src);
    }
#pragma unroll
    for (int i = 0; i < PX; ++i) out[i] = eotf_apply(out[i], p, oor);
  } else {
#pragma unroll
    for (int i = 0; i < PX; ++i) {
      long long off = (long long)y * p.sH + (long long)(x + i) * p.sW;
      if (p.C == 3) {
        float r = eotf_apply(load_sample(base, off, p.dtype), p, oor);
        float g = eotf_apply(load_sample(base, off + p.sC, p.dtype), p, oor);
        float b = eotf_apply(load_sample(base, off + 2 * p.sC, p.dtype), p, oor);
        out[i] = r * p.rgb2y[0] + g * p.rgb2y[1] + b * p.rgb2y[2];  // video_source.py:206
      } else {
        out[i] = eotf_apply(load_sample(base, off, p.dtype), p, oor);
      }
    }
  }
}

template <int PX>
__device__ __forceinline__ void store_px(float* dst, const float (&v)[PX]) {
  if (PX == 4) {
    *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1 % PX], v[2 % PX], v[3 % PX]);
  } else if (PX == 2) {
    *reinterpret_cast<float2*>(dst) = make_float2(v[0], v[1 % PX]);
  } else {
    dst[0] = v[0];
  }
}

// One thread owns PX consecutive pixels and walks the frames of the block keeping the last FL luminance
// samples of both streams in a register ring (position of slot s is (s + FL - fl) % FL, all indices static
// after unrolling).  Every input sample is read from HBM once per block and EOTF'd once.
template <int FL, int PX, bool CONTIG, bool PAIRS = false>
__global__ void __launch_bounds__(256) front_kernel(const __grid_constant__ FrontParams p) {
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long npx = (long long)p.H * p.W;
  const long long pix = gid * PX;
  if (pix >= npx) return;
  const int y = (int)(pix / p.W), x = (int)(pix % p.W);
  const int o = FL - p.fl;  // leading window positions that do not exist (weights are zero there)
  float win[2][FL][PX];
  bool oor = false;
#pragma unroll
  for (int k = 0; k < FL; ++k)
#pragma unroll
    for (int i = 0; i < PX; ++i) win[0][k][i] = win[1][k][i] = 0.0f;
  // pre-load window positions 0..FL-2 of output frame 0
#pragma unroll
  for (int k = 0; k < FL - 1; ++k) {
    const int s = k - o;
    if (s >= 0) {
      load_lum<PX, CONTIG>(p, p.slot[0][s], y, x, win[0][k], oor);
      load_lum<PX, CONTIG>(p, p.slot[1][s], y, x, win[1][k], oor);
    }
  }
  const long long plane = npx;
  for (int i0 = 0; i0 < p.n_frames; i0 += FL) {
#pragma unroll
    for (int j = 0; j < FL; ++j) {
      const int i = i0 + j;
      if (i < p.n_frames) {
        // newest frame of output i: window position FL-1 -> ring position (j + FL - 1) % FL, slot i + fl - 1
        const int rp = (j + FL - 1) % FL;
        load_lum<PX, CONTIG>(p, p.slot[0][i + p.fl - 1], y, x, win[0][rp], oor);
        load_lum<PX, CONTIG>(p, p.slot[1][i + p.fl - 1], y, x, win[1][rp], oor);
        float* dst = p.R + ((long long)i * p.nch) * plane + pix;
        const int ncc = p.nch >> 1;
        for (int cc = 0; cc < ncc; ++cc) {
          float at[PX], ar[PX];
#pragma unroll
          for (int e = 0; e < PX; ++e) at[e] = ar[e] = 0.0f;
#pragma unroll
          for (int k = 0; k < FL; ++k) {
            const float w = p.wgt[cc][k];
#pragma unroll
            for (int e = 0; e < PX; ++e) {
              at[e] = fmaf(win[0][(j + k) % FL][e], w, at[e]);
              ar[e] = fmaf(win[1][(j + k) % FL][e], w, ar[e]);
            }
          }
          if (PAIRS) {  // (test, reference) pairs in the pyramid layout, one plane per temporal channel and frame
            float* o = p.R + (((long long)cc * p.pairs_slots + i) * p.H + y) * p.pitch + 2 * x;
#pragma unroll
            for (int e = 0; e < PX; ++e) *reinterpret_cast<float2*>(o + 2 * e) = make_float2(at[e], ar[e]);
          } else {
            store_px<PX>(dst + (cc * 2 + 0) * plane, at);
            store_px<PX>(dst + (cc * 2 + 1) * plane, ar);
          }
        }
      }
    }
  }
  if (oor && p.flags) atomicOr(p.flags, 1u);
}

// Temporal filters for windows of 17..32 taps, (test, reference) pairs: one thread owns one pixel of both streams and walks the
// slots of the block with the last 31 luminance pairs in a register ring (62 registers; a window of fl taps needs fl - 1 stored
// frames: the newest frame's sustained tap is 1e-36 of the sum and the oldest frame's transient tap is exactly 0, fvvdp.py:609-626).
// The loop over the slots is unrolled by the ring length, so the ring position of a step is a compile-time constant and every
// weight a uniform operand of its FFMA2; a frame sits at position (index in the clip) mod 31, so the summation order does not
// depend on how the clip is cut into blocks.  Every input sample is read and converted once; the output is the two planes per frame
// (sustained, transient) the two-channel band kernel stages by TMA.
__device__ __forceinline__ unsigned long long fp_pk(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ unsigned long long fp_ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long fp_fadd2(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// The history of a pixel is a SHIFT REGISTER that moves once per FOUR steps: inside the four-times unrolled step the frame of age a
// sits at hist[a + 3 - d] (d = step within the group), so every filter weight is an operand at a static constant-bank address
// -- no loads, four code versions -- and the history is shifted by four positions (70 register moves) once per group.  The
// summation order goes by age and does not depend on how the clip is cut into blocks.  Measured alternatives (64-frame 4K clip,
// this kernel alone): the walk unrolled over 32 ring positions with uniform-register weights 12-15 ms (95-180 kB of code:
// instruction-cache bound), rotating weights from register-indexed constant loads 13.9 ms or from shared-memory broadcasts 14.7 ms
// (64 LDC / 128 shared-memory wavefronts per step and warp: load-issue bound), a shift per step 12.9 ms (62 moves per step, spills).
__device__ __forceinline__ float eotf_front(float v, const FrontParams& p, bool& oor) {
  if (p.eotf == FVVDP_B200_EOTF_SRGB) {  // the hot case without divisions (fvvdp_display_model.py:17-19)
    oor |= (v > 1.0f) | (v < 0.0f);
    const float t = __saturatef(fmaf(v, 1.0f / 1.055f, 0.055f / 1.055f));
    const float lin = (v > 0.04045f) ? fast_exp2(2.4f * fast_log2(t)) : __saturatef(v * (1.0f / 12.92f));
    return fmaf(p.Yscale, lin, p.Y_black);
  }
  return eotf_apply(v, p, oor);
}
template <bool CONTIG, int NA>   // NA = ages kept (0 = newest): 32, or 16 for windows of up to 16 taps
__global__ void __launch_bounds__(256, 2) front_pairs_kernel(const __grid_constant__ FrontParams p) {
  constexpr int PF = 4;   // steps per group = slots whose samples are in flight ahead of the one being filtered
  const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= (long long)p.H * p.W) return;
  const int y = (int)(pix / p.W), x = (int)(pix % p.W);
  float ht[NA + PF - 1], hr[NA + PF - 1];  // luminance history of the test / reference stream
#pragma unroll
  for (int k = 0; k < NA + PF - 1; ++k) ht[k] = hr[k] = 0.0f;
  bool oor = false;
  const int n_slots = p.n_frames + p.fl - 1;
  float* out = p.R + (long long)y * p.pitch + 2 * x;
  const long long frame_stride = (long long)p.H * p.pitch, chan_stride = (long long)p.pairs_slots * frame_stride;
  const long long in_off = (long long)y * p.sH + x;
  float pre[PF][2];  // CONTIG: raw samples of the slots ahead
  if (CONTIG) {
#pragma unroll
    for (int d = 0; d < PF; ++d) {
      pre[d][0] = d < n_slots ? __ldg(reinterpret_cast<const float*>(p.slot[0][d]) + in_off) : 0.0f;
      pre[d][1] = d < n_slots ? __ldg(reinterpret_cast<const float*>(p.slot[1][d]) + in_off) : 0.0f;
    }
  }
  for (int s4 = 0; s4 < n_slots; s4 += PF) {
    // make room for the four frames of this group: age a moves from hist[a] ... to hist[a + 4] (what falls off the end is older than 31)
#pragma unroll
    for (int k = NA + PF - 2; k >= PF; --k) { ht[k] = ht[k - PF]; hr[k] = hr[k - PF]; }
#pragma unroll
    for (int d = 0; d < PF; ++d) {
      const int s = s4 + d;
      if (s < n_slots) {
        float a[1], b[1];
        if (CONTIG) {
          a[0] = eotf_front(pre[d][0], p, oor);
          b[0] = eotf_front(pre[d][1], p, oor);
          if (s + PF < n_slots) {
            pre[d][0] = __ldg(reinterpret_cast<const float*>(p.slot[0][s + PF]) + in_off);
            pre[d][1] = __ldg(reinterpret_cast<const float*>(p.slot[1][s + PF]) + in_off);
          }
        } else {
          load_lum<1, false>(p, p.slot[0][s], y, x, a, oor);
          load_lum<1, false>(p, p.slot[1][s], y, x, b, oor);
        }
        ht[PF - 1 - d] = a[0];   // age 0 of step d; age k sits at hist[k + PF - 1 - d]
        hr[PF - 1 - d] = b[0];
        if (s >= p.fl - 1) {
          const int i = s - (p.fl - 1);
#pragma unroll
          for (int cc = 0; cc < 2; ++cc) {
            float st[2] = {0.0f, 0.0f}, sr[2] = {0.0f, 0.0f};  // two partial sums per stream
#pragma unroll
            for (int k = 0; k < NA; ++k) {
              st[k & 1] = fmaf(ht[k + PF - 1 - d], p.wage[cc][k], st[k & 1]);
              sr[k & 1] = fmaf(hr[k + PF - 1 - d], p.wage[cc][k], sr[k & 1]);
            }
            *reinterpret_cast<float2*>(out + cc * chan_stride + i * frame_stride) = make_float2(st[0] + st[1], sr[0] + sr[1]);
          }
        }
      }
    }
  }
  if (oor && p.flags) atomicOr(p.flags, 1u);
}

// ------------------------------------------------------------------------------------------------
// K_level: one pyramid level, fully fused
//   gausspyr_reduce (fvvdp_lpyr_dec.py:183-207) -> G_{l+1} tile (+1 halo) in shared memory (+ written out)
//   gausspyr_expand (:219-235,126-142), band = G_l - E, L_bkg = max(E[ref_sust],0.1), contrast (:259-269)
//   CSF look-up (fvvdp.py:520-537, interp.py:11-59), masking (:574-596), sum of D^beta (:467,598-607)
// ------------------------------------------------------------------------------------------------
constexpr int TH = 32, TW = 64, HALO = 4;
constexpr int SGH = TH + 2 * HALO, SGW = TW + 2 * HALO;  // G_l tile with halo: 40 x 72
constexpr int NH = TH / 2 + 2, NW = TW / 2 + 2;          // G_{l+1} tile with 1-px halo: 18 x 34
constexpr int LEVEL_THREADS = 256;

struct LevelParams {
  const float* G;     // [F][NCH][h][w]
  float* Gn;          // [F][NCH][h2][w2] or nullptr for the last band
  float* partial;     // [F][2][ntiles]
  int h, w, h2, w2, quirk, ntiles;
  float band_mul;
  float rho_band;
  // CSF
  CsfAxes ax;
  const float* csf1d;  // [2][32] pre-blended log2(S * sens_mul) over log2 Y for this band (non-foveated)
  const float* lut3d;  // [2][32(Y)][32(rho)][32(ecc)] log2 S
  float log2_sens_mul;
  // masking
  float mask_p, mask_q[2], mask_c_mul, beta, w_transient;
  // foveation
  const float* vx;     // [w] horizontal view direction of the band's pixel columns (deg)
  const float* vy;     // [h]
  const float* vmap;   // custom display geometry: view direction per pixel [2][h][w] (deg), else nullptr
  const float* rqmap;  // custom display geometry: log2(clamp(rho_band * res_mag)) per pixel [h][w]
  float res_k0, res_delta_rad;  // res_mag = res_k0 / (cos(a) cos(a+delta))
  float gaze[FVVDP_B200_MAX_BLOCK_FRAMES][2];
  // optional outputs
  float* tapC;  // [F][NCH][h][w]
  float* tapL;  // [F][h][w]
  float* tapS;  // [F][TC][h][w]
  float* tapD;  // [F][TC][h][w]
  float* dmap;  // [F][h][w]
};

__device__ __forceinline__ int mirror_idx(int i, int n) {
  if (i < 0) i = -1 - i;
  if (i >= n) i = 2 * n - 1 - i;
  return min(max(i, 0), n - 1);
}

// bucketize + fraction exactly as get_interpolants_v1 (interp.py:11-20): j1 = first index with x[j1] >= q
__device__ __forceinline__ void locate(float q, const float* __restrict__ x, const float* __restrict__ inv, float x0, float inv_dx,
                                       int& j0, int& j1, float& f) {
  int j = (int)ceilf((q - x0) * inv_dx);
  j = min(max(j, 0), 31);
  if (j > 0 && x[j - 1] >= q) --j;
  else if (j < 31 && x[j] < q) ++j;
  j1 = j;
  j0 = max(j - 1, 0);
  f = (j1 == j0) ? 0.0f : fmaxf((q - x[j0]) * inv[j1], 0.0f);
}

template <int NCH, bool FOV>
__global__ void __launch_bounds__(LEVEL_THREADS) level_kernel(const __grid_constant__ LevelParams p) {
  constexpr int TC = NCH / 2;
  extern __shared__ float smem[];
  float* sG = smem;                        // [NCH][SGH][SGW]
  float* sV = sG + NCH * SGH * SGW;        // [NCH][NH][SGW]
  float* sN = sV + NCH * NH * SGW;         // [NCH][NH][NW]
  float* sAx = sN + NCH * NH * NW;         // Y_log[32], invY[32], tab[2][32]
  __shared__ float sRed[2][LEVEL_THREADS / 32];

  const int tid = threadIdx.x;
  const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
  const int f = blockIdx.z;
  const int h = p.h, w = p.w, h2 = p.h2, w2 = p.w2;
  const int jx0 = tx0 >> 1, jy0 = ty0 >> 1;
  const long long plane = (long long)h * w;
  const float* Gf = p.G + (long long)f * NCH * plane;

  if (tid < 32) {
    sAx[tid] = p.ax.x[1][tid];
    sAx[32 + tid] = p.ax.inv[1][tid];
    if (!FOV) {
      sAx[64 + tid] = p.csf1d[tid];
      sAx[96 + tid] = p.csf1d[32 + tid];
    }
  }

  // ---- stage the G_l tile (mirror-extended at the image border) ----
  const bool interior = (ty0 - HALO >= 0) && (ty0 + TH + HALO <= h) && (tx0 - HALO >= 0) && (tx0 + TW + HALO <= w) && ((w & 3) == 0);
  if (interior) {
    constexpr int V4 = SGW / 4;
    for (int idx = tid; idx < NCH * SGH * V4; idx += LEVEL_THREADS) {
      const int c4 = idx % V4, r = (idx / V4) % SGH, ch = idx / (V4 * SGH);
      const float4 v = __ldg(reinterpret_cast<const float4*>(Gf + ch * plane + (long long)(ty0 - HALO + r) * w + (tx0 - HALO)) + c4);
      *reinterpret_cast<float4*>(sG + (ch * SGH + r) * SGW + c4 * 4) = v;
    }
  } else {
    for (int idx = tid; idx < NCH * SGH * SGW; idx += LEVEL_THREADS) {
      const int c = idx % SGW, r = (idx / SGW) % SGH, ch = idx / (SGW * SGH);
      const int gy = mirror_idx(ty0 - HALO + r, h), gx = mirror_idx(tx0 - HALO + c, w);
      sG[idx] = __ldg(Gf + ch * plane + (long long)gy * w + gx);
    }
  }
  __syncthreads();

  const float K0 = 0.05f, K1 = 0.25f, K2 = 0.4f, K3 = 0.25f, K4 = 0.05f;
  // ---- reduce, rows: sV[ch][a][c] = sum_k K[k] G[2j-2+k][c], j = clamp(jy0-1+a) ----
  for (int idx = tid; idx < NCH * NH * SGW; idx += LEVEL_THREADS) {
    const int c = idx % SGW, a = (idx / SGW) % NH, ch = idx / (SGW * NH);
    const int jc = min(max(jy0 - 1 + a, 0), h2 - 1);
    const float* g = sG + (ch * SGH + 2 * (jc - jy0) + 2) * SGW + c;
    sV[idx] = K0 * g[0] + K1 * g[SGW] + K2 * g[2 * SGW] + K3 * g[3 * SGW] + K4 * g[4 * SGW];
  }
  __syncthreads();
  // ---- reduce, columns (with the reference's row-parity quirk on the last column, fvvdp_lpyr_dec.py:202) ----
  for (int idx = tid; idx < NCH * NH * NW; idx += LEVEL_THREADS) {
    const int b = idx % NW, a = (idx / NW) % NH, ch = idx / (NW * NH);
    const int ic = min(max(jx0 - 1 + b, 0), w2 - 1);
    const float* v = sV + (ch * NH + a) * SGW + 2 * (ic - jx0) + 2;
    float o;
    if (p.quirk != 0 && ic == w2 - 1) {
      if (p.quirk == 1) o = (K0 * v[0] + K1 * v[1] + K2 * v[2] + K3 * v[3]) + (v[3] * K3 + v[2] * K4);  // W even, H odd
      else              o = (K0 * v[0] + K1 * v[1] + K2 * v[2]) + v[2] * K4;                            // W odd, H even
    } else {
      o = K0 * v[0] + K1 * v[1] + K2 * v[2] + K3 * v[3] + K4 * v[4];
    }
    sN[idx] = o;
    if (p.Gn != nullptr && a >= 1 && a <= TH / 2 && b >= 1 && b <= TW / 2) {
      const int j = jy0 - 1 + a, i = jx0 - 1 + b;
      if (j < h2 && i < w2) p.Gn[((long long)(f * NCH + ch) * h2 + j) * w2 + i] = o;
    }
  }
  __syncthreads();

  // ---- expand + contrast + CSF + masking + pooling; one thread per 2x2 output quad ----
  float acc[2] = {0.0f, 0.0f};
  const float* sY = sAx;
  const float* sYinv = sAx + 32;
  const float log2_dmax = 13.287712379549449f;  // log2(1e4), clamp of fvvdp.py:595
  for (int q = tid; q < (TH / 2) * (TW / 2); q += LEVEL_THREADS) {
    const int qa = q / (TW / 2), qb = q % (TW / 2);
    const int y0 = ty0 + 2 * qa, x0 = tx0 + 2 * qb;
    if (y0 >= h || x0 >= w) continue;
    float E[NCH][2][2];
#pragma unroll
    for (int ch = 0; ch < NCH; ++ch) {
      const float* n = sN + (ch * NH + qa) * NW + qb;
      float ve[3], vo[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float n0 = n[c], n1 = n[NW + c], n2 = n[2 * NW + c];
        ve[c] = 0.1f * n0 + 0.8f * n1 + 0.1f * n2;  // even row: taps 2K[0],2K[2],2K[4]
        vo[c] = 0.5f * n1 + 0.5f * n2;              // odd row:  taps 2K[1],2K[3]
      }
      E[ch][0][0] = 0.1f * ve[0] + 0.8f * ve[1] + 0.1f * ve[2];
      E[ch][0][1] = 0.5f * ve[1] + 0.5f * ve[2];
      E[ch][1][0] = 0.1f * vo[0] + 0.8f * vo[1] + 0.1f * vo[2];
      E[ch][1][1] = 0.5f * vo[1] + 0.5f * vo[2];
    }
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        const int y = y0 + dy, x = x0 + dx;
        if (y >= h || x >= w) continue;
        const float Lb = fmaxf(E[1][dy][dx], 0.1f);
        const float invL = fast_rcp(Lb);
        float C[NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          const float g = sG[(ch * SGH + 2 * qa + dy + HALO) * SGW + 2 * qb + dx + HALO];
          C[ch] = fminf((g - E[ch][dy][dx]) * invL, 1000.0f) * p.band_mul;
        }
        const long long pofs = (long long)y * w + x;
        // CSF query along Y (shared by both temporal channels)
        const float yq = fast_log2(fminf(fmaxf(Lb, p.ax.lo[1]), p.ax.hi[1]));
        int j0, j1;
        float fj;
        locate(yq, sY, sYinv, p.ax.x0[1], p.ax.inv_dx[1], j0, j1, fj);
        int i0 = 0, i1 = 0, k0 = 0, k1 = 0;
        float fi = 0.0f, fk = 0.0f;
        if (FOV) {
          float vx, vy, rq;
          if (p.vmap != nullptr) {  // maps computed by a fvvdp_display_geometry subclass
            vx = __ldg(p.vmap + pofs); vy = __ldg(p.vmap + plane + pofs); rq = __ldg(p.rqmap + pofs);
          } else {
            vx = __ldg(p.vx + x); vy = __ldg(p.vy + y);
            const float va = fminf(sqrtf(vx * vx + vy * vy), 89.9f) * 0.017453292519943295f;
            const float res_mag = p.res_k0 / (__cosf(va) * __cosf(va + p.res_delta_rad));
            rq = fast_log2(fminf(fmaxf(p.rho_band * res_mag, p.ax.lo[0]), p.ax.hi[0]));
          }
          const float ex = vx - p.gaze[f][0], ey = vy - p.gaze[f][1];
          const float ecc = sqrtf(ex * ex + ey * ey);
          const float eq = sqrtf(fminf(fmaxf(ecc, p.ax.lo[2]), p.ax.hi[2]));
          locate(rq, p.ax.x[0], p.ax.inv[0], p.ax.x0[0], p.ax.inv_dx[0], i0, i1, fi);
          locate(eq, p.ax.x[2], p.ax.inv[2], p.ax.x0[2], p.ax.inv_dx[2], k0, k1, fk);
        }
        float Dsum = 0.0f;
#pragma unroll
        for (int cc = 0; cc < TC; ++cc) {
          float ls;
          if (FOV) {
            const float* v = p.lut3d + cc * 32768;
            const float a00 = __ldg(v + (j0 * 32 + i0) * 32 + k0), a01 = __ldg(v + (j0 * 32 + i1) * 32 + k0);
            const float a10 = __ldg(v + (j1 * 32 + i0) * 32 + k0), a11 = __ldg(v + (j1 * 32 + i1) * 32 + k0);
            const float b00 = __ldg(v + (j0 * 32 + i0) * 32 + k1), b01 = __ldg(v + (j0 * 32 + i1) * 32 + k1);
            const float b10 = __ldg(v + (j1 * 32 + i0) * 32 + k1), b11 = __ldg(v + (j1 * 32 + i1) * 32 + k1);
            const float lo = (a00 * (1.0f - fi) + a01 * fi) * (1.0f - fj) + (a10 * (1.0f - fi) + a11 * fi) * fj;
            const float hi = (b00 * (1.0f - fi) + b01 * fi) * (1.0f - fj) + (b10 * (1.0f - fi) + b11 * fi) * fj;
            ls = lo * (1.0f - fk) + hi * fk + p.log2_sens_mul;
          } else {
            const float* tab = sAx + 64 + cc * 32;
            ls = tab[j0] * (1.0f - fj) + tab[j1] * fj;
          }
          const float S = fast_exp2(ls);
          const float Tn = C[cc * 2 + 0] * S, Rn = C[cc * 2 + 1] * S;
          const float d = fabsf(Tn - Rn);
          const float M = fminf(fabsf(Tn), fabsf(Rn)) * p.mask_c_mul;
          const float Mq = fast_exp2(p.mask_q[cc] * fast_log2(M));
          float lD = p.mask_p * fast_log2(d) - fast_log2(1.0f + Mq);
          lD = fminf(lD, log2_dmax);
          acc[cc] += fast_exp2(p.beta * lD);
          if (p.tapS) p.tapS[((long long)(f * TC + cc)) * plane + pofs] = S;
          if (p.tapD || p.dmap) {
            const float D = fast_exp2(lD);
            if (p.tapD) p.tapD[((long long)(f * TC + cc)) * plane + pofs] = D;
            Dsum += (cc == 0 ? 1.0f : p.w_transient) * D;
          }
        }
        if (p.dmap) p.dmap[(long long)f * plane + pofs] = Dsum / p.band_mul;
        if (p.tapL) p.tapL[(long long)f * plane + pofs] = Lb;
        if (p.tapC) {
#pragma unroll
          for (int ch = 0; ch < NCH; ++ch) p.tapC[((long long)(f * NCH + ch)) * plane + pofs] = C[ch];
        }
      }
    }
  }
  // ---- block reduction of sum D^beta (warp shuffles, then one partial per tile: deterministic) ----
#pragma unroll
  for (int cc = 0; cc < 2; ++cc) {
    float v = acc[cc];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((tid & 31) == 0) sRed[cc][tid >> 5] = v;
  }
  __syncthreads();
  if (tid < 2) {
    float v = 0.0f;
#pragma unroll
    for (int i = 0; i < LEVEL_THREADS / 32; ++i) v += sRed[tid][i];
    const int tile = blockIdx.y * gridDim.x + blockIdx.x;
    p.partial[((long long)f * 2 + tid) * p.ntiles + tile] = v;
  }
}

constexpr size_t level_smem_bytes(int nch) { return sizeof(float) * (size_t)(nch * SGH * SGW + nch * NH * SGW + nch * NH * NW + 128); }

// ------------------------------------------------------------------------------------------------
// K_final: Q[bb,cc,f] = (sum_tiles partial / Npix)^(1/beta)      (lp_norm, fvvdp.py:598-607)
// ------------------------------------------------------------------------------------------------
struct FinalParams {
  const float* partial[FVVDP_B200_MAX_LEVELS];
  int ntiles[FVVDP_B200_MAX_LEVELS];
  double npix[FVVDP_B200_MAX_LEVELS];
  float* q_out;
  long long q_stride, q_col0;
  int n_bands, n_frames, temp_ch;
  double inv_beta;
};

__global__ void __launch_bounds__(256) final_kernel(const __grid_constant__ FinalParams p) {
  const int cc = blockIdx.x & 1, bb = (blockIdx.x >> 1) % p.n_bands, f = (blockIdx.x >> 1) / p.n_bands;
  __shared__ double sd[8];
  double s = 0.0;
  if (cc < p.temp_ch) {
    const int n = p.ntiles[bb];
    const float* src = p.partial[bb] + ((long long)f * 2 + cc) * n;
    if ((n & 3) == 0) {  // level 0 holds 16 per-warp partials per tile (65 280 floats at 4K): 128-bit loads, four in flight
      const float4* s4 = reinterpret_cast<const float4*>(src);
#pragma unroll 4
      for (int i = threadIdx.x; i < (n >> 2); i += 256) {
        const float4 v = __ldg(s4 + i);
        s += ((double)v.x + (double)v.y) + ((double)v.z + (double)v.w);
      }
    } else {
      for (int i = threadIdx.x; i < n; i += 256) s += (double)src[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    const double t = ((sd[0] + sd[1]) + (sd[2] + sd[3])) + ((sd[4] + sd[5]) + (sd[6] + sd[7]));
    const double q = (cc < p.temp_ch) ? pow(t / p.npix[bb], p.inv_beta) : 0.0;
    p.q_out[((long long)bb * 2 + cc) * p.q_stride + p.q_col0 + f] = (float)q;
  }
}

// ------------------------------------------------------------------------------------------------
// K_pool: do_pooling_and_jods (fvvdp.py:337-357).  One block; fp64 accumulation (a few thousand terms).
//   Q_sc[cc,f] = (sum_bb |w_cc Q[bb,cc,f]|^b_sch)^(1/b_sch);  Q_tc[f] = (sum_cc Q_sc^b_tch)^(1/b_tch)
//   Q = (sum_f Q_tc^b_t / N)^(1/b_t);  JOD = 10 + sign(a) (|a|^(1/b) Q)^b,  b = 10^log_jod_exp
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pool_kernel(const float* __restrict__ q, int n_bands, long long n_frames, long long q_stride,
                                                   const fvvdp_b200_pool_params p, float* __restrict__ out) {
  __shared__ double sd[8];
  double acc = 0.0;
  for (long long f = threadIdx.x; f < n_frames; f += 256) {
    double tc = 0.0;
    for (int cc = 0; cc < 2; ++cc) {
      const double w = cc == 0 ? 1.0 : (double)p.w_transient;
      double sc = 0.0;
      for (int bb = 0; bb < n_bands; ++bb) {
        const double v = fabs((double)(q[((long long)bb * 2 + cc) * q_stride + f] * (float)w));
        sc += (p.beta_sch == 1.0f) ? v : pow(v, (double)p.beta_sch);
      }
      if (p.beta_sch != 1.0f) sc = pow(sc, 1.0 / (double)p.beta_sch);
      tc += sc > 0.0 ? pow(sc, (double)p.beta_tch) : 0.0;
    }
    tc = tc > 0.0 ? pow(tc, 1.0 / (double)p.beta_tch) : 0.0;
    acc += (p.beta_t == 1.0f) ? tc : pow(tc, (double)p.beta_t);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sd[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sd[i];
    double Q = t / (double)n_frames;
    if (p.beta_t != 1.0f) Q = pow(Q, 1.0 / (double)p.beta_t);
    const double b = pow(10.0, (double)p.log_jod_exp);
    const double a = (double)p.jod_a;
    const double sgn = a < 0.0 ? -1.0 : 1.0;
    out[0] = (float)(sgn * pow(pow(fabs(a), 1.0 / b) * Q, b) + 10.0);
    out[1] = (float)Q;
  }
}

// ------------------------------------------------------------------------------------------------
// K_recon: heat-map pyramid reconstruction  img_l = expand(img_{l+1}) + band_l   (fvvdp_lpyr_dec.py:94-101)
// and final |jod_a| * img^beta_jod -> fp16 (fvvdp.py:470-473)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) recon_kernel(const float* __restrict__ coarse, int h2, int w2, const float* __restrict__ band,
                                                    float* __restrict__ out, __half* __restrict__ out16, int h, int w, float beta_jod,
                                                    float jod_a_abs) {
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= w || y >= h) return;
  float e = 0.0f;
  if (coarse != nullptr) {
    // expand: rows first, then columns; z[m] = x[clamp(m/2-1)] on even m (see oracle/_expand_axis)
    float col[3];
    const int cx = x >> 1, cy = y >> 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const int xi = min(max(cx - 1 + c, 0), w2 - 1);
      const float n0 = coarse[(long long)min(max(cy - 1, 0), h2 - 1) * w2 + xi];
      const float n1 = coarse[(long long)min(cy, h2 - 1) * w2 + xi];
      const float n2 = coarse[(long long)min(cy + 1, h2 - 1) * w2 + xi];
      col[c] = (y & 1) ? (0.5f * n1 + 0.5f * n2) : (0.1f * n0 + 0.8f * n1 + 0.1f * n2);
    }
    e = (x & 1) ? (0.5f * col[1] + 0.5f * col[2]) : (0.1f * col[0] + 0.8f * col[1] + 0.1f * col[2]);
  }
  const float v = e + band[(long long)y * w + x];
  if (out16 != nullptr) out16[(long long)y * w + x] = __float2half(powf(v, beta_jod) * jod_a_abs);
  else out[(long long)y * w + x] = v;
}

// ------------------------------------------------------------------------------------------------
// Heat-map visualisation (visualize_diff_map.py:9-107): colour map of the difference map over the tone-mapped
// context frame (the sustained test frame, fvvdp.py:475).  Four small kernels per frame:
//   vis_range   min over the positive values and max of the context luminance
//   vis_hist    1024-bin histogram of its log (torch.histc semantics)
//   vis_curve   tone curve: cumulative sum of the cube root of the normalised histogram (vis_tonemap :26-50)
//   vis_apply   dmap = |jod_a| recon^beta_jod, colour-map look-up, times the tone-mapped context, clip, fp16
// ------------------------------------------------------------------------------------------------
struct VisWork {            // device scratch, one per context
  unsigned int range[2];    // bit patterns of (min positive y, max y)
  unsigned int hist[1024];
  float curve[1024];        // v of vis_tonemap
  float b_min, b_max, clampval;
  int linear;               // b_max - b_min < dr: no tone mapping
};
constexpr float VIS_DR = 0.6f;

__global__ void vis_reset_kernel(VisWork* ws) {
  const int i = threadIdx.x;
  ws->hist[i] = 0u;
  if (i == 0) { ws->range[0] = 0x7f800000u; ws->range[1] = 0u; }
}

__global__ void __launch_bounds__(256) vis_range_kernel(const float* __restrict__ y, long long n, VisWork* ws) {
  unsigned int lo = 0x7f800000u, hi = 0u;  // positive floats order like their bit patterns
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float v = __ldg(y + i);
    if (v > 0.0f) { lo = min(lo, __float_as_uint(v)); hi = max(hi, __float_as_uint(v)); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = min(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = max(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) { atomicMin(&ws->range[0], lo); atomicMax(&ws->range[1], hi); }
}

// log-luminance of the context frame (log_luminance, visualize_diff_map.py:20-23)
__device__ __forceinline__ float vis_log_lum(float y, float clampval) { return logf(fmaxf(y, clampval)); }

__global__ void __launch_bounds__(256) vis_hist_kernel(const float* __restrict__ y, long long n, VisWork* ws) {
  __shared__ unsigned int sh[1024];
  for (int i = threadIdx.x; i < 1024; i += 256) sh[i] = 0u;
  __syncthreads();
  const float clampval = __uint_as_float(ws->range[0]);
  const float b_min = logf(clampval), b_max = logf(fmaxf(__uint_as_float(ws->range[1]), clampval));
  const float range = b_max - b_min;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    const float b = vis_log_lum(__ldg(y + i), clampval);
    int pos = (int)((b - b_min) / range * 1024.0f);  // torch.histc; the maximum falls into the last bin
    pos = min(max(pos, 0), 1023);
    atomicAdd(&sh[pos], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 1024; i += 256)
    if (sh[i]) atomicAdd(&ws->hist[i], sh[i]);
}

__global__ void __launch_bounds__(1024) vis_curve_kernel(VisWork* ws, float n_pix) {
  __shared__ float sc[1024];
  __shared__ float total;
  const int i = threadIdx.x;
  const float clampval = __uint_as_float(ws->range[0]);
  const float b_min = logf(clampval), b_max = logf(fmaxf(__uint_as_float(ws->range[1]), clampval));
  const float pw = cbrtf((float)ws->hist[i] / n_pix);  // b_p^(1/t), t = 3
  sc[i] = pw;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan
    const float add = i >= o ? sc[i - o] : 0.0f;
    __syncthreads();
    sc[i] += add;
    __syncthreads();
  }
  if (i == 1023) total = sc[1023];
  __syncthreads();
  ws->curve[i] = sc[i] / total * VIS_DR + (1.0f - VIS_DR) / 2.0f;
  if (i == 0) {
    ws->b_min = b_min; ws->b_max = b_max; ws->clampval = clampval;
    ws->linear = (b_max - b_min < VIS_DR) ? 1 : 0;
  }
}

struct VisColorMap {  // colour-map nodes already divided by (their luminance + 1e-4), visualize_diff_map.py:96-98
  int n;
  float in[5];
  float ch[5][3];
};

// torch.linspace(b_min, b_max, 1024)[i] in float32
__device__ __forceinline__ float vis_scale(int i, float b_min, float b_max, float step) {
  return i < 512 ? b_min + step * (float)i : b_max - step * (float)(1023 - i);
}

__global__ void __launch_bounds__(256) vis_apply_kernel(const float* __restrict__ recon, const float* __restrict__ y, long long n,
                                                        const VisWork* __restrict__ ws, const __grid_constant__ VisColorMap cm, float beta_jod, float jod_a_abs,
                                                        __half* __restrict__ out) {
  const long long i = blockIdx.x * 256ll + threadIdx.x;
  if (i >= n) return;
  const float b_min = ws->b_min, b_max = ws->b_max;
  const float d = fminf(fmaxf(powf(recon[i], beta_jod) * jod_a_abs, 0.0f), 1.0f);
  const float b = vis_log_lum(__ldg(y + i), ws->clampval);
  float tmo;
  if (ws->linear) {
    tmo = (b - b_min) / (b_max - b_min + 1e-3f) * VIS_DR + (1.0f - VIS_DR) / 2.0f;
  } else {
    // interp1(b_scale, v, b) with get_interpolants_v1 (interp.py:11-20): imax = first index with b_scale[imax] >= b
    const float step = (b_max - b_min) / 1023.0f;
    int j = min(max((int)ceilf((b - b_min) / step), 0), 1023);
    while (j > 0 && vis_scale(j - 1, b_min, b_max, step) >= b) --j;
    while (j < 1023 && vis_scale(j, b_min, b_max, step) < b) ++j;
    const int j0 = max(j - 1, 0);
    const float x0 = vis_scale(j0, b_min, b_max, step), x1 = vis_scale(j, b_min, b_max, step);
    const float f = (j == j0) ? 0.0f : fmaxf((b - x0) / (x1 - x0 + 0.000001f), 0.0f);
    tmo = ws->curve[j0] * (1.0f - f) + ws->curve[j] * f;
  }
  int k = 0;
  while (k < cm.n - 1 && cm.in[k] < d) ++k;  // bucketize
  const int k0 = max(k - 1, 0);
  const float fk = (k == k0) ? 0.0f : fmaxf((d - cm.in[k0]) / (cm.in[k] - cm.in[k0] + 0.000001f), 0.0f);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float v = cm.ch[k0][c] * (1.0f - fk) + cm.ch[k][c] * fk;
    out[c * n + i] = __float2half(fminf(fmaxf(v * tmo, 0.0f), 1.0f));
  }
}

// ------------------------------------------------------------------------------------------------
// K_yuv: planar Y'CbCr frame (8 / 10..16 bit, 4:2:0 or 4:4:4, limited range) -> display-encoded RGB -> luminance
// (video_source_yuv.py:157-228, video_source_file.py:219-276: fixed2float, bilinear chroma upsampling = torch
// interpolate(scale_factor=2, mode='bilinear'), ycbcr2rgb matrix, clip; then the display model and RGB2Y of
// fvvdp_video_source_dm, video_source_yuv.py:299-302)
// ------------------------------------------------------------------------------------------------
struct YuvParams {
  const void* y;
  const void* u;
  const void* v;
  int W, H, cw, ch, is16, is420;
  float wy, oy, wc, oc;   // fixed2float: Y' = clip(wy * Y - oy, 0, 1), C = clip(wc * c - oc, -0.5, 0.5)
  float m[9];             // ycbcr2rgb, row-major
  int eotf;
  float Yscale, Y_black, Y_peak, gamma, L_min, L_max;
  float rgb2y[3];
  float* lum;             // [H][W] or nullptr
  float* rgb;             // [H][W][3] display-encoded, clipped to [0,1], or nullptr
};

__device__ __forceinline__ float yuv_sample(const void* p, long long i, int is16) {
  return is16 ? (float)__ldg(reinterpret_cast<const unsigned short*>(p) + i) : (float)__ldg(reinterpret_cast<const unsigned char*>(p) + i);
}
__device__ __forceinline__ float yuv_chroma(const YuvParams& p, const void* plane, int cy, int cx) {
  return fminf(fmaxf(p.wc * yuv_sample(plane, (long long)cy * p.cw + cx, p.is16) - p.oc, -0.5f), 0.5f);
}
// display EOTF of one clipped R'G'B' sample (KIND = fvvdp_b200_eotf); PQ takes its quotient in the log2 domain
template <int KIND>
__device__ __forceinline__ float yuv_eotf(float v, const YuvParams& p) {
  if (KIND == FVVDP_B200_EOTF_NONE) return v;
  if (KIND == FVVDP_B200_EOTF_ABSOLUTE) return fminf(fmaxf(v, p.L_min), p.L_max);
  if (KIND == FVVDP_B200_EOTF_LINEAR) return fminf(fmaxf(v, 0.005f), p.Y_peak) + p.Y_black;
  if (KIND == FVVDP_B200_EOTF_SRGB) {
    const float lin = (v > 0.04045f) ? fast_exp2(2.4f * fast_log2(fmaf(v, 1.0f / 1.055f, 0.055f / 1.055f))) : v * (1.0f / 12.92f);
    return fmaf(p.Yscale, lin, p.Y_black);
  }
  if (KIND == FVVDP_B200_EOTF_GAMMA) return fmaf(p.Yscale, fast_pow(v, p.gamma), p.Y_black);
  // PQ, fvvdp_display_model.py:100-112: L = 1e4 (max(t - c1, 0) / (c2 - c3 t))^(1/n), t = V^(1/m)
  const float t = fast_exp2(fast_log2(v) * (1.0f / 78.843750000000000f));
  const float L = fast_exp2(fmaf(fast_log2(fmaxf(t - 0.83593750000000000f, 0.0f)) - fast_log2(fmaf(-18.687500000000000f, t, 18.851562500000000f)),
                                 1.0f / 0.15930175781250000f, 13.287712379549449f));
  return fminf(fmaxf(L, 0.005f), p.Y_peak) + p.Y_black;
}

// one thread: two horizontally adjacent pixels (they share the rows of the chroma taps)
template <int KIND>
__global__ void __launch_bounds__(256) yuv_kernel(const YuvParams p) {
  const int x = 2 * (blockIdx.x * 32 + (threadIdx.x & 31)), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (x >= p.W || y >= p.H) return;
  const bool two = x + 1 < p.W;
  float cb[2], cr[2];
  if (p.is420) {
    // bilinear, align_corners = False: source coordinate max((dst + 0.5) / 2 - 0.5, 0); neighbour clamped to the last sample
    const float sy = fmaxf(0.5f * (float)y - 0.25f, 0.0f);
    const int y0 = (int)sy, y1 = min(y0 + 1, p.ch - 1);
    const float ly = sy - (float)y0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float sx = fmaxf(0.5f * (float)(x + k) - 0.25f, 0.0f);
      const int x0 = (int)sx, x1 = min(x0 + 1, p.cw - 1);
      const float lx = sx - (float)x0;
      cb[k] = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, p.u, y0, x0) + lx * yuv_chroma(p, p.u, y0, x1)) +
              ly * ((1.0f - lx) * yuv_chroma(p, p.u, y1, x0) + lx * yuv_chroma(p, p.u, y1, x1));
      cr[k] = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, p.v, y0, x0) + lx * yuv_chroma(p, p.v, y0, x1)) +
              ly * ((1.0f - lx) * yuv_chroma(p, p.v, y1, x0) + lx * yuv_chroma(p, p.v, y1, x1));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      cb[k] = yuv_chroma(p, p.u, y, min(x + k, p.W - 1));
      cr[k] = yuv_chroma(p, p.v, y, min(x + k, p.W - 1));
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (k == 1 && !two) break;
    const long long i = (long long)y * p.W + x + k;
    const float Y = fminf(fmaxf(p.wy * yuv_sample(p.y, i, p.is16) - p.oy, 0.0f), 1.0f);
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = fminf(fmaxf(p.m[3 * c] * Y + p.m[3 * c + 1] * cb[k] + p.m[3 * c + 2] * cr[k], 0.0f), 1.0f);
    if (p.rgb) { p.rgb[3 * i] = rgb[0]; p.rgb[3 * i + 1] = rgb[1]; p.rgb[3 * i + 2] = rgb[2]; }
    if (p.lum) p.lum[i] = yuv_eotf<KIND>(rgb[0], p) * p.rgb2y[0] + yuv_eotf<KIND>(rgb[1], p) * p.rgb2y[1] + yuv_eotf<KIND>(rgb[2], p) * p.rgb2y[2];
  }
}

// Block version for the metric: the window slots of BOTH streams in one launch, written as (test, reference) planes in the
// pyramid layout ([slot][row][2 * column + stream]) so that level 0 of the band kernels stages them by TMA like any other level.
// Frames are DEVICE copies of the file's frames as stored (Y plane, Cb plane, Cr plane).
struct YuvBlockParams {
  YuvParams f;                                    // format and display model (y / u / v / lum / rgb unused)
  const void* frame[2][FVVDP_B200_MAX_SLOTS];     // [test|ref][slot]
  unsigned char skip[FVVDP_B200_MAX_SLOTS];       // slots the band kernels never stage (repeats of the first frame)
  long long y_elems, c_elems;                     // samples of the luma plane / of one chroma plane
  float* out;                                     // [slot][H][pitch]
  long long slot_stride;
  int pitch;
};

template <int KIND>
__device__ __forceinline__ void yuv_two_pixels(YuvParams p, int x, int y, float (&lum)[2]) {
  float cb[2], cr[2];
  if (p.is420) {
    const float sy = fmaxf(0.5f * (float)y - 0.25f, 0.0f);
    const int y0 = (int)sy, y1 = min(y0 + 1, p.ch - 1);
    const float ly = sy - (float)y0;
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      const float sx = fmaxf(0.5f * (float)min(x + k, p.W - 1) - 0.25f, 0.0f);
      const int x0 = (int)sx, x1 = min(x0 + 1, p.cw - 1);
      const float lx = sx - (float)x0;
      cb[k] = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, p.u, y0, x0) + lx * yuv_chroma(p, p.u, y0, x1)) +
              ly * ((1.0f - lx) * yuv_chroma(p, p.u, y1, x0) + lx * yuv_chroma(p, p.u, y1, x1));
      cr[k] = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, p.v, y0, x0) + lx * yuv_chroma(p, p.v, y0, x1)) +
              ly * ((1.0f - lx) * yuv_chroma(p, p.v, y1, x0) + lx * yuv_chroma(p, p.v, y1, x1));
    }
  } else {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      cb[k] = yuv_chroma(p, p.u, y, min(x + k, p.W - 1));
      cr[k] = yuv_chroma(p, p.v, y, min(x + k, p.W - 1));
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const long long i = (long long)y * p.W + min(x + k, p.W - 1);
    const float Y = fminf(fmaxf(p.wy * yuv_sample(p.y, i, p.is16) - p.oy, 0.0f), 1.0f);
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = fminf(fmaxf(p.m[3 * c] * Y + p.m[3 * c + 1] * cb[k] + p.m[3 * c + 2] * cr[k], 0.0f), 1.0f);
    lum[k] = yuv_eotf<KIND>(rgb[0], p) * p.rgb2y[0] + yuv_eotf<KIND>(rgb[1], p) * p.rgb2y[1] + yuv_eotf<KIND>(rgb[2], p) * p.rgb2y[2];
  }
}

template <int KIND>
__global__ void __launch_bounds__(256) yuv_planes_kernel(const __grid_constant__ YuvBlockParams q) {
  const int x = 2 * (blockIdx.x * 32 + (threadIdx.x & 31)), y = blockIdx.y * 8 + (threadIdx.x >> 5), slot = blockIdx.z;
  if (x >= q.f.W || y >= q.f.H || q.skip[slot]) return;
  float lum[2][2];
#pragma unroll
  for (int st = 0; st < 2; ++st) {
    YuvParams p = q.f;
    const char* base = reinterpret_cast<const char*>(q.frame[st][slot]);
    const int esz = p.is16 ? 2 : 1;
    p.y = base;
    p.u = base + q.y_elems * esz;
    p.v = base + (q.y_elems + q.c_elems) * esz;
    yuv_two_pixels<KIND>(p, x, y, lum[st]);
  }
  float* o = q.out + slot * q.slot_stride + (long long)y * q.pitch + 2 * x;
  if (x + 1 < q.f.W) *reinterpret_cast<float4*>(o) = make_float4(lum[0][0], lum[1][0], lum[0][1], lum[1][1]);
  else *reinterpret_cast<float2*>(o) = make_float2(lum[0][0], lum[1][0]);
}

// ------------------------------------------------------------------------------------------------
// Full-screen resize of .yuv clips (fvvdp_video_source_yuv_file._get_frame, video_source_yuv.py:293-297: the display-encoded
// R'G'B' frame goes through torch.nn.functional.interpolate(size=display resolution, mode=nearest|bilinear|bicubic|area),
// align_corners unset, then .clip(0, 1), then the display model).  Here a CTA converts the source pixels its 32x8 output
// pixels touch from the planar Y'CbCr frame into shared memory (like yuv_kernel does) and every output pixel gathers its taps
// from there, so no R'G'B' frame at the clip's own resolution is ever written to HBM.  Tap positions and weights follow ATen's upsample kernels: scale = in / out in float; nearest
// min(floor(dst * scale), in - 1); bilinear src = max(scale (dst + 0.5) - 0.5, 0); bicubic src = scale (dst + 0.5) - 0.5,
// A = -0.75, taps clamped to the frame; area = adaptive average pooling over [floor(i in / out), ceil((i + 1) in / out)).
// ------------------------------------------------------------------------------------------------
struct ResizeParams {
  int mode;        // fvvdp_b200_resize
  int outW, outH;
  float sx, sy;    // (float)in / out
};

// display-encoded, clipped R'G'B' of ONE source pixel (py / pu / pv: the frame's planes; p: format, read where it lies -- the
// kernel parameters -- so nothing of it is copied to the stack)
__device__ __forceinline__ void yuv_rgb_at(const YuvParams& p, const void* py, const void* pu, const void* pv, int x, int y, float (&rgb)[3]) {
  float cb, cr;
  if (p.is420) {
    const float sy = fmaxf(0.5f * (float)y - 0.25f, 0.0f), sx = fmaxf(0.5f * (float)x - 0.25f, 0.0f);
    const int y0 = (int)sy, y1 = min(y0 + 1, p.ch - 1), x0 = (int)sx, x1 = min(x0 + 1, p.cw - 1);
    const float ly = sy - (float)y0, lx = sx - (float)x0;
    cb = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, pu, y0, x0) + lx * yuv_chroma(p, pu, y0, x1)) +
         ly * ((1.0f - lx) * yuv_chroma(p, pu, y1, x0) + lx * yuv_chroma(p, pu, y1, x1));
    cr = (1.0f - ly) * ((1.0f - lx) * yuv_chroma(p, pv, y0, x0) + lx * yuv_chroma(p, pv, y0, x1)) +
         ly * ((1.0f - lx) * yuv_chroma(p, pv, y1, x0) + lx * yuv_chroma(p, pv, y1, x1));
  } else {
    cb = yuv_chroma(p, pu, y, x);
    cr = yuv_chroma(p, pv, y, x);
  }
  const float Y = fminf(fmaxf(p.wy * yuv_sample(py, (long long)y * p.W + x, p.is16) - p.oy, 0.0f), 1.0f);
#pragma unroll
  for (int c = 0; c < 3; ++c) rgb[c] = fminf(fmaxf(p.m[3 * c] * Y + p.m[3 * c + 1] * cb + p.m[3 * c + 2] * cr, 0.0f), 1.0f);
}

__device__ __forceinline__ void cubic_weights(float t, float (&w)[4]) {  // ATen get_cubic_upsample_coefficients, A = -0.75
  const float A = -0.75f;
  const float a = t + 1.0f, b = t, c = 1.0f - t, d = 2.0f - t;
  w[0] = ((A * a - 5.0f * A) * a + 8.0f * A) * a - 4.0f * A;
  w[1] = ((A + 2.0f) * b - (A + 3.0f)) * b * b + 1.0f;
  w[2] = ((A + 2.0f) * c - (A + 3.0f)) * c * c + 1.0f;
  w[3] = ((A * d - 5.0f * A) * d + 8.0f * A) * d - 4.0f * A;
}

// where the taps of a resized pixel come from: converted on the fly from the planar frame, or read from the CTA's patch of
// already converted source pixels in shared memory ([3][RESIZE_PATCH], origin (x_lo, y_lo), row length pw)
constexpr int RESIZE_PATCH = 3072;   // source pixels a CTA converts once for its 32x8 output pixels (36 kB); larger footprints
                                     // (down-scaling by more than ~3.5) convert per tap
struct YuvTapDirect {
  const YuvParams& p;
  const void *py, *pu, *pv;
  __device__ __forceinline__ void operator()(int x, int y, float (&rgb)[3]) const { yuv_rgb_at(p, py, pu, pv, x, y, rgb); }
};
struct YuvTapPatch {
  const float* s;
  int x_lo, y_lo, pw;
  __device__ __forceinline__ void operator()(int x, int y, float (&rgb)[3]) const {
    const int i = (y - y_lo) * pw + (x - x_lo);
    rgb[0] = s[i]; rgb[1] = s[RESIZE_PATCH + i]; rgb[2] = s[2 * RESIZE_PATCH + i];
  }
};

// first and last source sample along one axis that the output samples o_first..o_last touch (the tap rules of resized_rgb)
__device__ __forceinline__ void resize_src_range(int mode, float s, int n_in, int n_out, int o_first, int o_last, int& lo, int& hi) {
  if (mode == FVVDP_B200_RESIZE_NEAREST) {
    lo = min((int)floorf((float)o_first * s), n_in - 1);
    hi = min((int)floorf((float)o_last * s), n_in - 1);
  } else if (mode == FVVDP_B200_RESIZE_BILINEAR) {
    lo = min((int)fmaxf(s * ((float)o_first + 0.5f) - 0.5f, 0.0f), n_in - 1);
    hi = min(min((int)fmaxf(s * ((float)o_last + 0.5f) - 0.5f, 0.0f), n_in - 1) + 1, n_in - 1);
  } else if (mode == FVVDP_B200_RESIZE_BICUBIC) {
    lo = max(min((int)floorf(s * ((float)o_first + 0.5f) - 0.5f) - 1, n_in - 1), 0);
    hi = max(min((int)floorf(s * ((float)o_last + 0.5f) - 0.5f) + 2, n_in - 1), 0);
  } else {
    lo = (int)(((long long)o_first * n_in) / n_out);
    hi = (int)((((long long)o_last + 1) * n_in + n_out - 1) / n_out) - 1;
  }
}

// resized, clipped R'G'B' of output pixel (ox, oy) of a W x H source
template <class Tap>
__device__ __forceinline__ void resized_rgb(const Tap& tap, int W, int H, const ResizeParams& r, int ox, int oy, float (&rgb)[3]) {
  float t[3];
  rgb[0] = rgb[1] = rgb[2] = 0.0f;
  if (r.mode == FVVDP_B200_RESIZE_NEAREST) {
    tap(min((int)floorf((float)ox * r.sx), W - 1), min((int)floorf((float)oy * r.sy), H - 1), rgb);
  } else if (r.mode == FVVDP_B200_RESIZE_BILINEAR) {
    const float fx = fmaxf(r.sx * ((float)ox + 0.5f) - 0.5f, 0.0f), fy = fmaxf(r.sy * ((float)oy + 0.5f) - 0.5f, 0.0f);
    const int x0 = min((int)fx, W - 1), y0 = min((int)fy, H - 1), x1 = x0 + (x0 < W - 1 ? 1 : 0), y1 = y0 + (y0 < H - 1 ? 1 : 0);
    const float lx = fx - (float)x0, ly = fy - (float)y0, kx = 1.0f - lx, ky = 1.0f - ly;
    float a[3], b[3];
    tap(x0, y0, a);
    tap(x1, y0, b);
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] = ky * (kx * a[c] + lx * b[c]);
    tap(x0, y1, a);
    tap(x1, y1, b);
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] += ly * (kx * a[c] + lx * b[c]);
  } else if (r.mode == FVVDP_B200_RESIZE_BICUBIC) {
    const float fx = r.sx * ((float)ox + 0.5f) - 0.5f, fy = r.sy * ((float)oy + 0.5f) - 0.5f;
    const float bx = floorf(fx), by = floorf(fy);
    float wx[4], wy[4];
    cubic_weights(fx - bx, wx);
    cubic_weights(fy - by, wy);
    for (int j = 0; j < 4; ++j) {
      const int yy = max(min((int)by - 1 + j, H - 1), 0);
      float row[3] = {0.0f, 0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        tap(max(min((int)bx - 1 + i, W - 1), 0), yy, t);
#pragma unroll
        for (int c = 0; c < 3; ++c) row[c] += t[c] * wx[i];
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) rgb[c] += row[c] * wy[j];
    }
  } else {  // area
    const int x0 = (int)(((long long)ox * W) / r.outW), x1 = (int)((((long long)ox + 1) * W + r.outW - 1) / r.outW);
    const int y0 = (int)(((long long)oy * H) / r.outH), y1 = (int)((((long long)oy + 1) * H + r.outH - 1) / r.outH);
    for (int yy = y0; yy < y1; ++yy)
      for (int xx = x0; xx < x1; ++xx) {
        tap(xx, yy, t);
#pragma unroll
        for (int c = 0; c < 3; ++c) rgb[c] += t[c];
      }
    const float n = (float)((y1 - y0) * (x1 - x0));
#pragma unroll
    for (int c = 0; c < 3; ++c) rgb[c] /= n;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) rgb[c] = fminf(fmaxf(rgb[c], 0.0f), 1.0f);
}

// The CTA's 32x8 output pixels of one frame: the source pixels they touch are converted ONCE into shared memory (a 2x up-scale
// with bicubic taps converts 19x7 source pixels instead of 256 x 16 taps), then every output pixel gathers its taps from there.
// All 256 threads call this (barriers inside); `inside` = this thread's pixel lies in the output frame.
__device__ __forceinline__ void yuv_resized_tile(const YuvParams& p, const void* py, const void* pu, const void* pv, const ResizeParams& r, float* sP,
                                                 int ox, int oy, bool inside, float (&rgb)[3]) {
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * 8;
  int x_lo, x_hi, y_lo, y_hi;
  resize_src_range(r.mode, r.sx, p.W, r.outW, ox0, min(ox0 + 31, r.outW - 1), x_lo, x_hi);
  resize_src_range(r.mode, r.sy, p.H, r.outH, oy0, min(oy0 + 7, r.outH - 1), y_lo, y_hi);
  const int pw = x_hi - x_lo + 1, n = pw * (y_hi - y_lo + 1);
  if (n <= RESIZE_PATCH) {  // uniform over the CTA
    __syncthreads();        // the previous user of the patch is done with it
    for (int i = threadIdx.x; i < n; i += 256) {
      float t[3];
      yuv_rgb_at(p, py, pu, pv, x_lo + i % pw, y_lo + i / pw, t);
      sP[i] = t[0]; sP[RESIZE_PATCH + i] = t[1]; sP[2 * RESIZE_PATCH + i] = t[2];
    }
    __syncthreads();
    if (inside) resized_rgb(YuvTapPatch{sP, x_lo, y_lo, pw}, p.W, p.H, r, ox, oy, rgb);
  } else if (inside) {
    resized_rgb(YuvTapDirect{p, py, pu, pv}, p.W, p.H, r, ox, oy, rgb);
  }
}

// one frame: luminance [outH][outW] and / or resized R'G'B' [outH][outW][3]
template <int KIND>
__global__ void __launch_bounds__(256) yuv_resize_kernel(const __grid_constant__ YuvParams p, const __grid_constant__ ResizeParams r) {
  __shared__ float sP[3 * RESIZE_PATCH];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
  const bool inside = x < r.outW && y < r.outH;
  float rgb[3] = {0.0f, 0.0f, 0.0f};
  yuv_resized_tile(p, p.y, p.u, p.v, r, sP, x, y, inside, rgb);
  if (!inside) return;
  const long long i = (long long)y * r.outW + x;
  if (p.rgb) { p.rgb[3 * i] = rgb[0]; p.rgb[3 * i + 1] = rgb[1]; p.rgb[3 * i + 2] = rgb[2]; }
  if (p.lum) p.lum[i] = yuv_eotf<KIND>(rgb[0], p) * p.rgb2y[0] + yuv_eotf<KIND>(rgb[1], p) * p.rgb2y[1] + yuv_eotf<KIND>(rgb[2], p) * p.rgb2y[2];
}

// block version: the window slots of both streams into the (test, reference) planes level 0 stages (see yuv_planes_kernel)
template <int KIND>
__global__ void __launch_bounds__(256) yuv_resize_planes_kernel(const __grid_constant__ YuvBlockParams q, const __grid_constant__ ResizeParams r) {
  __shared__ float sP[3 * RESIZE_PATCH];
  const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5), slot = blockIdx.z;
  const bool inside = x < r.outW && y < r.outH;
  const YuvParams& p = q.f;
  float lum[2];
#pragma unroll 1
  for (int st = 0; st < 2; ++st) {
    const char* base = reinterpret_cast<const char*>(q.frame[st][slot]);
    const int esz = p.is16 ? 2 : 1;
    float rgb[3] = {0.0f, 0.0f, 0.0f};
    yuv_resized_tile(p, base, base + q.y_elems * esz, base + (q.y_elems + q.c_elems) * esz, r, sP, x, y, inside, rgb);
    lum[st] = yuv_eotf<KIND>(rgb[0], p) * p.rgb2y[0] + yuv_eotf<KIND>(rgb[1], p) * p.rgb2y[1] + yuv_eotf<KIND>(rgb[2], p) * p.rgb2y[2];
  }
  if (inside) *reinterpret_cast<float2*>(q.out + slot * q.slot_stride + (long long)y * q.pitch + 2 * x) = make_float2(lum[0], lum[1]);
}

// ------------------------------------------------------------------------------------------------
// K_pu: PU21-PSNR frame term (pupsnr.py:52-79, utils.py:157-202): sum over the frame of (PU(T) - PU(R))^2 with
// PU(Y) = p6 (((p0 + p1 Y^p3) / (1 + p2 Y^p3))^p4 - p5), Y clipped to [L_min, L_max]
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pu_sqerr_kernel(const float* __restrict__ t, const float* __restrict__ r, long long n, PuParams q,
                                                       double* __restrict__ acc) {
  float s = 0.0f;
  for (long long i = blockIdx.x * 256ll + threadIdx.x; i < n; i += 256ll * gridDim.x) {
    // p6 * ((a - p5) - (b - p5)) = p6 * (a - b): exactly 0 for identical frames (a fused multiply-add of the two scaled terms is not)
    const float d = q.p[6] * (pu_encode(__ldg(t + i), q) - pu_encode(__ldg(r + i), q));
    s = fmaf(d, d, s);
  }
  double v = (double)s;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __shared__ double sh[8];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double tot = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += sh[k];
    atomicAdd(acc, tot);
  }
}

}  // namespace fvvdp
