// Host side of libfvvdp_b200.so: the C ABI declared in include/fvvdp_b200.h.
// Everything is enqueued on the caller's stream; no host synchronisation in score_block.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <algorithm>
#include <new>
#include <vector>

#include <stdlib.h>

#include "fvvdp_fused.cuh"
#include "fvvdp_fused_launch.h"
#include "fvvdp_kernels.cuh"
#include "fvvdp_ws_geometry.h"

using namespace fvvdp;

struct fvvdp_b200_ctx {
  fvvdp_b200_config cfg;
  int dev = 0;
  int nch = 4, n_bands = 0, T = 1;
  int lh[FVVDP_B200_MAX_LEVELS], lw[FVVDP_B200_MAX_LEVELS];
  int tiles_x[FVVDP_B200_MAX_LEVELS], tiles_y[FVVDP_B200_MAX_LEVELS];
  bool fused = false;                          // band kernels (fused / warp-specialised) or the general v1 path
  bool ch2 = false;                            // band kernels on two filtered planes per slot (filter_len 17..32): front_kernel, pairs output
  float* P[FVVDP_B200_MAX_LEVELS] = {};        // fused: luminance pyramid, level >= 1: [slots][h_l][pitch_l], (test, ref) interleaved
  int pitch[FVVDP_B200_MAX_LEVELS] = {};
  float* cell = nullptr;                       // fused: [n_bands][32][8] CSF cells over log2 Y
  CUtensorMap pmap[FVVDP_B200_MAX_LEVELS];     // fused: TMA descriptors of P[l] (2x + stream, y, slot), box = staged tile
  CUtensorMap pmap_ws[FVVDP_B200_MAX_LEVELS];  //   the same tensors with the staged-tile box of the warp-specialised kernel
  int ws_th = ws::TH, ws_rp = ws::RP;                   //   its tile height / ring positions: ws (<= 8 taps) or ws16 (<= 16 taps)
  int ws_max_level = -1;                       // warp-specialised kernel on levels 0..ws_max_level (video, <= 8 taps, no debug outputs)
  int ws_min_tiles = 0;                        // ... and, past level 0, only where a level has at least this many tiles (one per SM); the
                                               //   small levels go to the fused kernel (2 CTAs per SM, time chunks).  0 with FVVDP_B200_WS_LEVELS
  int ntiles_used[FVVDP_B200_MAX_LEVELS] = {}; // tiles of the kernel that scored each level of the last block
  bool no_dup_skip = false;                    // A/B switch FVVDP_B200_NO_DUP_SKIP
  float* G[FVVDP_B200_MAX_LEVELS] = {};        // v1 (and taps): G[0] = R; [T][nch][h_l][w_l]
  float* partial[FVVDP_B200_MAX_LEVELS] = {};  // [T][2][ntiles_l]
  float* tapC[FVVDP_B200_MAX_LEVELS] = {};
  float* tapL[FVVDP_B200_MAX_LEVELS] = {};
  float* tapS[FVVDP_B200_MAX_LEVELS] = {};
  float* tapD[FVVDP_B200_MAX_LEVELS] = {};
  float* dmap[FVVDP_B200_MAX_LEVELS] = {};
  float* recon[2] = {};                        // ping-pong buffers for the heat-map reconstruction
  float* ctxmap = nullptr;                     // fused, want_dmap == 2: [T][H][W] sustained test frames (context of the visualisation)
  VisWork* vis = nullptr;                      // want_dmap == 2: scratch of the heat-map visualisation
  const float* fov_view[FVVDP_B200_MAX_LEVELS] = {};  // custom display geometry (caller-owned): [2][h_l][w_l] view directions
  const float* fov_rq[FVVDP_B200_MAX_LEVELS] = {};    //   and [h_l][w_l] log2(clamped rho)
  float* axes = nullptr;                       // x[3][32], inv[3][32]
  float* csf1d = nullptr;                      // [n_bands][2][32]
  float* lut3d = nullptr;                      // [2][32][32][32]
  float4* lut4 = nullptr;                      // fused, foveated: [rho 32][ecc 32][Y 32] (t0, dt0/dY-cell, t1, dt1/dY-cell), t = log2(S * sens_mul)
  float* vx[FVVDP_B200_MAX_LEVELS] = {};
  float* vy[FVVDP_B200_MAX_LEVELS] = {};
  CsfAxes ax;
  float log2_sens_mul = 0.f;
  int last_n_frames = 0;
  int64_t launches = 0;
  bool profiling = false;
  std::vector<cudaEvent_t> ev;  // pairs (start, stop)
  std::vector<int> ev_class;
  size_t ev_used = 0;
  double bytes_alg = 0, bytes_plan = 0;
  char err[512] = {0};
};

static char g_create_err[512] = "";

static int fail(fvvdp_b200_ctx* c, int code, const char* fmt, ...) {
  char* dst = c ? c->err : g_create_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  return code;
}

#define CU(call)                                                                                       \
  do {                                                                                                 \
    cudaError_t e_ = (call);                                                                           \
    if (e_ != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

static void locate_host(float q, const float* x, int& j0, int& j1, float& f) {
  // get_interpolants_v1, interp.py:11-20
  int j = 0;
  while (j < 32 && !(x[j] >= q)) ++j;  // bucketize(right=False)
  if (j > 31) j = 31;
  j1 = j;
  j0 = j - 1 < 0 ? 0 : j - 1;
  f = (q - x[j0]) / (x[j1] - x[j0] + 0.000001f);
  if (j1 == j0) f = 0.f;
  if (f < 0.f) f = 0.f;
}

static void free_ctx(fvvdp_b200_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->dev);
  for (int l = 0; l < FVVDP_B200_MAX_LEVELS; ++l) {
    cudaFree(c->G[l]); cudaFree(c->partial[l]); cudaFree(c->tapC[l]); cudaFree(c->tapL[l]);
    cudaFree(c->tapS[l]); cudaFree(c->tapD[l]); cudaFree(c->dmap[l]); cudaFree(c->vx[l]); cudaFree(c->vy[l]);
    cudaFree(c->P[l]);
  }
  cudaFree(c->cell);
  cudaFree(c->recon[0]); cudaFree(c->recon[1]);
  cudaFree(c->ctxmap); cudaFree(c->vis);
  cudaFree(c->axes); cudaFree(c->csf1d); cudaFree(c->lut3d); cudaFree(c->lut4);
  for (cudaEvent_t e : c->ev) cudaEventDestroy(e);
  delete c;
}

// cuTensorMapEncodeTiled resolved through the runtime (no link-time dependency on libcuda)
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = (encode_tiled_fn)p;
  }
  return fn;
}
// 3-D float tensor (dims[0] innermost, strides in bytes for dims 1..), box = staged tile (box0 x LH x 1), zero fill outside
static bool make_tile_map(CUtensorMap* map, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes, int box0,
                          int box1 = fused::LH) {
  encode_tiled_fn fn = get_encode_tiled();
  if (!fn) return false;
  cuuint32_t box[4] = {(cuuint32_t)box0, (cuuint32_t)box1, 1, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

extern "C" int fvvdp_b200_abi_version(void) { return FVVDP_B200_ABI_VERSION; }

extern "C" const char* fvvdp_b200_last_error(const fvvdp_b200_ctx* ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int fvvdp_b200_create(const fvvdp_b200_config* cfg, int cuda_device, fvvdp_b200_ctx** out) {
  fvvdp_b200_ctx* ctx = nullptr;  // errors before allocation go to g_create_err
  if (!cfg || !out) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->abi_version != FVVDP_B200_ABI_VERSION) return fail(ctx, FVVDP_B200_ERR_INVALID, "ABI version mismatch (%d vs %d)", cfg->abi_version, FVVDP_B200_ABI_VERSION);
  if (cfg->width < 4 || cfg->height < 4) return fail(ctx, FVVDP_B200_ERR_INVALID, "frame too small (%dx%d)", cfg->width, cfg->height);
  if (cfg->n_levels < 2 || cfg->n_levels > FVVDP_B200_MAX_LEVELS) return fail(ctx, FVVDP_B200_ERR_INVALID, "n_levels out of range: %d", cfg->n_levels);
  if (cfg->temp_ch != 1 && cfg->temp_ch != 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "temp_ch must be 1 or 2");
  if (cfg->filter_len < 1 || cfg->filter_len > FVVDP_B200_MAX_FILTER_LEN) return fail(ctx, FVVDP_B200_ERR_INVALID, "filter_len %d not in 1..%d", cfg->filter_len, FVVDP_B200_MAX_FILTER_LEN);
  if (cfg->in_channels != 1 && cfg->in_channels != 3) return fail(ctx, FVVDP_B200_ERR_INVALID, "The content must have either 1 or 3 colour channels.");
  if (cfg->in_dtype < 0 || cfg->in_dtype > 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "Only uint8, uint16 and float32 is currently supported");
  if (cfg->eotf < 0 || cfg->eotf > 5) return fail(ctx, FVVDP_B200_ERR_INVALID, "Unknown EOTF %d", cfg->eotf);
  if (cfg->max_block_frames < 1 || cfg->max_block_frames > FVVDP_B200_MAX_BLOCK_FRAMES) return fail(ctx, FVVDP_B200_ERR_INVALID, "max_block_frames %d not in 1..%d", cfg->max_block_frames, FVVDP_B200_MAX_BLOCK_FRAMES);
  if (!cfg->csf_rho_log || !cfg->csf_Y_log || !cfg->csf_ecc_sqrt || !cfg->csf_S_log) return fail(ctx, FVVDP_B200_ERR_INVALID, "CSF look-up tables missing");
  {
    int hh = cfg->height, ww = cfg->width;
    for (int l = 0; l + 1 < cfg->n_levels; ++l) {  // every scored band needs >= 2 rows and columns
      if (hh < 2 || ww < 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "too many pyramid levels (%d) for %dx%d", cfg->n_levels, cfg->width, cfg->height);
      hh = (hh + 1) / 2; ww = (ww + 1) / 2;
    }
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) return fail(ctx, FVVDP_B200_ERR_CUDA, "no CUDA device available (the B200 core has no CPU fallback)");
  if (cuda_device < 0 || cuda_device >= ndev) return fail(ctx, FVVDP_B200_ERR_INVALID, "cuda_device %d out of range", cuda_device);

  fvvdp_b200_ctx* c = new (std::nothrow) fvvdp_b200_ctx();
  if (!c) return fail(ctx, FVVDP_B200_ERR_NOMEM, "out of host memory");
  ctx = c;
  c->cfg = *cfg;
  c->cfg.csf_rho_log = c->cfg.csf_Y_log = c->cfg.csf_ecc_sqrt = c->cfg.csf_S_log = nullptr;  // host tables are not retained
  c->dev = cuda_device;
  c->nch = 2 * cfg->temp_ch;
  c->n_bands = cfg->n_levels - 1;
  c->T = cfg->max_block_frames;
#define CUC(call)                                                                                     \
  do {                                                                                                \
    cudaError_t e_ = (call);                                                                          \
    if (e_ != cudaSuccess) {                                                                          \
      fail(nullptr, FVVDP_B200_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_));                    \
      free_ctx(c);                                                                                    \
      return e_ == cudaErrorMemoryAllocation ? FVVDP_B200_ERR_NOMEM : FVVDP_B200_ERR_CUDA;            \
    }                                                                                                 \
  } while (0)
  CUC(cudaSetDevice(cuda_device));
  const int T = c->T, nch = c->nch;
  int hh = cfg->height, ww = cfg->width;
  {
    // A/B switches, read once: FVVDP_B200_PATH=v1 (general kernels) | fused (no warp-specialised kernel);
    // FVVDP_B200_WS_LEVELS=n (warp-specialised kernel on levels < n; default: level 0 only for windows of up to 8 taps, where it
    // only ties the fused kernel on the pyramid levels, every level for 9..16 taps)
    const char* force = getenv("FVVDP_B200_PATH");
    const bool v1 = force && strcmp(force, "v1") == 0;
    // windows of 17..32 taps: the filters run in a register-ring walk of their own and the band kernels take two planes per slot
    const char* ct = getenv("FVVDP_B200_CH2_TAPS");  // A/B switch: smallest window that takes this route (default 17)
    const int ch2_min = ct ? atoi(ct) : fused::MAXRING + 1;
    c->ch2 = cfg->filter_len >= ch2_min && cfg->filter_len >= 2 && cfg->temp_ch == 2 && !cfg->want_taps && !cfg->want_dmap && !v1;
    c->fused = (cfg->filter_len <= fused::MAXRING || c->ch2) && !v1;
    const bool ws_ok = c->fused && !c->ch2 && !(force && strcmp(force, "fused") == 0) && cfg->temp_ch == 2 && cfg->filter_len >= 2 &&
                       cfg->filter_len <= ws16::RP + 1 && !cfg->want_taps && !cfg->want_dmap;
    if (cfg->filter_len > ws::RP + 1) { c->ws_th = ws16::TH; c->ws_rp = ws16::RP; }
    const char* wl = getenv("FVVDP_B200_WS_LEVELS");
    c->ws_max_level = ws_ok ? (wl ? atoi(wl) - 1 : (c->ws_rp == ws::RP ? 0 : FVVDP_B200_MAX_LEVELS)) : -1;
    c->ws_min_tiles = wl ? 0 : 148;
    c->no_dup_skip = getenv("FVVDP_B200_NO_DUP_SKIP") != nullptr;
  }
  const int tile_w = c->fused ? fused::TW : TW, tile_h = c->fused ? fused::TH : TH;
  for (int l = 0; l < cfg->n_levels; ++l) {
    c->lh[l] = hh; c->lw[l] = ww;
    c->tiles_x[l] = (ww + tile_w - 1) / tile_w; c->tiles_y[l] = (hh + tile_h - 1) / tile_h;
    c->pitch[l] = (2 * ww + 3) & ~3;  // floats per row of the interleaved (test, ref) pyramid planes
    hh = (hh + 1) / 2; ww = (ww + 1) / 2;
  }
  for (int l = 0; l < cfg->n_levels; ++l) {
    const size_t px = (size_t)c->lh[l] * c->lw[l];
    // v1: 4-channel Gaussian levels of the temporal channels; the last level (base band) only lives in shared memory
    if (l < c->n_bands && (!c->fused || cfg->want_taps)) CUC(cudaMalloc(&c->G[l], sizeof(float) * px * nch * T));
    // fused: 2-plane luminance pyramid per window slot; the row padding must stay zero (it is read as zero padding).  Level 0 has
    // planes of its own when the input is known not to be contiguous single-channel float frames (uint8 / uint16 / RGB go
    // through the luminance front end); strided float views and raw .yuv blocks get theirs on first use
    const bool planes0 = l == 0 && (c->ch2 || cfg->in_dtype != FVVDP_B200_F32 || cfg->in_channels != 1);
    const size_t plane_slots = c->ch2 ? 2 * (size_t)T : (size_t)(T + cfg->filter_len - 1);  // ch2: [2 temporal channels][T frames]
    if (c->fused && ((l >= 1 && l < c->n_bands) || planes0)) {
      const size_t n = plane_slots * c->lh[l] * c->pitch[l];
      CUC(cudaMalloc(&c->P[l], sizeof(float) * n));
      CUC(cudaMemset(c->P[l], 0, sizeof(float) * n));
      const cuuint64_t dims[3] = {(cuuint64_t)(2 * c->lw[l]), (cuuint64_t)c->lh[l], (cuuint64_t)plane_slots};
      const cuuint64_t str[2] = {(cuuint64_t)c->pitch[l] * 4, (cuuint64_t)c->lh[l] * c->pitch[l] * 4};
      if (!make_tile_map(&c->pmap[l], c->P[l], 3, dims, str, 2 * fused::LW) ||
          !make_tile_map(&c->pmap_ws[l], c->P[l], 3, dims, str, 2 * ws::LW, c->ws_th + 8)) {
        fail(nullptr, FVVDP_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed for pyramid level %d", l);
        free_ctx(c);
        return FVVDP_B200_ERR_CUDA;
      }
    }
    if (l < c->n_bands) {
      // the warp-specialised kernels keep one partial sum per consumer warp (16 per tile)
      CUC(cudaMalloc(&c->partial[l], sizeof(float) * (size_t)T * 2 * c->tiles_x[l] * c->tiles_y[l] * (c->ws_max_level >= 0 ? 16 : 1)));
      if (cfg->want_taps) {
        CUC(cudaMalloc(&c->tapC[l], sizeof(float) * px * nch * T));
        CUC(cudaMalloc(&c->tapL[l], sizeof(float) * px * T));
        CUC(cudaMalloc(&c->tapS[l], sizeof(float) * px * cfg->temp_ch * T));
        CUC(cudaMalloc(&c->tapD[l], sizeof(float) * px * cfg->temp_ch * T));
      }
      if (cfg->want_dmap) CUC(cudaMalloc(&c->dmap[l], sizeof(float) * px * T));
    }
  }
  if (cfg->want_taps) {  // keep the base level readable too
    const int l = c->n_bands;
    CUC(cudaMalloc(&c->G[l], sizeof(float) * (size_t)c->lh[l] * c->lw[l] * nch * T));
  }
  if (cfg->want_dmap) {
    CUC(cudaMalloc(&c->recon[0], sizeof(float) * (size_t)cfg->height * cfg->width));
    CUC(cudaMalloc(&c->recon[1], sizeof(float) * (size_t)cfg->height * cfg->width));
  }
  if (cfg->want_dmap >= 2) {
    CUC(cudaMalloc(&c->vis, sizeof(VisWork)));
    if (c->fused && !cfg->want_taps) CUC(cudaMalloc(&c->ctxmap, sizeof(float) * (size_t)cfg->height * cfg->width * T));
  }

  // ---- CSF tables ----
  float hax[6][32];
  const float* src_ax[3] = {cfg->csf_rho_log, cfg->csf_Y_log, cfg->csf_ecc_sqrt};
  for (int a = 0; a < 3; ++a) {
    for (int j = 0; j < 32; ++j) {
      hax[a][j] = src_ax[a][j];
      hax[3 + a][j] = (j == 0) ? 0.f : 1.0f / (src_ax[a][j] - src_ax[a][j - 1] + 0.000001f);
    }
  }
  CUC(cudaMalloc(&c->axes, sizeof(hax)));
  CUC(cudaMemcpy(c->axes, hax, sizeof(hax), cudaMemcpyHostToDevice));
  for (int a = 0; a < 3; ++a) {
    c->ax.x[a] = c->axes + a * 32;
    c->ax.inv[a] = c->axes + (3 + a) * 32;
    c->ax.x0[a] = src_ax[a][0];
    c->ax.inv_dx[a] = 31.0f / (src_ax[a][31] - src_ax[a][0]);
  }
  c->ax.lo[0] = cfg->csf_rho_range[0]; c->ax.hi[0] = cfg->csf_rho_range[1];
  c->ax.lo[1] = cfg->csf_Y_range[0];   c->ax.hi[1] = cfg->csf_Y_range[1];
  c->ax.lo[2] = cfg->csf_ecc_range[0]; c->ax.hi[2] = cfg->csf_ecc_range[1];
  c->log2_sens_mul = log2f(cfg->sens_mul);
  CUC(cudaMalloc(&c->lut3d, sizeof(float) * 2 * 32768));
  CUC(cudaMemcpy(c->lut3d, cfg->csf_S_log, sizeof(float) * 2 * 32768, cudaMemcpyHostToDevice));
  if (cfg->foveated) {
    // both temporal channels and the step to the next Y entry in one 16-byte record, Y innermost: the 8 corners of a
    // trilinear look-up for both channels are four 16-byte loads
    std::vector<float4> l4(32768);
    for (int i = 0; i < 32; ++i)
      for (int k = 0; k < 32; ++k)
        for (int j = 0; j < 32; ++j) {
          const int j1 = j < 31 ? j + 1 : 31;
          const float* v0 = cfg->csf_S_log;
          const float* v1 = cfg->csf_S_log + 32768;
          const float a0 = v0[(j * 32 + i) * 32 + k], b0 = v0[(j1 * 32 + i) * 32 + k];
          const float a1 = v1[(j * 32 + i) * 32 + k], b1 = v1[(j1 * 32 + i) * 32 + k];
          l4[(i * 32 + k) * 32 + j] = make_float4(a0 + c->log2_sens_mul, b0 - a0, a1 + c->log2_sens_mul, b1 - a1);
        }
    CUC(cudaMalloc(&c->lut4, sizeof(float4) * l4.size()));
    CUC(cudaMemcpy(c->lut4, l4.data(), sizeof(float4) * l4.size(), cudaMemcpyHostToDevice));
  }
  {
    // non-foveated: rho and ecc (=0) are constant per band -> 32-entry table over log2 Y per (band, cc)
    // (cached_sensitivity fvvdp.py:520-537 with rho = rho_band[bb], ecc = 0, fvvdp.py:438-442)
    std::vector<float> tab((size_t)c->n_bands * 2 * 32);
    for (int bb = 0; bb < c->n_bands; ++bb) {
      float rho = fminf(fmaxf(cfg->band_freq[bb], cfg->csf_rho_range[0]), cfg->csf_rho_range[1]);
      int i0, i1; float fi;
      locate_host(log2f(rho), cfg->csf_rho_log, i0, i1, fi);
      for (int cc = 0; cc < 2; ++cc)
        for (int j = 0; j < 32; ++j) {
          const float* v = cfg->csf_S_log + (size_t)cc * 32768 + (size_t)j * 1024;
          tab[((size_t)bb * 2 + cc) * 32 + j] = (v[i0 * 32] * (1.0f - fi) + v[i1 * 32] * fi) + c->log2_sens_mul;
        }
    }
    CUC(cudaMalloc(&c->csf1d, sizeof(float) * tab.size()));
    CUC(cudaMemcpy(c->csf1d, tab.data(), sizeof(float) * tab.size(), cudaMemcpyHostToDevice));
    // fused kernels: per band 32 cells {Y_log[j], 1/(Y_log[j+1]-Y_log[j]+1e-6), t0[j], t0[j+1]-t0[j], t1[j], t1[j+1]-t1[j], 0, 0}
    std::vector<float> cell((size_t)c->n_bands * 32 * 8, 0.f);
    for (int bb = 0; bb < c->n_bands; ++bb)
      for (int j = 0; j < 32; ++j) {
        float* q = &cell[((size_t)bb * 32 + j) * 8];
        const int j1 = j < 31 ? j + 1 : 31;
        q[0] = cfg->csf_Y_log[j];
        q[1] = j < 31 ? 1.0f / (cfg->csf_Y_log[j1] - cfg->csf_Y_log[j] + 0.000001f) : 0.f;
        for (int cc = 0; cc < 2; ++cc) {
          q[2 + 2 * cc] = tab[((size_t)bb * 2 + cc) * 32 + j];
          q[3 + 2 * cc] = tab[((size_t)bb * 2 + cc) * 32 + j1] - tab[((size_t)bb * 2 + cc) * 32 + j];
        }
      }
    CUC(cudaMalloc(&c->cell, sizeof(float) * cell.size()));
    CUC(cudaMemcpy(c->cell, cell.data(), sizeof(float) * cell.size(), cudaMemcpyHostToDevice));
  }
  if (cfg->foveated == 1) {
    // pix2view_direction of each band's pixel centres, the band spanning the whole display
    // (fvvdp.py:422-428, fvvdp_display_model.py:498-510)
    for (int l = 0; l < c->n_bands; ++l) {
      std::vector<float> vx(c->lw[l]), vy(c->lh[l]);
      for (int x = 0; x < c->lw[l]; ++x) {
        float xr = ((float)x + 0.5f) - (float)(c->lw[l] / 2.0);
        float xm = xr * cfg->display_size_m[0] / (float)c->lw[l];
        vx[x] = (float)(atan((double)(xm / cfg->distance_m)) * 180.0 / M_PI);
      }
      for (int y = 0; y < c->lh[l]; ++y) {
        float yr = ((float)y + 0.5f) - (float)(c->lh[l] / 2.0);
        float ym = -yr * cfg->display_size_m[1] / (float)c->lh[l];
        vy[y] = (float)(atan((double)(ym / cfg->distance_m)) * 180.0 / M_PI);
      }
      CUC(cudaMalloc(&c->vx[l], sizeof(float) * vx.size()));
      CUC(cudaMalloc(&c->vy[l], sizeof(float) * vy.size()));
      CUC(cudaMemcpy(c->vx[l], vx.data(), sizeof(float) * vx.size(), cudaMemcpyHostToDevice));
      CUC(cudaMemcpy(c->vy[l], vy.data(), sizeof(float) * vy.size(), cudaMemcpyHostToDevice));
    }
  }
  CUC(fused::configure_band_kernels());
  CUC(ws::configure_band_ws_kernels());
  CUC(ws16::configure_band_ws_kernels());
  CUC(cudaFuncSetAttribute(level_kernel<4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)level_smem_bytes(4)));
  CUC(cudaFuncSetAttribute(level_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)level_smem_bytes(4)));
  CUC(cudaFuncSetAttribute(level_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)level_smem_bytes(2)));
  CUC(cudaFuncSetAttribute(level_kernel<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)level_smem_bytes(2)));
#undef CUC
  *out = c;
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_destroy(fvvdp_b200_ctx* ctx) {
  free_ctx(ctx);
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_level_size(const fvvdp_b200_ctx* ctx, int level, int32_t* h, int32_t* w) {
  if (!ctx || level < 0 || level >= ctx->cfg.n_levels) return FVVDP_B200_ERR_INVALID;
  if (h) *h = ctx->lh[level];
  if (w) *w = ctx->lw[level];
  return FVVDP_B200_OK;
}

extern "C" int64_t fvvdp_b200_launch_count(const fvvdp_b200_ctx* ctx) { return ctx ? ctx->launches : 0; }

extern "C" int fvvdp_b200_traffic_model(const fvvdp_b200_ctx* ctx, double out_bytes[2]) {
  if (!ctx || !out_bytes) return FVVDP_B200_ERR_INVALID;
  out_bytes[0] = ctx->bytes_alg;
  out_bytes[1] = ctx->bytes_plan;
  return FVVDP_B200_OK;
}

// RAII bracket: records a start event now and a stop event when it goes out of scope (profiling only)
struct ProfScope {
  fvvdp_b200_ctx* c;
  cudaStream_t st;
  cudaEvent_t stop = nullptr;
  ProfScope(fvvdp_b200_ctx* c_, int cls, cudaStream_t st_) : c(c_), st(st_) {
    if (!c->profiling) return;
    if (c->ev_used * 2 >= c->ev.size()) {
      cudaEvent_t a, b;
      if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) return;
      c->ev.push_back(a); c->ev.push_back(b); c->ev_class.push_back(cls);
    }
    c->ev_class[c->ev_used] = cls;
    cudaEventRecord(c->ev[2 * c->ev_used], st);
    stop = c->ev[2 * c->ev_used + 1];
    c->ev_used++;
  }
  ~ProfScope() { if (stop) cudaEventRecord(stop, st); }
};

extern "C" int fvvdp_b200_profile(fvvdp_b200_ctx* ctx, int enable) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  ctx->profiling = enable != 0;
  ctx->ev_used = 0;
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_profile_read(fvvdp_b200_ctx* ctx, float* ms, int32_t* count) {
  if (!ctx || !ms || !count) return FVVDP_B200_ERR_INVALID;
  for (int i = 0; i < FVVDP_B200_PROFILE_CLASSES; ++i) { ms[i] = 0.f; count[i] = 0; }
  CU(cudaSetDevice(ctx->dev));
  if (ctx->ev_used > 0) CU(cudaEventSynchronize(ctx->ev[2 * ctx->ev_used - 1]));
  for (size_t i = 0; i < ctx->ev_used; ++i) {
    float t = 0.f;
    CU(cudaEventElapsedTime(&t, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
    ms[ctx->ev_class[i]] += t;
    count[ctx->ev_class[i]]++;
  }
  ctx->ev_used = 0;
  return FVVDP_B200_OK;
}

template <int FL, int PX, bool CONTIG, bool PAIRS = false>
static cudaError_t launch_front(const FrontParams& fp, cudaStream_t st) {
  const long long npx = (long long)fp.H * fp.W;
  const long long nthreads = (npx + PX - 1) / PX;
  const unsigned blocks = (unsigned)((nthreads + 255) / 256);
  front_kernel<FL, PX, CONTIG, PAIRS><<<blocks, 256, 0, st>>>(fp);
  return cudaGetLastError();
}

// gaze view direction [deg]: foveated == 1: pix2view_direction at frame resolution of fixation + 0.5 (fvvdp.py:429-431,
// fvvdp_display_model.py:498-510); foveated == 2 (custom geometry): the caller's plugin already converted it
static void gaze_direction(const fvvdp_b200_config& cfg, const float* fix, float out[2]) {
  if (cfg.foveated == 2) { out[0] = fix[0]; out[1] = fix[1]; return; }
  const float gx = fix[0] + 0.5f, gy = fix[1] + 0.5f;
  const float xm = (gx - (float)(cfg.width / 2.0)) * cfg.display_size_m[0] / (float)cfg.width;
  const float ym = -(gy - (float)(cfg.height / 2.0)) * cfg.display_size_m[1] / (float)cfg.height;
  out[0] = (float)(atan((double)(xm / cfg.distance_m)) * 180.0 / M_PI);
  out[1] = (float)(atan((double)(ym / cfg.distance_m)) * 180.0 / M_PI);
}

extern "C" int fvvdp_b200_set_foveation_maps(fvvdp_b200_ctx* ctx, int level, const float* view_xy, const float* log2_rho) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  if (ctx->cfg.foveated != 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "ctx was not created with foveated = 2 (custom display geometry)");
  if (level < 0 || level >= ctx->n_bands) return fail(ctx, FVVDP_B200_ERR_INVALID, "level %d is not a scored band", level);
  if (!view_xy || !log2_rho) return fail(ctx, FVVDP_B200_ERR_INVALID, "null map");
  ctx->fov_view[level] = view_xy;
  ctx->fov_rq[level] = log2_rho;
  return FVVDP_B200_OK;
}

static void fill_yuv_params(const fvvdp_b200_yuv_desc* d, YuvParams& p) {
  memset(&p, 0, sizeof(p));
  p.W = d->width; p.H = d->height;
  p.is420 = d->chroma_420 ? 1 : 0;
  p.cw = p.is420 ? d->width / 2 : d->width; p.ch = p.is420 ? d->height / 2 : d->height;
  p.is16 = d->bit_depth > 8;
  const float scale = (float)(1 << (d->bit_depth - 8));
  p.wy = 1.0f / (scale * 219.0f); p.oy = 16.0f / 219.0f;    // fixed2float, video_source_yuv.py:198-212
  p.wc = 1.0f / (scale * 224.0f); p.oc = 128.0f / 224.0f;
  for (int i = 0; i < 9; ++i) p.m[i] = d->ycbcr2rgb[i];
  p.eotf = d->eotf;
  p.Yscale = d->Y_peak - d->Y_black; p.Y_black = d->Y_black; p.Y_peak = d->Y_peak; p.gamma = d->gamma; p.L_min = d->L_min; p.L_max = d->L_max;
  for (int i = 0; i < 3; ++i) p.rgb2y[i] = d->rgb2y[i];
}

static const char* check_yuv_desc(const fvvdp_b200_yuv_desc* d) {
  if (d->width < 1 || d->height < 1) return "bad frame size";
  if (d->bit_depth < 8 || d->bit_depth > 16) return "bit depth not in 8..16";
  if (d->chroma_420 && ((d->width | d->height) & 1)) return "4:2:0 frames need an even width and height";
  if (d->eotf < 0 || d->eotf > 5) return "unknown EOTF";
  if (d->resize < FVVDP_B200_RESIZE_NONE || d->resize > FVVDP_B200_RESIZE_AREA) return "unknown resize mode";
  if (d->resize != FVVDP_B200_RESIZE_NONE && (d->out_width < 1 || d->out_height < 1)) return "bad output size";
  return nullptr;
}

static void fill_resize_params(const fvvdp_b200_yuv_desc* d, ResizeParams& r) {
  r.mode = d->resize; r.outW = d->out_width; r.outH = d->out_height;
  r.sx = (float)d->width / (float)d->out_width; r.sy = (float)d->height / (float)d->out_height;  // ATen area_pixel_compute_scale
}

// score_block for array frames (yuv == nullptr) or for DEVICE copies of raw planar Y'CbCr frames as stored in a .yuv file
static int score_block_impl(fvvdp_b200_ctx* ctx, const void* const* test_slots, const void* const* ref_slots, const int64_t strides[3],
                            int n_frames, const float* fixation_xy, float* q_out, int64_t q_stride, int64_t q_col0, uint32_t* flags_out,
                            void* cuda_stream, const fvvdp_b200_yuv_desc* yuv) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  const fvvdp_b200_config& cfg = ctx->cfg;
  if (!test_slots || !ref_slots || !strides || !q_out) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (n_frames < 1 || n_frames > ctx->T) return fail(ctx, FVVDP_B200_ERR_INVALID, "n_frames %d not in 1..%d", n_frames, ctx->T);
  if (q_col0 < 0 || q_col0 + n_frames > q_stride) return fail(ctx, FVVDP_B200_ERR_INVALID, "q_out columns out of range");
  if (cfg.foveated && !fixation_xy) return fail(ctx, FVVDP_B200_ERR_INVALID, "foveated scoring needs fixation points");
  if (cfg.foveated == 2)
    for (int l = 0; l < ctx->n_bands; ++l)
      if (!ctx->fov_view[l]) return fail(ctx, FVVDP_B200_ERR_INVALID, "custom display geometry: no foveation maps set for band %d", l);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CU(cudaSetDevice(ctx->dev));
  const int fl = cfg.filter_len, n_slots = n_frames + fl - 1;
  const int H = cfg.height, W = cfg.width;

  cudaError_t le = cudaSuccess;
  if (ctx->fused) {
    // ---- fused band kernels: one launch per pyramid level (fvvdp_fused.cuh) ----
    static_assert(sizeof(fused::BandParams) <= 16384, "kernel parameter space (32764 bytes since CUDA 12.1 on sm_70+)");
    fused::BandParams bp;
    memset(&bp, 0, sizeof(bp));
    bool aligned = true;
    for (int s = 0; s < n_slots; ++s) {
      if (!test_slots[s] || !ref_slots[s]) return fail(ctx, FVVDP_B200_ERR_INVALID, "null frame pointer in slot %d", s);
      bp.slot[0][s] = test_slots[s];
      bp.slot[1][s] = ref_slots[s];
      aligned = aligned && (((uintptr_t)test_slots[s] | (uintptr_t)ref_slots[s]) % 16 == 0);
    }
    const bool contig = !yuv && cfg.in_dtype == FVVDP_B200_F32 && cfg.in_channels == 1 && strides[2] == 1 && aligned && W % 4 == 0 &&
                        strides[1] % 4 == 0 && strides[1] >= W && strides[1] * (int64_t)H < (1ll << 31);
    const bool ch2 = ctx->ch2;
    const int mode = ch2 ? 3 : (cfg.temp_ch == 2 ? (fl > fused::RING ? 2 : 1) : 0);
    const int ring_len = mode == 2 ? fused::MAXRING : fused::RING;
    for (int cc = 0; cc < cfg.temp_ch; ++cc)
      for (int i = 0; i < 2 * ring_len; ++i) {
        const int age = i % ring_len;  // 0 = newest frame; cfg.filt[cc][0] weighs the newest (corr_filter = F.flip(0), fvvdp.py:298)
        const float wv = age < fl ? cfg.filt[cc][age] : 0.0f;
        uint32_t bits;
        memcpy(&bits, &wv, 4);
        bp.wext[cc][i] = ((unsigned long long)bits << 32) | bits;
      }
    // contiguous float frames at a common stride from one base address: level 0 is staged by TMA as well
    bool l0_tma = false;
    if (contig) {
      uintptr_t base[2], step[2] = {0, 0};
      const void* const* lists[2] = {test_slots, ref_slots};
      l0_tma = true;
      for (int st = 0; st < 2 && l0_tma; ++st) {
        uintptr_t lo = (uintptr_t)lists[st][0], hi = lo;
        for (int s = 1; s < n_slots; ++s) { const uintptr_t a = (uintptr_t)lists[st][s]; if (a < lo) lo = a; if (a > hi) hi = a; }
        uintptr_t g = 0;  // smallest positive distance to the base
        for (int s = 0; s < n_slots; ++s) { const uintptr_t d = (uintptr_t)lists[st][s] - lo; if (d && (!g || d < g)) g = d; }
        if (!g) g = (uintptr_t)strides[1] * H * 4;
        for (int s = 0; s < n_slots; ++s) {
          const uintptr_t d = (uintptr_t)lists[st][s] - lo;
          if (d % g || d / g > 65535) l0_tma = false; else bp.slot_frame[st][s] = (unsigned short)(d / g);
        }
        if (g % 16 || g < (uintptr_t)strides[1] * (H - 1) * 4 + (uintptr_t)W * 4) l0_tma = false;
        base[st] = lo; step[st] = g;
        if (l0_tma) {
          const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)((hi - lo) / g + 1)};
          const cuuint64_t str[2] = {(cuuint64_t)strides[1] * 4, (cuuint64_t)g};
          if (!make_tile_map(&bp.tmap[st], (const void*)lo, 3, dims, str, fused::LW)) l0_tma = false;
          if (l0_tma && ctx->ws_max_level >= 0 && !make_tile_map(&bp.tmap_ws[st], (const void*)lo, 3, dims, str, ws::LW, ctx->ws_th + 8)) l0_tma = false;
        }
      }
      (void)base; (void)step;
    }
    if (ch2) {
      // ---- temporal filters of the frames (register-ring walk, every input sample read and converted once) -> two (test, ref)
      //      planes per frame in the pyramid layout: P[0][temporal channel][frame]
      FrontParams fp;
      memset(&fp, 0, sizeof(fp));
      for (int s = 0; s < n_slots; ++s) { fp.slot[0][s] = test_slots[s]; fp.slot[1][s] = ref_slots[s]; }
      for (int cc = 0; cc < 2; ++cc)
        for (int k = 0; k < 32; ++k) {
          const int kk = k - (32 - fl);  // window position within the real filter, 0 = oldest
          fp.wgt[cc][k] = kk >= 0 ? cfg.filt[cc][fl - 1 - kk] : 0.0f;  // corr_filter = F.flip(0), fvvdp.py:298
        }
      fp.R = ctx->P[0]; fp.pitch = ctx->pitch[0]; fp.pairs_slots = ctx->T;
      fp.flags = flags_out;
      fp.sC = strides[0]; fp.sH = strides[1]; fp.sW = strides[2];
      fp.H = H; fp.W = W; fp.n_frames = n_frames; fp.fl = fl; fp.nch = 4; fp.C = cfg.in_channels;
      fp.dtype = cfg.in_dtype; fp.eotf = cfg.eotf;
      fp.Yscale = cfg.Y_peak - cfg.Y_black; fp.Y_black = cfg.Y_black; fp.Y_peak = cfg.Y_peak; fp.gamma = cfg.gamma;
      fp.L_min = cfg.L_min; fp.L_max = cfg.L_max;
      for (int i = 0; i < 3; ++i) fp.rgb2y[i] = cfg.rgb2y[i];
      const bool fcontig = cfg.in_dtype == FVVDP_B200_F32 && cfg.in_channels == 1 && strides[2] == 1 && aligned;
      ProfScope prof(ctx, 0, st);
      for (int cc = 0; cc < 2; ++cc)
        for (int a = 0; a < 32; ++a) fp.wage[cc][a] = a < fl ? cfg.filt[cc][a] : 0.0f;  // cfg.filt[cc][0] weighs the newest frame (corr_filter = F.flip(0), fvvdp.py:298)
      fp.ring_phase = (int)(((q_col0 - (fl - 1)) % 32 + 32) % 32);  // slot 0 is the frame shown at time q_col0 - (fl-1)
      {
        const long long npx = (long long)H * W;
        const unsigned blocks = (unsigned)((npx + 255) / 256);
        if (fl <= 16) {
          if (fcontig) front_pairs_kernel<true, 16><<<blocks, 256, 0, st>>>(fp);
          else front_pairs_kernel<false, 16><<<blocks, 256, 0, st>>>(fp);
        } else {
          if (fcontig) front_pairs_kernel<true, 32><<<blocks, 256, 0, st>>>(fp);
          else front_pairs_kernel<false, 32><<<blocks, 256, 0, st>>>(fp);
        }
      }
      cudaError_t lf = cudaGetLastError();
      if (lf != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "front_kernel launch: %s", cudaGetErrorString(lf));
      ctx->launches++;
      bp.ch2_slots_in = bp.ch2_slots_out = ctx->T;
    }
    bp.n_frames = n_frames; bp.fl = ch2 ? 1 : fl;
    bp.ring_phase = (int)(((q_col0 - (fl - 1)) % ring_len + ring_len) % ring_len);  // slot 0 is the frame shown at time q_col0 - (fl-1)
    bp.ring_phase_ws = (int)(((q_col0 - (fl - 1)) % ctx->ws_rp + ctx->ws_rp) % ctx->ws_rp);
    while (!ch2 && bp.dup_prefix + 1 < n_slots && test_slots[bp.dup_prefix + 1] == test_slots[0] && ref_slots[bp.dup_prefix + 1] == ref_slots[0]) bp.dup_prefix++;
    if (ctx->no_dup_skip) bp.dup_prefix = 0;  // A/B switch
    bp.sC = strides[0]; bp.sH = strides[1]; bp.sW = strides[2];
    bp.C = cfg.in_channels; bp.dtype = cfg.in_dtype; bp.eotf = cfg.eotf;
    bp.Yscale = cfg.Y_peak - cfg.Y_black; bp.Y_black = cfg.Y_black; bp.Y_peak = cfg.Y_peak; bp.gamma = cfg.gamma;
    bp.L_min = cfg.L_min; bp.L_max = cfg.L_max;
    for (int i = 0; i < 3; ++i) bp.rgb2y[i] = cfg.rgb2y[i];
    bp.flags = flags_out;
    bp.y0 = ctx->ax.x0[1]; bp.inv_dy = ctx->ax.inv_dx[1]; bp.lg_y_hi = log2f(cfg.csf_Y_range[1]);
    bp.mask_p = cfg.mask_p; bp.mask_q[0] = cfg.mask_q[0]; bp.mask_q[1] = cfg.mask_q[1];
    bp.log2_mask_c = log2f(cfg.mask_c_mul); bp.beta = cfg.beta; bp.w_transient = cfg.w_transient;
    {
      auto pack2 = [](float lo, float hi) { uint32_t a, b; memcpy(&a, &lo, 4); memcpy(&b, &hi, 4); return ((unsigned long long)b << 32) | a; };
      bp.m_q = pack2(cfg.mask_q[0], cfg.mask_q[1]);
      bp.m_qlmc = pack2(cfg.mask_q[0] * bp.log2_mask_c, cfg.mask_q[1] * bp.log2_mask_c);
      bp.m_bp = pack2(cfg.beta * cfg.mask_p, cfg.beta * cfg.mask_p);
      bp.m_nbeta = pack2(-cfg.beta, -cfg.beta);
      bp.m_cap = cfg.beta * 13.287712379549449f;
    }
    bp.ax = ctx->ax; bp.lut4 = ctx->lut4; bp.log2_sens_mul = ctx->log2_sens_mul;
    if (cfg.foveated) {
      const double delta = (1.0 / cfg.ppd_centre) / 2.0 * M_PI / 180.0;
      bp.res_k0 = (float)cos(delta);
      bp.res_delta_rad = (float)delta;
      for (int i = 0; i < n_frames; ++i) gaze_direction(cfg, fixation_xy + 2 * i, bp.gaze[i]);
    }
    const bool extra = cfg.want_taps || cfg.want_dmap;
    for (int l = 0; l < ctx->n_bands; ++l) {
      bp.Pn = (l + 1 < ctx->n_bands) ? ctx->P[l + 1] : nullptr;
      bp.pitch2 = ctx->pitch[l + 1];
      bp.Pn_slot_stride = (long long)ctx->lh[l + 1] * ctx->pitch[l + 1];
      bp.partial = ctx->partial[l];
      bp.h = ctx->lh[l]; bp.w = ctx->lw[l]; bp.h2 = ctx->lh[l + 1]; bp.w2 = ctx->lw[l + 1];
      bp.h_odd = bp.h & 1;
      // the warp-specialised kernel (one CTA of 24 warps per SM, 32x64 tiles) where it applies; it stages with TMA only
      const int ws_tx = (bp.w + ws::TW - 1) / ws::TW, ws_ty = (bp.h + ctx->ws_th - 1) / ctx->ws_th;
      const bool use_ws = l <= ctx->ws_max_level && (l > 0 || l0_tma || !contig) && cfg.foveated != 2 &&  // custom geometry maps: fused kernel
                          (l == 0 || ws_tx * ws_ty >= ctx->ws_min_tiles);
      const int tx = use_ws ? ws_tx : ctx->tiles_x[l], ty = use_ws ? ws_ty : ctx->tiles_y[l];
      const int tiles = tx * ty;
      bp.ntiles = tiles;
      ctx->ntiles_used[l] = use_ws ? tiles * 16 : tiles;
      // small levels: split the time walk so that the grid still fills the machine (each chunk re-walks fl-1 frames)
      int nchunks = ((use_ws ? 1 : 4) * 148 + tiles - 1) / tiles;
      const int max_chunks = (n_frames + 3) / 4;
      if (nchunks > max_chunks) nchunks = max_chunks;
      if (nchunks < 1) nchunks = 1;
      bp.chunk = (n_frames + nchunks - 1) / nchunks;
      nchunks = (n_frames + bp.chunk - 1) / bp.chunk;
      bp.cell = ctx->cell + (size_t)l * 256;
      bp.band_mul = (l == 0) ? 1.0f : 2.0f;  // get_band, fvvdp_lpyr_dec.py:57-63 (the base band is never scored)
      bp.log2_m = (l == 0) ? 0.0f : 1.0f;
      bp.rho_band = cfg.band_freq[l];
      bp.vx = ctx->vx[l]; bp.vy = ctx->vy[l];
      bp.vmap = ctx->fov_view[l]; bp.rqmap = ctx->fov_rq[l];
      bp.ctxmap = (l == 0) ? ctx->ctxmap : nullptr;
      bp.tapR = (l == 0) ? ctx->G[0] : nullptr;
      bp.tapG = cfg.want_taps ? ctx->G[l + 1] : nullptr;
      bp.tapC = ctx->tapC[l]; bp.tapL = ctx->tapL[l]; bp.tapS = ctx->tapS[l]; bp.tapD = ctx->tapD[l];
      bp.dmap = ctx->dmap[l];
      int kind = l == 0 && !ch2 ? (l0_tma ? fused::IN_LEVEL0_TMA : fused::IN_LEVEL0_CPASYNC) : fused::IN_PYRAMID_TMA;
      if (l == 0 && ch2) { bp.tmap[0] = ctx->pmap[0]; bp.tmap_ws[0] = ctx->pmap_ws[0]; }
      if (l == 0 && !contig && !ch2) {
        // any other input format: one luminance pass into planes laid out like the pyramid, then level 0 is TMA-staged too
        if (!ctx->P[0]) {
          const size_t n = (size_t)(ctx->T + cfg.filter_len - 1) * H * ctx->pitch[0];
          CU(cudaMalloc(&ctx->P[0], sizeof(float) * n));
          CU(cudaMemsetAsync(ctx->P[0], 0, sizeof(float) * n, st));
          const cuuint64_t dims[3] = {(cuuint64_t)(2 * W), (cuuint64_t)H, (cuuint64_t)(ctx->T + cfg.filter_len - 1)};
          const cuuint64_t str[2] = {(cuuint64_t)ctx->pitch[0] * 4, (cuuint64_t)H * ctx->pitch[0] * 4};
          if (!make_tile_map(&ctx->pmap[0], ctx->P[0], 3, dims, str, 2 * fused::LW) ||
              !make_tile_map(&ctx->pmap_ws[0], ctx->P[0], 3, dims, str, 2 * ws::LW, ctx->ws_th + 8)) return fail(ctx, FVVDP_B200_ERR_CUDA, "cuTensorMapEncodeTiled failed for the luminance planes");
        }
        if (yuv) {
          // planar Y'CbCr frames: unpack, chroma upsampling, Y'CbCr -> R'G'B', display EOTF, RGB -> Y for the window slots of both
          // streams in one launch, straight into the (test, reference) planes level 0 stages by TMA
          ProfScope prof(ctx, 0, st);
          YuvBlockParams yp;
          memset(&yp, 0, sizeof(yp));
          fill_yuv_params(yuv, yp.f);
          for (int s = 0; s < n_slots; ++s) {
            yp.frame[0][s] = test_slots[s]; yp.frame[1][s] = ref_slots[s];
            yp.skip[s] = 0;  // (repeats of slot 0 are converted too: time chunks of small frames start their walk inside them)
          }
          yp.y_elems = (long long)yp.f.W * yp.f.H; yp.c_elems = (long long)yp.f.cw * yp.f.ch;
          yp.out = ctx->P[0]; yp.slot_stride = (long long)H * ctx->pitch[0]; yp.pitch = ctx->pitch[0];
          dim3 yg(((W + 1) / 2 + 31) / 32, (H + 7) / 8, n_slots);
          if (yuv->resize) {
            // full-screen resize: every output pixel gathers its taps from the planar frames (yuv_resize_planes_kernel)
            ResizeParams rs;
            fill_resize_params(yuv, rs);
            dim3 rg((W + 31) / 32, (H + 7) / 8, n_slots);
            switch (yuv->eotf) {
              case FVVDP_B200_EOTF_NONE: yuv_resize_planes_kernel<FVVDP_B200_EOTF_NONE><<<rg, 256, 0, st>>>(yp, rs); break;
              case FVVDP_B200_EOTF_SRGB: yuv_resize_planes_kernel<FVVDP_B200_EOTF_SRGB><<<rg, 256, 0, st>>>(yp, rs); break;
              case FVVDP_B200_EOTF_GAMMA: yuv_resize_planes_kernel<FVVDP_B200_EOTF_GAMMA><<<rg, 256, 0, st>>>(yp, rs); break;
              case FVVDP_B200_EOTF_PQ: yuv_resize_planes_kernel<FVVDP_B200_EOTF_PQ><<<rg, 256, 0, st>>>(yp, rs); break;
              case FVVDP_B200_EOTF_LINEAR: yuv_resize_planes_kernel<FVVDP_B200_EOTF_LINEAR><<<rg, 256, 0, st>>>(yp, rs); break;
              default: yuv_resize_planes_kernel<FVVDP_B200_EOTF_ABSOLUTE><<<rg, 256, 0, st>>>(yp, rs); break;
            }
          } else
          switch (yuv->eotf) {
            case FVVDP_B200_EOTF_NONE: yuv_planes_kernel<FVVDP_B200_EOTF_NONE><<<yg, 256, 0, st>>>(yp); break;
            case FVVDP_B200_EOTF_SRGB: yuv_planes_kernel<FVVDP_B200_EOTF_SRGB><<<yg, 256, 0, st>>>(yp); break;
            case FVVDP_B200_EOTF_GAMMA: yuv_planes_kernel<FVVDP_B200_EOTF_GAMMA><<<yg, 256, 0, st>>>(yp); break;
            case FVVDP_B200_EOTF_PQ: yuv_planes_kernel<FVVDP_B200_EOTF_PQ><<<yg, 256, 0, st>>>(yp); break;
            case FVVDP_B200_EOTF_LINEAR: yuv_planes_kernel<FVVDP_B200_EOTF_LINEAR><<<yg, 256, 0, st>>>(yp); break;
            default: yuv_planes_kernel<FVVDP_B200_EOTF_ABSOLUTE><<<yg, 256, 0, st>>>(yp); break;
          }
          cudaError_t le1 = cudaGetLastError();
          if (le1 != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "yuv_planes_kernel launch: %s", cudaGetErrorString(le1));
          ctx->launches++;
        } else {
          ProfScope prof(ctx, 0, st);
          // 4 pixels per thread with vector loads when rows are contiguous and every row / plane / frame starts aligned
          const bool vec_ok = strides[2] == 1 && aligned && strides[1] % 4 == 0 && (cfg.in_channels == 1 || strides[0] % 4 == 0);
          cudaError_t le1 = fused::launch_luminance(bp, ctx->P[0], (long long)H * ctx->pitch[0], ctx->pitch[0], n_slots, vec_ok, st);
          if (le1 != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "luminance_kernel launch: %s", cudaGetErrorString(le1));
          ctx->launches++;
        }
        kind = fused::IN_PYRAMID_TMA;
        bp.tmap[0] = ctx->pmap[0];
        bp.tmap_ws[0] = ctx->pmap_ws[0];
      }
      if (l >= 1) { bp.tmap[0] = ctx->pmap[l]; bp.tmap_ws[0] = ctx->pmap_ws[l]; }
      dim3 grid(tx, ty, nchunks);
      ProfScope prof(ctx, 1 + l, st);
      cudaError_t le2 = use_ws ? (ctx->ws_rp == ws::RP ? ws::launch_band_ws(kind, cfg.foveated != 0, bp, grid, st)
                                                        : ws16::launch_band_ws(kind, cfg.foveated != 0, bp, grid, st))
                               : fused::launch_band(kind, mode, cfg.foveated != 0, extra, bp, grid, st);
      if (le2 != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "band_kernel[%d] launch: %s", l, cudaGetErrorString(le2));
      ctx->launches++;
    }
  } else {
  if (yuv) return fail(ctx, FVVDP_B200_ERR_INVALID, "raw .yuv frame blocks need a temporal window of at most %d taps", fused::MAXRING);
  // ---- K_front ----
  FrontParams fp;
  memset(&fp, 0, sizeof(fp));
  bool aligned = true;
  for (int s = 0; s < n_slots; ++s) {
    if (!test_slots[s] || !ref_slots[s]) return fail(ctx, FVVDP_B200_ERR_INVALID, "null frame pointer in slot %d", s);
    fp.slot[0][s] = test_slots[s];
    fp.slot[1][s] = ref_slots[s];
    aligned = aligned && (((uintptr_t)test_slots[s] | (uintptr_t)ref_slots[s]) % 16 == 0);
  }
  const int FLT = fl <= 8 ? 8 : (fl <= 16 ? 16 : 32);
  for (int cc = 0; cc < cfg.temp_ch; ++cc)
    for (int k = 0; k < FLT; ++k) {
      const int kk = k - (FLT - fl);  // window position within the real filter, 0 = oldest
      fp.wgt[cc][k] = kk >= 0 ? cfg.filt[cc][fl - 1 - kk] : 0.0f;  // corr_filter = F.flip(0), fvvdp.py:298
    }
  fp.R = ctx->G[0];
  fp.flags = flags_out;
  fp.sC = strides[0]; fp.sH = strides[1]; fp.sW = strides[2];
  fp.H = H; fp.W = W; fp.n_frames = n_frames; fp.fl = fl; fp.nch = ctx->nch; fp.C = cfg.in_channels;
  fp.dtype = cfg.in_dtype; fp.eotf = cfg.eotf;
  fp.Yscale = cfg.Y_peak - cfg.Y_black; fp.Y_black = cfg.Y_black; fp.Y_peak = cfg.Y_peak; fp.gamma = cfg.gamma;
  fp.L_min = cfg.L_min; fp.L_max = cfg.L_max;
  for (int i = 0; i < 3; ++i) fp.rgb2y[i] = cfg.rgb2y[i];
  const bool contig = cfg.in_dtype == FVVDP_B200_F32 && cfg.in_channels == 1 && strides[2] == 1 && aligned;
  {
  ProfScope prof(ctx, 0, st);
  if (FLT == 8) {
    if (contig && W % 4 == 0 && strides[1] % 4 == 0) le = launch_front<8, 4, true>(fp, st);
    else le = launch_front<8, 1, false>(fp, st);
  } else if (FLT == 16) {
    if (contig && W % 2 == 0 && strides[1] % 2 == 0) le = launch_front<16, 2, true>(fp, st);
    else le = launch_front<16, 1, false>(fp, st);
  } else {
    if (contig) le = launch_front<32, 1, true>(fp, st);
    else le = launch_front<32, 1, false>(fp, st);
  }
  }
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "front_kernel launch: %s", cudaGetErrorString(le));
  ctx->launches++;

  // ---- K_level per band ----
  LevelParams lp;  // POD incl. the per-frame gaze table; filled per launch, passed by value
  for (int l = 0; l < ctx->n_bands; ++l) {
    memset(&lp, 0, sizeof(lp));
    lp.G = ctx->G[l];
    lp.Gn = ctx->G[l + 1];  // nullptr for the base level unless taps are kept
    lp.partial = ctx->partial[l];
    lp.h = ctx->lh[l]; lp.w = ctx->lw[l]; lp.h2 = ctx->lh[l + 1]; lp.w2 = ctx->lw[l + 1];
    // gausspyr_reduce keys the last-column term on the ROW count (fvvdp_lpyr_dec.py:202)
    const bool h_odd = lp.h & 1, w_odd = lp.w & 1;
    lp.quirk = (h_odd && !w_odd) ? 1 : ((!h_odd && w_odd) ? 2 : 0);
    lp.ntiles = ctx->tiles_x[l] * ctx->tiles_y[l];
    ctx->ntiles_used[l] = lp.ntiles;
    lp.band_mul = (l == 0) ? 1.0f : 2.0f;  // get_band, fvvdp_lpyr_dec.py:57-63 (the base band is never scored)
    lp.rho_band = cfg.band_freq[l];
    lp.ax = ctx->ax;
    lp.csf1d = ctx->csf1d + (size_t)l * 64;
    lp.lut3d = ctx->lut3d;
    lp.log2_sens_mul = ctx->log2_sens_mul;
    lp.mask_p = cfg.mask_p; lp.mask_q[0] = cfg.mask_q[0]; lp.mask_q[1] = cfg.mask_q[1];
    lp.mask_c_mul = cfg.mask_c_mul; lp.beta = cfg.beta; lp.w_transient = cfg.w_transient;
    if (cfg.foveated) {
      lp.vx = ctx->vx[l]; lp.vy = ctx->vy[l];
      // ppd(a)/ppd_c = (tan(a+d) - tan a)/tan d = cos d / (cos a cos(a+d)), d = half a central pixel
      const double delta = (1.0 / cfg.ppd_centre) / 2.0 * M_PI / 180.0;
      lp.res_k0 = (float)cos(delta);
      lp.res_delta_rad = (float)delta;
      lp.vmap = ctx->fov_view[l]; lp.rqmap = ctx->fov_rq[l];
      for (int i = 0; i < n_frames; ++i) gaze_direction(cfg, fixation_xy + 2 * i, lp.gaze[i]);
    }
    lp.tapC = ctx->tapC[l]; lp.tapL = ctx->tapL[l]; lp.tapS = ctx->tapS[l]; lp.tapD = ctx->tapD[l];
    lp.dmap = ctx->dmap[l];
    dim3 grid(ctx->tiles_x[l], ctx->tiles_y[l], n_frames);
    const size_t smem = level_smem_bytes(ctx->nch);
    ProfScope prof(ctx, 1 + l, st);
    if (ctx->nch == 4) {
      if (cfg.foveated) level_kernel<4, true><<<grid, LEVEL_THREADS, smem, st>>>(lp);
      else level_kernel<4, false><<<grid, LEVEL_THREADS, smem, st>>>(lp);
    } else {
      if (cfg.foveated) level_kernel<2, true><<<grid, LEVEL_THREADS, smem, st>>>(lp);
      else level_kernel<2, false><<<grid, LEVEL_THREADS, smem, st>>>(lp);
    }
    le = cudaGetLastError();
    if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "level_kernel[%d] launch: %s", l, cudaGetErrorString(le));
    ctx->launches++;
  }

  }  // v1 path

  // ---- K_final ----
  FinalParams fin;
  memset(&fin, 0, sizeof(fin));
  for (int l = 0; l < ctx->n_bands; ++l) {
    fin.partial[l] = ctx->partial[l];
    fin.ntiles[l] = ctx->ntiles_used[l];
    fin.npix[l] = (double)ctx->lh[l] * ctx->lw[l];
  }
  fin.q_out = q_out; fin.q_stride = q_stride; fin.q_col0 = q_col0;
  fin.n_bands = ctx->n_bands; fin.n_frames = n_frames; fin.temp_ch = cfg.temp_ch;
  fin.inv_beta = 1.0 / (double)cfg.beta;
  {
    ProfScope prof(ctx, FVVDP_B200_MAX_LEVELS + 1, st);
    final_kernel<<<n_frames * ctx->n_bands * 2, 256, 0, st>>>(fin);
  }
  le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "final_kernel launch: %s", cudaGetErrorString(le));
  ctx->launches++;
  ctx->last_n_frames = n_frames;

  // traffic model: compulsory input bytes; bytes this kernel plan moves through global memory
  const double esz = cfg.in_dtype == FVVDP_B200_F32 ? 4.0 : (cfg.in_dtype == FVVDP_B200_U8 ? 1.0 : 2.0);
  const double P0 = (double)H * W;
  ctx->bytes_alg = 2.0 * P0 * cfg.in_channels * esz * n_frames;
  double plan = 2.0 * P0 * cfg.in_channels * esz * n_slots + 4.0 * P0 * ctx->nch * n_frames;
  for (int l = 0; l < ctx->n_bands; ++l) {
    plan += 4.0 * ctx->nch * n_frames * (double)ctx->lh[l] * ctx->lw[l];
    if (l + 1 < ctx->n_bands) plan += 4.0 * ctx->nch * n_frames * (double)ctx->lh[l + 1] * ctx->lw[l + 1];
  }
  if (ctx->fused) {  // inputs once per window slot + write and read-back of the 2-plane luminance pyramid
    plan = 2.0 * P0 * cfg.in_channels * esz * n_slots;
    for (int l = 1; l < ctx->n_bands; ++l) plan += 2.0 * 4.0 * n_slots * (double)ctx->lh[l] * ctx->pitch[l];
  }
  ctx->bytes_plan = plan;
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_score_block(fvvdp_b200_ctx* ctx, const void* const* test_slots, const void* const* ref_slots,
                                      const int64_t strides[3], int n_frames, const float* fixation_xy, float* q_out,
                                      int64_t q_stride, int64_t q_col0, uint32_t* flags_out, void* cuda_stream) {
  return score_block_impl(ctx, test_slots, ref_slots, strides, n_frames, fixation_xy, q_out, q_stride, q_col0, flags_out, cuda_stream, nullptr);
}

extern "C" int fvvdp_b200_score_block_yuv(fvvdp_b200_ctx* ctx, const fvvdp_b200_yuv_desc* desc, const void* const* test_frames,
                                          const void* const* ref_frames, int n_frames, const float* fixation_xy, float* q_out,
                                          int64_t q_stride, int64_t q_col0, void* cuda_stream) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  if (!desc) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (const char* why = check_yuv_desc(desc)) return fail(ctx, FVVDP_B200_ERR_INVALID, "yuv frame description: %s", why);
  const int ow = desc->resize ? desc->out_width : desc->width, oh = desc->resize ? desc->out_height : desc->height;
  if (ow != ctx->cfg.width || oh != ctx->cfg.height) return fail(ctx, FVVDP_B200_ERR_INVALID, "frame size %dx%d does not match the context (%dx%d)", ow, oh, ctx->cfg.width, ctx->cfg.height);
  const int64_t strides[3] = {0, ctx->cfg.width, 1};
  return score_block_impl(ctx, test_frames, ref_frames, strides, n_frames, fixation_xy, q_out, q_stride, q_col0, nullptr, cuda_stream, desc);
}

extern "C" int64_t fvvdp_b200_read_tap(fvvdp_b200_ctx* ctx, int tap, int level, int frame, float* dst, int64_t cap, void* cuda_stream) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  if (!dst) return fail(ctx, FVVDP_B200_ERR_INVALID, "null destination");
  if (frame < 0 || frame >= ctx->last_n_frames) return fail(ctx, FVVDP_B200_ERR_INVALID, "frame %d not in the last block", frame);
  if (tap == FVVDP_B200_TAP_R) level = 0;
  if (level < 0 || level >= ctx->cfg.n_levels) return fail(ctx, FVVDP_B200_ERR_INVALID, "level out of range");
  const size_t px = (size_t)ctx->lh[level] * ctx->lw[level];
  const float* src = nullptr;
  size_t n = 0;
  switch (tap) {
    case FVVDP_B200_TAP_R:
    case FVVDP_B200_TAP_GAUSS: src = ctx->G[level]; n = px * ctx->nch; break;
    case FVVDP_B200_TAP_CONTRAST: src = level < ctx->n_bands ? ctx->tapC[level] : nullptr; n = px * ctx->nch; break;
    case FVVDP_B200_TAP_LBKG: src = level < ctx->n_bands ? ctx->tapL[level] : nullptr; n = px; break;
    case FVVDP_B200_TAP_S: src = level < ctx->n_bands ? ctx->tapS[level] : nullptr; n = px * ctx->cfg.temp_ch; break;
    case FVVDP_B200_TAP_D: src = level < ctx->n_bands ? ctx->tapD[level] : nullptr; n = px * ctx->cfg.temp_ch; break;
    case FVVDP_B200_TAP_DMAP_BAND: src = level < ctx->n_bands ? ctx->dmap[level] : nullptr; n = px; break;
    default: return fail(ctx, FVVDP_B200_ERR_INVALID, "unknown tap %d", tap);
  }
  if (!src) return fail(ctx, FVVDP_B200_ERR_INVALID, "tap %d level %d not kept (create the ctx with want_taps / want_dmap)", tap, level);
  if ((int64_t)n > cap) return fail(ctx, FVVDP_B200_ERR_INVALID, "destination too small (%lld < %lld floats)", (long long)cap, (long long)n);
  CU(cudaSetDevice(ctx->dev));
  CU(cudaMemcpyAsync(dst, src + (size_t)frame * n, n * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)cuda_stream));
  return (int64_t)n;
}

// reconstruct (fvvdp_lpyr_dec.py:94-101) with a zero base band: coarse -> fine, expand + add.  The level-0 result goes to
// out16 as |jod_a| recon^beta_jod, or (out16 == nullptr) stays in ctx->recon[0] as the plain reconstruction.
static int reconstruct_dmap(fvvdp_b200_ctx* ctx, int frame, float beta_jod, float jod_a_abs, __half* out16, cudaStream_t st) {
  const float* coarse = nullptr;
  int ch = 0, cw = 0;
  for (int l = ctx->n_bands - 1; l >= 0; --l) {
    const int h = ctx->lh[l], w = ctx->lw[l];
    const float* band = ctx->dmap[l] + (size_t)frame * h * w;
    float* out = ctx->recon[l & 1];
    dim3 grid((w + 31) / 32, (h + 7) / 8);
    recon_kernel<<<grid, 256, 0, st>>>(coarse, ch, cw, band, out, l == 0 ? out16 : nullptr, h, w, beta_jod, jod_a_abs);
    cudaError_t le = cudaGetLastError();
    if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "recon_kernel launch: %s", cudaGetErrorString(le));
    ctx->launches++;
    coarse = out; ch = h; cw = w;
  }
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_heatmap(fvvdp_b200_ctx* ctx, int frame, float beta_jod, float jod_a_abs, void* dmap_out_f16, void* cuda_stream) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  if (!ctx->cfg.want_dmap) return fail(ctx, FVVDP_B200_ERR_INVALID, "ctx was created without want_dmap");
  if (!dmap_out_f16) return fail(ctx, FVVDP_B200_ERR_INVALID, "null destination");
  if (frame < 0 || frame >= ctx->last_n_frames) return fail(ctx, FVVDP_B200_ERR_INVALID, "frame %d not in the last block", frame);
  CU(cudaSetDevice(ctx->dev));
  return reconstruct_dmap(ctx, frame, beta_jod, jod_a_abs, (__half*)dmap_out_f16, (cudaStream_t)cuda_stream);
}

extern "C" int fvvdp_b200_heatmap_visualize(fvvdp_b200_ctx* ctx, int frame, float beta_jod, float jod_a_abs, int colormap, void* rgb_out_f16,
                                            void* cuda_stream) {
  if (!ctx) return FVVDP_B200_ERR_INVALID;
  if (ctx->cfg.want_dmap < 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "ctx was created without want_dmap = 2");
  if (!rgb_out_f16) return fail(ctx, FVVDP_B200_ERR_INVALID, "null destination");
  if (frame < 0 || frame >= ctx->last_n_frames) return fail(ctx, FVVDP_B200_ERR_INVALID, "frame %d not in the last block", frame);
  // colour maps of visualize_diff_map.py:66-82
  static const float kThr[5][3] = {{0.2f, 0.2f, 1.0f}, {0.2f, 1.0f, 1.0f}, {0.2f, 1.0f, 0.2f}, {1.0f, 1.0f, 0.2f}, {1.0f, 0.2f, 0.2f}};
  static const float kSup[3][3] = {{0.2f, 1.0f, 1.0f}, {1.0f, 1.0f, 1.0f}, {1.0f, 1.0f, 0.2f}};
  VisColorMap cm;
  memset(&cm, 0, sizeof(cm));
  const float (*src)[3];
  if (colormap == FVVDP_B200_CMAP_THRESHOLD) { cm.n = 5; src = kThr; }
  else if (colormap == FVVDP_B200_CMAP_SUPRA_THRESHOLD) { cm.n = 3; src = kSup; }
  else return fail(ctx, FVVDP_B200_ERR_INVALID, "Unknown colormap: %d", colormap);
  for (int i = 0; i < cm.n; ++i) {
    cm.in[i] = (float)i / (float)(cm.n - 1);
    const float lum = src[i][0] * 0.212656f + src[i][1] * 0.715158f + src[i][2] * 0.072186f;
    for (int c = 0; c < 3; ++c) cm.ch[i][c] = src[i][c] / (lum + 0.0001f);
  }
  cudaStream_t st = (cudaStream_t)cuda_stream;
  CU(cudaSetDevice(ctx->dev));
  int rc = reconstruct_dmap(ctx, frame, beta_jod, jod_a_abs, nullptr, st);
  if (rc != FVVDP_B200_OK) return rc;
  const long long n = (long long)ctx->cfg.height * ctx->cfg.width;
  // context image = R[:,0], the sustained (or, for an image, the only) channel of the TEST stream (fvvdp.py:475)
  const float* y = ctx->ctxmap ? ctx->ctxmap + (size_t)frame * n : ctx->G[0] + (size_t)frame * ctx->nch * n;
  const unsigned blocks = (unsigned)((n + 255) / 256), rblocks = blocks < 1184u ? blocks : 1184u;
  vis_reset_kernel<<<1, 1024, 0, st>>>(ctx->vis);
  vis_range_kernel<<<rblocks, 256, 0, st>>>(y, n, ctx->vis);
  vis_hist_kernel<<<rblocks, 256, 0, st>>>(y, n, ctx->vis);
  vis_curve_kernel<<<1, 1024, 0, st>>>(ctx->vis, (float)n);
  vis_apply_kernel<<<blocks, 256, 0, st>>>(ctx->recon[0], y, n, ctx->vis, cm, beta_jod, jod_a_abs, (__half*)rgb_out_f16);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "heat-map visualisation launch: %s", cudaGetErrorString(le));
  ctx->launches += 5;
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_yuv_to_luminance(const fvvdp_b200_yuv_desc* d, const void* y_plane, const void* u_plane, const void* v_plane,
                                           float* lum_out, float* rgb_out, int cuda_device, void* cuda_stream) {
  fvvdp_b200_ctx* ctx = nullptr;  // ctx-free: errors are reported through fvvdp_b200_last_error(NULL)
  if (!d || !y_plane || !u_plane || !v_plane || (!lum_out && !rgb_out)) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (const char* why = check_yuv_desc(d)) return fail(ctx, FVVDP_B200_ERR_INVALID, "yuv frame description: %s", why);
  CU(cudaSetDevice(cuda_device));
  YuvParams p;
  fill_yuv_params(d, p);
  p.y = y_plane; p.u = u_plane; p.v = v_plane;
  p.lum = lum_out; p.rgb = rgb_out;
  dim3 grid(((d->width + 1) / 2 + 31) / 32, (d->height + 7) / 8);
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (d->resize) {
    ResizeParams rs;
    fill_resize_params(d, rs);
    dim3 rg((rs.outW + 31) / 32, (rs.outH + 7) / 8);
    switch (lum_out ? d->eotf : FVVDP_B200_EOTF_NONE) {
      case FVVDP_B200_EOTF_NONE: yuv_resize_kernel<FVVDP_B200_EOTF_NONE><<<rg, 256, 0, st>>>(p, rs); break;
      case FVVDP_B200_EOTF_SRGB: yuv_resize_kernel<FVVDP_B200_EOTF_SRGB><<<rg, 256, 0, st>>>(p, rs); break;
      case FVVDP_B200_EOTF_GAMMA: yuv_resize_kernel<FVVDP_B200_EOTF_GAMMA><<<rg, 256, 0, st>>>(p, rs); break;
      case FVVDP_B200_EOTF_PQ: yuv_resize_kernel<FVVDP_B200_EOTF_PQ><<<rg, 256, 0, st>>>(p, rs); break;
      case FVVDP_B200_EOTF_LINEAR: yuv_resize_kernel<FVVDP_B200_EOTF_LINEAR><<<rg, 256, 0, st>>>(p, rs); break;
      default: yuv_resize_kernel<FVVDP_B200_EOTF_ABSOLUTE><<<rg, 256, 0, st>>>(p, rs); break;
    }
  } else
  switch (lum_out ? d->eotf : FVVDP_B200_EOTF_NONE) {
    case FVVDP_B200_EOTF_NONE: yuv_kernel<FVVDP_B200_EOTF_NONE><<<grid, 256, 0, st>>>(p); break;
    case FVVDP_B200_EOTF_SRGB: yuv_kernel<FVVDP_B200_EOTF_SRGB><<<grid, 256, 0, st>>>(p); break;
    case FVVDP_B200_EOTF_GAMMA: yuv_kernel<FVVDP_B200_EOTF_GAMMA><<<grid, 256, 0, st>>>(p); break;
    case FVVDP_B200_EOTF_PQ: yuv_kernel<FVVDP_B200_EOTF_PQ><<<grid, 256, 0, st>>>(p); break;
    case FVVDP_B200_EOTF_LINEAR: yuv_kernel<FVVDP_B200_EOTF_LINEAR><<<grid, 256, 0, st>>>(p); break;
    default: yuv_kernel<FVVDP_B200_EOTF_ABSOLUTE><<<grid, 256, 0, st>>>(p); break;
  }
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "yuv_kernel launch: %s", cudaGetErrorString(le));
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_pu_sq_err(const float* lum_test, const float* lum_ref, int64_t n, const fvvdp_b200_pu_params* params,
                                    double* sq_err_acc, int cuda_device, void* cuda_stream) {
  fvvdp_b200_ctx* ctx = nullptr;  // ctx-free: errors are reported through fvvdp_b200_last_error(NULL)
  if (!lum_test || !lum_ref || !params || !sq_err_acc) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (n < 1) return fail(ctx, FVVDP_B200_ERR_INVALID, "empty frame");
  CU(cudaSetDevice(cuda_device));
  PuParams q;
  for (int i = 0; i < 7; ++i) q.p[i] = params->p[i];
  q.L_min = params->L_min; q.L_max = params->L_max;
  const long long blocks = (n + 255) / 256;
  pu_sqerr_kernel<<<(unsigned)(blocks < 1184 ? blocks : 1184), 256, 0, (cudaStream_t)cuda_stream>>>(lum_test, lum_ref, (long long)n, q, sq_err_acc);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "pu_sqerr_kernel launch: %s", cudaGetErrorString(le));
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_pu_sq_err_frames(const fvvdp_b200_frame_format* fmt, const void* const* test_frames, const void* const* ref_frames,
                                           const int64_t strides[3], int n_frames, const fvvdp_b200_pu_params* params, double* sq_err_out,
                                           int cuda_device, void* cuda_stream) {
  fvvdp_b200_ctx* ctx = nullptr;  // ctx-free: errors are reported through fvvdp_b200_last_error(NULL)
  if (!fmt || !test_frames || !ref_frames || !strides || !params || !sq_err_out) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (n_frames < 1 || n_frames > FVVDP_B200_MAX_SLOTS) return fail(ctx, FVVDP_B200_ERR_INVALID, "n_frames %d not in 1..%d", n_frames, FVVDP_B200_MAX_SLOTS);
  if (fmt->width < 1 || fmt->height < 1) return fail(ctx, FVVDP_B200_ERR_INVALID, "bad frame size %dx%d", fmt->width, fmt->height);
  if (fmt->in_channels != 1 && fmt->in_channels != 3) return fail(ctx, FVVDP_B200_ERR_INVALID, "The content must have either 1 or 3 colour channels.");
  if (fmt->in_dtype < 0 || fmt->in_dtype > 2) return fail(ctx, FVVDP_B200_ERR_INVALID, "Only uint8, uint16 and float32 is currently supported");
  if (fmt->eotf < 0 || fmt->eotf > 5) return fail(ctx, FVVDP_B200_ERR_INVALID, "Unknown EOTF %d", fmt->eotf);
  CU(cudaSetDevice(cuda_device));
  fused::BandParams bp;
  memset(&bp, 0, sizeof(bp));
  bool aligned = true;
  for (int s = 0; s < n_frames; ++s) {
    if (!test_frames[s] || !ref_frames[s]) return fail(ctx, FVVDP_B200_ERR_INVALID, "null frame pointer %d", s);
    bp.slot[0][s] = test_frames[s];
    bp.slot[1][s] = ref_frames[s];
    aligned = aligned && (((uintptr_t)test_frames[s] | (uintptr_t)ref_frames[s]) % 16 == 0);
  }
  bp.w = fmt->width; bp.h = fmt->height;
  bp.sC = strides[0]; bp.sH = strides[1]; bp.sW = strides[2];
  bp.C = fmt->in_channels; bp.dtype = fmt->in_dtype; bp.eotf = fmt->eotf;
  bp.Yscale = fmt->Y_peak - fmt->Y_black; bp.Y_black = fmt->Y_black; bp.Y_peak = fmt->Y_peak; bp.gamma = fmt->gamma;
  bp.L_min = fmt->L_min; bp.L_max = fmt->L_max;
  for (int i = 0; i < 3; ++i) bp.rgb2y[i] = fmt->rgb2y[i];
  const bool vec_ok = strides[2] == 1 && aligned && strides[1] % 4 == 0 && (fmt->in_channels == 1 || strides[0] % 4 == 0);
  static_assert(sizeof(fvvdp_b200_pu_params) == sizeof(PuParams), "PU21 parameter block");
  cudaError_t le = fused::launch_pu_frames(bp, params, sq_err_out, n_frames, vec_ok, (cudaStream_t)cuda_stream);
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "pu_frames_kernel launch: %s", cudaGetErrorString(le));
  return FVVDP_B200_OK;
}

extern "C" int fvvdp_b200_pool_jod(const float* q, int n_bands, int64_t n_frames, int64_t q_stride, const fvvdp_b200_pool_params* params,
                                   int cuda_device, float* jod_out, void* cuda_stream) {
  fvvdp_b200_ctx* ctx = nullptr;  // ctx-free: errors are reported through fvvdp_b200_last_error(NULL)
  if (!q || !params || !jod_out) return fail(ctx, FVVDP_B200_ERR_INVALID, "null argument");
  if (n_bands < 1 || n_bands >= FVVDP_B200_MAX_LEVELS || n_frames < 1 || q_stride < n_frames) return fail(ctx, FVVDP_B200_ERR_INVALID, "bad q_per_ch shape");
  CU(cudaSetDevice(cuda_device));
  pool_kernel<<<1, 256, 0, (cudaStream_t)cuda_stream>>>(q, n_bands, (long long)n_frames, (long long)q_stride, *params, jod_out);
  cudaError_t le = cudaGetLastError();
  if (le != cudaSuccess) return fail(ctx, FVVDP_B200_ERR_CUDA, "pool_kernel launch: %s", cudaGetErrorString(le));
  return FVVDP_B200_OK;
}
