/*
 * fvvdp_b200.h -- C ABI of the B200-native FovVideoVDP per-frame core (libfvvdp_b200.so).
 *
 * The reference (gfxdisp/FovVideoVDP, pyfvvdp) has no native boundary: its hot path is Python that
 * issues PyTorch tensor ops.  The entry points below are what a binding for that path replaces:
 *
 *   fvvdp_b200_create        <- fvvdp.__init__ / set_display_model / load_config / preload_cache
 *                               (pyfvvdp/fvvdp.py:59-100,113-161,505-518) and the per-clip set-up in
 *                               predict_video_source (:201-230: pyramid layout, filter_len, temporal filters)
 *   fvvdp_b200_score_block   <- the frame loop of predict_video_source (:246-311) for a block of frames:
 *                               _get_frame / display EOTF (video_source.py:180-208, fvvdp_display_model.py:147-165),
 *                               sliding window + temporal FIR (:258-300), process_block_of_frames (:359-478):
 *                               fvvdp_contrast_pyr.decompose (fvvdp_lpyr_dec.py:248-273), cached_sensitivity
 *                               (:520-537, interp.py:11-59), apply_masking_model (:574-596), lp_norm (:598-607)
 *   fvvdp_b200_pool_jod      <- do_pooling_and_jods (:337-357), run after the optional all-reduce of q_per_ch
 *   fvvdp_b200_read_tap      <- the debug tap points of fvvdp.py:364,410-411,456 (verify_against_matlab)
 *   fvvdp_b200_destroy       <- object lifetime (Python GC in the reference)
 *
 * Conventions: plain C, POD structs, device pointers as void*; every call returns 0 on success or a
 * negative fvvdp_b200_status and never throws; all device work is enqueued on the caller's stream and is
 * asynchronous (no host synchronisation inside score_block); the caller keeps every buffer it passes alive
 * until the stream has consumed it; a ctx owns its workspace and must be used from one thread/stream at a
 * time.  The final pooling over bands/channels/frames into JOD (do_pooling_and_jods, fvvdp.py:337-357) is a
 * separate, ctx-free entry point because in the multi-GPU path it follows the all-reduce of q_per_ch.
 */
#ifndef FVVDP_B200_H_
#define FVVDP_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FVVDP_B200_ABI_VERSION 1
#define FVVDP_B200_MAX_LEVELS 16
#define FVVDP_B200_MAX_FILTER_LEN 32
#define FVVDP_B200_MAX_BLOCK_FRAMES 96
#define FVVDP_B200_MAX_SLOTS (FVVDP_B200_MAX_BLOCK_FRAMES + FVVDP_B200_MAX_FILTER_LEN)

typedef enum {
  FVVDP_B200_OK = 0,
  FVVDP_B200_ERR_INVALID = -1,  /* bad argument / unsupported configuration */
  FVVDP_B200_ERR_CUDA = -2,     /* a CUDA runtime call failed; see fvvdp_b200_last_error */
  FVVDP_B200_ERR_NOMEM = -3
} fvvdp_b200_status;

typedef enum {                /* photometric model applied to every input sample */
  FVVDP_B200_EOTF_NONE = 0,   /* input already is luminance in cd/m^2 (generic video sources) */
  FVVDP_B200_EOTF_SRGB = 1,   /* fvvdp_display_model.py:17-19,157 */
  FVVDP_B200_EOTF_GAMMA = 2,  /* :159 (also fvvdp_display_photo_gog :266-279) */
  FVVDP_B200_EOTF_PQ = 3,     /* :100-112,161 */
  FVVDP_B200_EOTF_LINEAR = 4, /* :163 */
  FVVDP_B200_EOTF_ABSOLUTE = 5 /* fvvdp_display_photo_absolute.forward :203-212 */
} fvvdp_b200_eotf;

typedef enum {
  FVVDP_B200_F32 = 0,
  FVVDP_B200_U8 = 1,          /* /255          video_source.py:198 */
  FVVDP_B200_U16 = 2          /* /65535, passed as the int16 bit pattern  video_source.py:186-196 */
} fvvdp_b200_dtype;

typedef enum {                /* fvvdp_b200_read_tap selectors */
  FVVDP_B200_TAP_R = 0,       /* temporal channels (n_ch,H,W), level ignored */
  FVVDP_B200_TAP_GAUSS = 1,   /* Gaussian level l (n_ch,h_l,w_l), l >= 1 */
  FVVDP_B200_TAP_CONTRAST = 2,/* band l contrast x band_mul (n_ch,h_l,w_l)  = T_f/R_f of fvvdp.py:395-396 */
  FVVDP_B200_TAP_LBKG = 3,    /* (h_l,w_l) */
  FVVDP_B200_TAP_S = 4,       /* sensitivity x sensitivity_correction (temp_ch,h_l,w_l) */
  FVVDP_B200_TAP_D = 5,       /* masked difference (temp_ch,h_l,w_l) */
  FVVDP_B200_TAP_DMAP_BAND = 6 /* heat-map band: (D_sust + w_transient D_trans)/band_mul (h_l,w_l) */
} fvvdp_b200_tap;

typedef struct fvvdp_b200_config {
  int32_t abi_version;        /* FVVDP_B200_ABI_VERSION */
  int32_t width, height;      /* frame size in pixels */
  int32_t n_levels;           /* Gaussian levels = lpyr.height + 1; bands scored = n_levels - 1 */
  float band_freq[FVVDP_B200_MAX_LEVELS]; /* lpyr.get_freqs(), cycles/degree */
  int32_t temp_ch;            /* 1 = image (channels [T,R]); 2 = video ([T_sust,R_sust,T_trans,R_trans]) */
  int32_t filter_len;         /* taps of the temporal filters (1 for images) */
  float filt[2][FVVDP_B200_MAX_FILTER_LEN]; /* get_temporal_filters(): F[0] sustained, F[1] transient; F[.][0] weighs the NEWEST frame */
  /* photometry */
  int32_t eotf;               /* fvvdp_b200_eotf */
  float Y_peak, Y_black, gamma, L_min, L_max;
  float rgb2y[3];
  int32_t in_dtype;           /* fvvdp_b200_dtype */
  int32_t in_channels;        /* 1 or 3 */
  /* CSF look-up tables (host pointers, copied at create): axes of 32 entries and S_log[2][32][32][32] = [omega][Y][rho][ecc] */
  const float* csf_rho_log;
  const float* csf_Y_log;
  const float* csf_ecc_sqrt;
  const float* csf_S_log;
  float csf_rho_range[2], csf_Y_range[2], csf_ecc_range[2]; /* clamp ranges (linear units) */
  /* masking / pooling constants */
  float mask_p, mask_q[2], mask_c_mul; /* 10^mask_c */
  float sens_mul;             /* 10^(sensitivity_correction/20) */
  float beta;                 /* spatial pooling exponent */
  float w_transient;          /* used only for the heat-map bands */
  /* foveation: 0 = off; 1 = stock fvvdp_display_geometry (maps computed from display_size_m / distance_m / ppd_centre);
   * 2 = custom geometry plugin: per-band maps are supplied with fvvdp_b200_set_foveation_maps and fixation_xy of
   * score_block holds the gaze DIRECTION in degrees (the plugin's pix2view_direction of the fixation point) */
  int32_t foveated;
  float display_size_m[2], distance_m, ppd_centre;
  /* options */
  int32_t want_dmap;          /* 1: keep per-band difference maps for the heat-map path; 2: also keep the sustained test
                               * frames (context image of fvvdp_b200_heatmap_visualize) */
  int32_t want_taps;          /* keep intermediate tensors readable through fvvdp_b200_read_tap (tests) */
  int32_t max_block_frames;   /* frames scored per call, 1..FVVDP_B200_MAX_BLOCK_FRAMES */
} fvvdp_b200_config;

typedef struct fvvdp_b200_ctx fvvdp_b200_ctx;

/* Allocate a scoring context (workspace for max_block_frames frames) on `cuda_device`. */
int fvvdp_b200_create(const fvvdp_b200_config* cfg, int cuda_device, fvvdp_b200_ctx** out);

/*
 * Score frames [f, f+n_frames) of a test/reference pair.
 *   test_slots/ref_slots: HOST arrays of n_frames + filter_len - 1 DEVICE pointers, oldest first; slot s
 *       is the frame that sits at time (f - (filter_len-1) + s), i.e. the caller resolves the temporal
 *       padding rule (fvvdp.py:258-285) and any frame-block halo by choosing the pointers.  Each pointer
 *       addresses sample (c=0,y=0,x=0) of one frame; strides[3] = element strides of (channel,row,column).
 *   fixation_xy: HOST array n_frames x 2 (x,y in full-resolution pixels) or NULL (ignored unless foveated).
 *   q_out: DEVICE float array laid out (n_bands, 2, q_stride); frame i of this call is written to column
 *       q_col0 + i:  Q[bb,cc] = (sum D^beta / Npix)^(1/beta)   (fvvdp.py:467,598-607).
 *   flags_out: optional DEVICE uint32; bit 0 is OR-ed in when an input sample was outside [0,1]
 *       (the reference logs "Pixel outside the valid range 0-1", fvvdp_display_model.py:149-151).
 */
int fvvdp_b200_score_block(fvvdp_b200_ctx* ctx, const void* const* test_slots, const void* const* ref_slots,
                           const int64_t strides[3], int n_frames, const float* fixation_xy, float* q_out,
                           int64_t q_stride, int64_t q_col0, uint32_t* flags_out, void* cuda_stream);

/*
 * Heat-map (fvvdp.py:469-473, heatmap="raw"): reconstruct the difference-map pyramid of frame
 * `frame_in_block` of the last score_block call and write |jod_a| * recon^beta_jod as fp16 (H,W) to
 * `dmap_out` (DEVICE).  Requires want_dmap.
 */
int fvvdp_b200_heatmap(fvvdp_b200_ctx* ctx, int frame_in_block, float beta_jod, float jod_a_abs, void* dmap_out_f16,
                       void* cuda_stream);

/*
 * Heat-map visualisation (fvvdp.py:474-476 -> visualize_diff_map, visualize_diff_map.py:58-107; heatmap="threshold" /
 * "supra-threshold"): the difference map of frame `frame_in_block`, clamped to [0,1], goes through the colour map
 * and is multiplied by the tone-mapped context image (log luminance of the sustained test frame, 1024-bin histogram
 * tone curve, vis_tonemap :26-50).  Writes fp16 sRGB planes (3,H,W) to `rgb_out_f16` (DEVICE).  Requires want_dmap = 2.
 */
typedef enum { FVVDP_B200_CMAP_THRESHOLD = 0, FVVDP_B200_CMAP_SUPRA_THRESHOLD = 1 } fvvdp_b200_colormap;
int fvvdp_b200_heatmap_visualize(fvvdp_b200_ctx* ctx, int frame_in_block, float beta_jod, float jod_a_abs, int colormap,
                                 void* rgb_out_f16, void* cuda_stream);

/*
 * Foveation maps of a custom fvvdp_display_geometry subclass (the reference calls the plugin's pix2view_direction and
 * get_resolution_magnification per band, fvvdp.py:422-438; example: pytorch_examples/ex_custom_ppd.py:38-57).
 *   view_xy:  DEVICE (2, h_l, w_l) view direction of every band pixel in degrees
 *   log2_rho: DEVICE (h_l, w_l) log2 of rho_band[l] * res_mag clamped to the CSF table's rho range
 * The caller keeps both alive while the ctx scores.  Requires cfg.foveated = 2; every scored band needs its maps.
 */
int fvvdp_b200_set_foveation_maps(fvvdp_b200_ctx* ctx, int level, const float* view_xy, const float* log2_rho);

/*
 * Video front end (ctx-free): one planar Y'CbCr frame -> luminance, replacing YUVReader.get_frame_rgb_tensor /
 * video_reader_yuv_pytorch.unpack + the display model of fvvdp_video_source_dm (video_source_yuv.py:157-228,299-302,
 * video_source_file.py:219-276): limited-range fixed2float, bilinear 4:2:0 chroma upsampling, ycbcr2rgb, clip to [0,1],
 * display EOTF, RGB2Y.  Planes are DEVICE pointers (uint8, or uint16 when bit_depth > 8); 4:2:0 needs even width/height.
 * lum_out: DEVICE float (H,W) in cd/m^2, or NULL; rgb_out: DEVICE float (H,W,3) display-encoded RGB (for callers that
 * apply their own photometry), or NULL.  With desc->resize set, both have the output size (out_height, out_width).
 */
typedef enum fvvdp_b200_resize {
  FVVDP_B200_RESIZE_NONE = 0, FVVDP_B200_RESIZE_NEAREST = 1, FVVDP_B200_RESIZE_BILINEAR = 2, FVVDP_B200_RESIZE_BICUBIC = 3,
  FVVDP_B200_RESIZE_AREA = 4
} fvvdp_b200_resize;
typedef struct fvvdp_b200_yuv_desc {
  int32_t width, height;
  int32_t bit_depth;          /* 8..16 */
  int32_t chroma_420;         /* 1: 4:2:0, 0: 4:4:4 */
  float ycbcr2rgb[9];         /* row-major 3x3 */
  int32_t eotf;               /* fvvdp_b200_eotf */
  float Y_peak, Y_black, gamma, L_min, L_max;
  float rgb2y[3];
  /* full-screen resize (fvvdp_video_source_yuv_file(full_screen_resize=..., resize_resolution=...), video_source_yuv.py:293-297):
   * with resize != NONE the R'G'B' frame is resampled to out_width x out_height like torch.nn.functional.interpolate(mode=...)
   * and clipped to [0,1] before the display EOTF; lum_out / rgb_out / the context then have the OUTPUT size */
  int32_t resize;             /* fvvdp_b200_resize */
  int32_t out_width, out_height;
} fvvdp_b200_yuv_desc;
int fvvdp_b200_yuv_to_luminance(const fvvdp_b200_yuv_desc* desc, const void* y_plane, const void* u_plane, const void* v_plane,
                                float* lum_out, float* rgb_out, int cuda_device, void* cuda_stream);

/*
 * score_block for raw planar Y'CbCr clips: test_frames / ref_frames are DEVICE copies of the n_frames + filter_len - 1 window
 * slots' frames exactly as stored in the .yuv file (Y plane, Cb plane, Cr plane; 8-bit or 16-bit little-endian samples).  One
 * launch unpacks, upsamples the chroma planes, converts Y'CbCr -> R'G'B', applies the display EOTF and RGB -> Y for both streams
 * (YUVReader.get_frame_rgb_tensor + fvvdp_video_source_yuv_file._get_frame, video_source_yuv.py:157-228,290-302) and writes the
 * (test, reference) luminance planes level 0 of the metric stages by TMA; the rest is score_block.  The ctx is created with
 * eotf = NONE, in_dtype = F32, in_channels = 1 and filter_len <= 16.
 */
int fvvdp_b200_score_block_yuv(fvvdp_b200_ctx* ctx, const fvvdp_b200_yuv_desc* desc, const void* const* test_frames,
                               const void* const* ref_frames, int n_frames, const float* fixation_xy, float* q_out, int64_t q_stride,
                               int64_t q_col0, void* cuda_stream);

/*
 * PU21-PSNR (the reference CLI's second metric, pupsnr.py:52-79 with utils.PU, utils.py:157-202), ctx-free: adds
 * sum((PU(test) - PU(ref))^2) over `n` luminance samples (DEVICE floats, cd/m^2) to *sq_err_acc (DEVICE double, zeroed by
 * the caller).  The caller finishes a frame as 20 log10(peak / sqrt(acc / n)) and averages the frames.
 */
typedef struct fvvdp_b200_pu_params {
  float p[7];                 /* PU21 parameters (utils.py:172-179) */
  float L_min, L_max;         /* luminance clip range, 0.005 .. 10000 cd/m^2 */
} fvvdp_b200_pu_params;
int fvvdp_b200_pu_sq_err(const float* lum_test, const float* lum_ref, int64_t n, const fvvdp_b200_pu_params* params,
                         double* sq_err_acc, int cuda_device, void* cuda_stream);

/*
 * PU21-PSNR for a block of frames in ONE launch, straight from the frames as the user holds them (replaces the per-frame loop
 * of pupsnr.py:52-79 incl. get_test_frame/get_reference_frame -> display_photometry.forward, video_source.py:180-208):
 * sample -> [0,1] -> display EOTF -> RGB2Y -> PU21 of both streams -> squared difference, sq_err_out[i] (DEVICE double, zeroed
 * by the caller) += the sum over frame i.  Frames are DEVICE pointers with the element strides of score_block; n_frames <=
 * FVVDP_B200_MAX_SLOTS per call.
 */
typedef struct fvvdp_b200_frame_format {
  int32_t width, height;
  int32_t in_dtype;           /* fvvdp_b200_dtype */
  int32_t in_channels;        /* 1 or 3 */
  int32_t eotf;               /* fvvdp_b200_eotf */
  float Y_peak, Y_black, gamma, L_min, L_max;
  float rgb2y[3];
} fvvdp_b200_frame_format;
int fvvdp_b200_pu_sq_err_frames(const fvvdp_b200_frame_format* fmt, const void* const* test_frames, const void* const* ref_frames,
                                const int64_t strides[3], int n_frames, const fvvdp_b200_pu_params* params, double* sq_err_out,
                                int cuda_device, void* cuda_stream);

typedef struct fvvdp_b200_pool_params {
  float beta_sch, beta_tch, beta_t; /* Lp exponents over spatial bands, temporal channels, frames */
  float w_transient;                /* weight of the transient channel */
  float jod_a, log_jod_exp;         /* JOD regression */
} fvvdp_b200_pool_params;

/*
 * do_pooling_and_jods (fvvdp.py:337-357): q (DEVICE, (n_bands, 2, q_stride), columns [0, n_frames) used) ->
 * jod_out[0] = JOD, jod_out[1] = pooled Q before the JOD regression (DEVICE, 2 floats).  Stream-ordered.
 */
int fvvdp_b200_pool_jod(const float* q, int n_bands, int64_t n_frames, int64_t q_stride, const fvvdp_b200_pool_params* params,
                        int cuda_device, float* jod_out, void* cuda_stream);

/* Copy an intermediate tensor of frame `frame_in_block` of the last score_block call to `dst` (DEVICE floats).
 * Requires want_taps (want_dmap for TAP_DMAP_BAND).  Returns the number of floats written, or <0. */
int64_t fvvdp_b200_read_tap(fvvdp_b200_ctx* ctx, int tap, int level, int frame_in_block, float* dst, int64_t dst_capacity,
                            void* cuda_stream);

/* Size (rows, columns) of pyramid level l. */
int fvvdp_b200_level_size(const fvvdp_b200_ctx* ctx, int level, int32_t* h, int32_t* w);

/* Kernel launches enqueued by this ctx so far (for bench.py's gpu_launches). */
int64_t fvvdp_b200_launch_count(const fvvdp_b200_ctx* ctx);

/*
 * Per-kernel device timing (bench.py's roofline): when enabled, every kernel launch of score_block is bracketed by
 * CUDA events on the caller's stream.  profile_read synchronises with the last recorded event, adds up the elapsed
 * milliseconds and launch counts per kernel class since the previous read, and recycles the events.
 * Classes: 0 = front (EOTF + temporal FIR), 1 + l = pyramid level l, FVVDP_B200_MAX_LEVELS + 1 = final.
 * `ms` and `count` hold FVVDP_B200_PROFILE_CLASSES entries.
 */
#define FVVDP_B200_PROFILE_CLASSES (FVVDP_B200_MAX_LEVELS + 2)
int fvvdp_b200_profile(fvvdp_b200_ctx* ctx, int enable);
int fvvdp_b200_profile_read(fvvdp_b200_ctx* ctx, float* ms, int32_t* count);

/* Algorithmic / structural byte counts of the last score_block call: [0] compulsory input bytes,
 * [1] bytes the kernels of this plan read+write to global memory. */
int fvvdp_b200_traffic_model(const fvvdp_b200_ctx* ctx, double out_bytes[2]);

int fvvdp_b200_destroy(fvvdp_b200_ctx* ctx);

/* Last error message of `ctx` (or of the last failed create when ctx is NULL).  Never NULL. */
const char* fvvdp_b200_last_error(const fvvdp_b200_ctx* ctx);

int fvvdp_b200_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* FVVDP_B200_H_ */
