/*
 * blobsplat.h — C ABI of the B200-native (sm_100a) blob-splat library.
 *
 * This is the drop-in boundary for ONE hot path of TencentARC/BlobCtrl: the BlobGAN-style blob
 * renderer in blobctrl/utils/utils.py.  The reference has no FFI of its own (it is pure Python over
 * ATen), so each entry point below cites the reference Python interface (file:line, relative to the
 * reference tree) whose arithmetic it replaces.  The host side that keeps the reference's Python
 * signatures lives in blobctrl_b200/ and binds these symbols with ctypes.
 *
 * Conventions
 *   - Plain pointers and sizes only; every pointer is a DEVICE pointer unless stated otherwise.
 *   - The caller owns and allocates every buffer; the library never allocates, frees or retains
 *     device memory, keeps no mutable global state and is re-entrant.
 *   - All work is enqueued asynchronously on `stream` (a cudaStream_t passed as void*; NULL = the
 *     legacy default stream).  No host synchronisation, no allocation: every call is CUDA-graph
 *     capturable.
 *   - `device` is the CUDA ordinal that owns the buffers, or -1 for "the calling thread's current
 *     device".  When >= 0 the library makes it current for the duration of the call and restores
 *     the previous one.
 *   - Every function returns 0 on success or a negative blobsplat_status; the message is available
 *     per thread through blobsplat_last_error().  There is NO CPU fallback: unsupported
 *     combinations fail with BLOBSPLAT_E_UNSUPPORTED.
 *   - Tensors are dense row-major in the shapes given; K = M + 1 (channel 0 = background).
 */
#ifndef BLOBSPLAT_H_
#define BLOBSPLAT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BLOBSPLAT_ABI_VERSION 4

#if defined(__GNUC__)
#define BLOBSPLAT_API __attribute__((visibility("default")))
#else
#define BLOBSPLAT_API
#endif

typedef enum {
  BLOBSPLAT_OK = 0,
  BLOBSPLAT_E_INVALID = -1,      /* bad shape / null pointer / misaligned buffer / bad enum */
  BLOBSPLAT_E_UNSUPPORTED = -2,  /* valid request this build cannot run (dtype combo, K or C limits) */
  BLOBSPLAT_E_CUDA = -3          /* a CUDA runtime call failed; message carries cudaGetErrorString */
} blobsplat_status;

typedef enum {
  BLOBSPLAT_F32 = 0,
  BLOBSPLAT_F64 = 1,
  BLOBSPLAT_BF16 = 2,
  BLOBSPLAT_F16 = 3
} blobsplat_dtype;

/* utils.py:183-191 — only_splatting_bg / only_splatting_fg */
typedef enum {
  BLOBSPLAT_SELECT_ALL = 0, /* channels 0..M   -> K planes  */
  BLOBSPLAT_SELECT_FG = 1,  /* channels 1..M   -> M planes  */
  BLOBSPLAT_SELECT_BG = 2   /* channel 0 only  -> 1 plane   */
} blobsplat_select;

/* Stage-2 thread mapping (both are measured; AUTO picks by shape) */
typedef enum {
  BLOBSPLAT_COMPOSITE_AUTO = 0,
  BLOBSPLAT_COMPOSITE_LANE_PIXEL = 1, /* one thread walks k = M..0 with the transmittance in a register */
  BLOBSPLAT_COMPOSITE_WARP_SCAN = 2   /* lane = blob; multiplicative suffix scan with warp shuffles */
} blobsplat_composite_mode;

/* Stage-3 engine */
typedef enum {
  BLOBSPLAT_ENGINE_AUTO = 0,
  BLOBSPLAT_ENGINE_FMA = 1,    /* CUDA-core FP32/FP64 FMA tiles */
  BLOBSPLAT_ENGINE_TENSOR = 2, /* tcgen05 MMA, weights staged into tensor memory by threads (bf16/f16: kind::f16; f32: split precision) */
  BLOBSPLAT_ENGINE_TMA = 3     /* 16-bit maps only (16-byte aligned, pixel-contiguous, H*W and C multiples of 8, K <= 256): operands by
                                  TMA loads, tcgen05 MMA from shared memory, 128-byte line stores; whole pyramids per launch */
} blobsplat_engine;

typedef struct {
  int abi_version;
  int sm_arch;            /* 100 */
  int max_blobs;          /* largest M accepted by blobsplat_scores */
  int tensor_max_k;       /* largest K = M+1 the tensor-core render accepts */
  int tensor_c_multiple;  /* C must be a multiple of this for the tensor-core render (1: any C) */
  int tensor_max_c;       /* largest C the tensor-core render accepts */
} blobsplat_caps;

BLOBSPLAT_API int blobsplat_abi_version(void);
BLOBSPLAT_API int blobsplat_get_caps(blobsplat_caps* out);
/* copies the calling thread's last error message (NUL-terminated) into buf; returns its length */
BLOBSPLAT_API int blobsplat_last_error(char* buf, size_t cap);

/*
 * (1) scores — stages 1+2.  Replaces utils.py:120-194 (splat_features up to `return_d_score`):
 *     delta = (pixel - centre*size)/size, q = delta^T Sigma^-1 delta, s = min(1, 2*sigmoid(-q)),
 *     sizes < 0.5 -> 1e-6, background alpha 1 prepended, d_k = s_k * prod_{j>k}(1 - s_j), d_M = s_M.
 *
 *   xs, ys   [N, M]        blob centres in [0,1] (fractions of W, H)
 *   covs     [N, M, 2, 2]  covariances, normalised by the image diagonal^2
 *   sizes    [N, M]        float32 existence flags (the reference compares `< 0.5`)
 *   param_dtype            BLOBSPLAT_F32 or BLOBSPLAT_F64 (dtype of xs/ys/covs; also the compute type)
 *   composed [N, Ksel, H, W] (Ksel = K, M or 1 by `select`), dtype composed_dtype; may be NULL
 *   raw      [N, K, H, W]  raw scores incl. the background plane of ones, dtype raw_dtype; may be NULL
 *            (the reference's layout dict holds them as [N,H,W,K]; the host wrapper returns a
 *             permuted view of this planar buffer — same shape and values)
 *   Output dtypes: F32/BF16/F16 when param_dtype is F32; F64 when param_dtype is F64.
 */
BLOBSPLAT_API int blobsplat_scores(const void* xs, const void* ys, const void* covs, const float* sizes,
                     int param_dtype, int N, int M, int H, int W, int select,
                     void* composed, int composed_dtype, void* raw, int raw_dtype,
                     int composite_mode, int device, void* stream);

/*
 * (1a) scores from ellipses — the device-side front end.  Same maps as (1), but the blobs are
 *      given as OpenCV-convention ellipses instead of (centre, covariance): replaces the host recipe
 *      get_gs_from_ellipse -> normalize_gs -> get_blob_dict_from_norm_gs (scripts/blobctrl_inference.py:71-109, same
 *      code at scripts/blobctrl_app.py:604-629; ellipse_to_gaussian at blobctrl/utils/utils.py:297-341) plus (1).
 *   ellipses [N, M, 5] float32: xc, yc, d1, d2 in pixels of an img_w x img_h image; angle in degrees
 *            (cv2.fitEllipse convention).  sizes [N, M] as in (1).  Blob index = depth order (highest in front).
 */
BLOBSPLAT_API int blobsplat_scores_ellipse(const float* ellipses, const float* sizes, float img_w, float img_h,
                             int N, int M, int H, int W, int select,
                             void* composed, int composed_dtype, void* raw, int raw_dtype,
                             int device, void* stream);

/*
 * (1c) preview — stages 1+2 and the C = 3 colour splat in ONE launch, the score maps never materialised.  Replaces
 *      splat_features(..., is_viz=True, only_vis=True, viz_score_fn=identity) — utils.py:198-223 -> visualize_features
 *      (utils.py:244-270) -> splat_features_from_scores (utils.py:57-77) — i.e. the UI preview of
 *      scripts/blobctrl_app.py:637-650:   image[n, ch, y, x] = sum_k d_k[n, y, x] * colors[k, ch].
 *   param_dtype F32 or F64 = dtype of xs / ys / covs / colors / image / composed (the app renders in float64).
 *   colors [K, 3] (colors_per_image = 0) or [N, K, 3] (1); image [N, 3, H, W]; composed [N, K, H, W] or NULL.
 */
BLOBSPLAT_API int blobsplat_preview(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype,
                      const void* colors, int colors_per_image, int N, int M, int H, int W,
                      void* image, void* composed, int device, void* stream);

/*
 * (1d) preview as the 8-bit picture the UI shows.  Replaces the whole of get_blob_vis_img_from_blob_dict up to
 *      Image.fromarray (scripts/blobctrl_app.py:637-648): the preview above, then
 *      blob_vis[0].permute(1, 2, 0).contiguous().cpu().numpy(); (img * 255).astype(np.uint8)
 *      — the permute and the conversion happen in the render launch, and the device-to-host copy moves 3 bytes per
 *      pixel instead of 12 (float32) or 24 (float64).
 *   image_hwc [N, H, W, 3] uint8 = truncation of value * 255 computed in param_dtype (numpy's astype).
 */
BLOBSPLAT_API int blobsplat_preview_u8(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype,
                         const void* colors, int colors_per_image, int N, int M, int H, int W,
                         unsigned char* image_hwc, int device, void* stream);

/*
 * (1b) composite only.  Replaces utils.py:179-181 / :205-206 applied to caller-modified raw scores
 *      (the `viz_score_fn` branch).  scores_in / composed: planar [N, K, H, W], same dtype.
 */
BLOBSPLAT_API int blobsplat_composite(const void* scores_in, void* composed, int N, int K, int H, int W,
                        int dtype, int device, void* stream);

/*
 * (2) bilinear resize, align_corners = False.  Replaces torch.nn.functional.interpolate(mode=
 *     'bilinear') as used by pyramid_resize (utils.py:280-294) and by splat_features_from_scores
 *     (utils.py:70-73).  in [B, Hin, Win] -> out [B, Hout, Wout], same dtype.
 */
BLOBSPLAT_API int blobsplat_resize_bilinear(const void* in, void* out, int B, int Hin, int Win, int Hout, int Wout,
                              int dtype, int device, void* stream);

/*
 * (2b) whole pyramid in one launch: in [B, S, S] -> levels S/2, S/4, ... (n_levels of them), each an
 *      exact 2x2 mean of the previous (== the reference's bilinear halving for even sizes, SURVEY
 *      probe B7).  outs: HOST array of n_levels device pointers.  S must be divisible by 2^n_levels.
 */
BLOBSPLAT_API int blobsplat_pyramid(const void* in, void* const* outs, int n_levels, int B, int S, int dtype,
                      int device, void* stream);

/*
 * (3) feature splat — stage 3.  Replaces splat_features_from_scores (utils.py:57-77) and the
 *     duplicate pipeline method (pipelines/pipeline_blobnet.py:706-721) after the optional resize:
 *         out[n, c, y, x] = sum_k scores[n, k, y, x] * features[n, k, c]      (NCHW, contiguous)
 *   scores: element strides (in elements) stride_n, stride_k, stride_p with pixel p = y*W + x
 *           ([N,K,H,W] contiguous: K*P, P, 1;  [N,H,W,K] contiguous: P*K, 1, K)
 *   features [N, K, C] contiguous, same dtype as scores and out.
 *   engine: AUTO runs the contraction on tcgen05 tensor cores when it is a real dense one (C >= 64 and K >= 12, or
 *           for 16-bit maps any K once the output has >= 2^24 elements;
 *           K <= 128, not float64; float32 uses the 3xTF32 split) and on CUDA-core FMA tiles
 *           otherwise; FMA / TENSOR force one (TENSOR fails with BLOBSPLAT_E_UNSUPPORTED outside its envelope).
 */
BLOBSPLAT_API int blobsplat_feature_splat(const void* scores, int64_t stride_n, int64_t stride_k, int64_t stride_p,
                            const void* features, void* out, int N, int K, int C, int H, int W,
                            int dtype, int engine, int device, void* stream);

/*
 * (3b) feature splat of a whole pyramid — the multi-resolution BlobNet conditioning (BASELINE configs[2]): the same
 *      contraction as (3) for n_levels score maps of one batch (same N, K, dtype; per-level H, W, C), i.e. the loop
 *      `for s in sizes: splat_features_from_scores(scores_pyramid[s], features[s], s)` over the pyramid returned by
 *      splat_features (utils.py:235-241, 57-77).  All arrays are HOST arrays of n_levels entries (pointers are
 *      device pointers).  engine FMA: level by level exactly as (3).  engine TMA, and AUTO for 16-bit pyramids
 *      that are real contractions (2..4 levels, K >= 12, C[i] >= 64, maps inside the TMA envelope): ONE launch of the
 *      TMA engine over all levels.  engine TENSOR: when every level shares an operand tiling (2..4 levels,
 *      C[i] >= 64 with one common channel tile) ONE launch of the thread-staged tensor engine over the concatenated
 *      tile sequence, otherwise level by level.  AUTO otherwise: level by level as (3).
 */
BLOBSPLAT_API int blobsplat_feature_splat_levels(int n_levels, const void* const* scores, const int64_t* stride_n,
                              const int64_t* stride_k, const int64_t* stride_p, const void* const* features,
                              void* const* outs, int N, int K, const int* C, const int* H, const int* W, int dtype,
                              int engine, int device, void* stream);

/*
 * (3b) conditioning fill — fused construct_blobnet_input for the loop-invariant channels.
 *      Replaces the stage-3 splat at pipelines/pipeline_blobnet.py:984 plus the per-step torch.cat of
 *      construct_blobnet_input (:724-739, called at :1043-1049 and :1071-1076) for everything except the 4 latent
 *      channels: writes, into a persistent buffer out [B, c_total, h, halves*w] (halves = 2: left | right),
 *          planes c_off .. c_off+K-1      = scores[b, k]                     (when write_scores != 0)
 *          planes ..   .. +C-1            = sum_k scores[b, k] * features[b, k, c]
 *      identically in every width half.  scores [B, K, h, w], features [B, K, C] (NULL when C == 0), same dtype
 *      (F32/BF16/F16) as out; w % 4 == 0; out 16-byte aligned.
 */
BLOBSPLAT_API int blobsplat_conditioning_fill(const void* scores, const void* features, void* out, int B, int K, int C,
                                int h, int w, int c_total, int c_off, int halves, int write_scores, int dtype,
                                int device, void* stream);

/*
 * (3c) residual injection — fused scale + right-half slice + add.  Replaces, per BlobNet
 *      residual and denoising step, `residual * conditioning_scale` (models/blobnet.py:936-938), the slice
 *      `residual[..., -h:]` (pipelines/pipeline_blobnet.py:1085-1087) and `sample[..., -h:] += residual`
 *      (diffusers/src/diffusers/models/unets/unet_2d_condition.py:1215-1219; unet_2d_blocks.py:1303-1319, 2598-2615):
 *          hidden[b, c, y, Wh - cols + x] += scale_b * residual[b, c, y, Wr - cols + x],   x < cols   (in place)
 *      hidden [B, C, H, Wh], residual [B, C, H, Wr] contiguous, same dtype; scale_per_sample: device float[B] or NULL
 *      (then the scalar `scale` applies).  Rounded op by op like the reference (bit-identical in fp16/bf16/fp32).
 */
BLOBSPLAT_API int blobsplat_residual_inject(void* hidden, const void* residual, const float* scale_per_sample, float scale,
                              int B, int C, int H, int Wh, int Wr, int cols, int dtype, int device, void* stream);

/*
 * (3d) hoisted BlobNet conv_in.  BlobNet's first layer, Conv2d(lc + 1 + C, O, 3, padding=1)
 *      (models/blobnet.py:241-245, applied at :840) over the canvas of construct_blobnet_input
 *      (pipelines/pipeline_blobnet.py:724-739, rebuilt at :1043-1049 every step), split by linearity into the part that
 *      changes per step (the lc latent planes) and the loop-invariant conditioning planes, whose C rank-K feature planes
 *      (:984) collapse into per-sample 3x3 kernels.
 *   blobsplat_conv_in_weights — once per edit:   weff [B, O, 1 + K, 12] float32 (9 taps padded to 12),
 *        weff[b, o, 0]     = weight[o, lc]                                   (the score plane's own kernel)
 *        weff[b, o, 1 + k] = sum_c weight[o, lc + 1 + c] * features[b, k, c]
 *      weight [O, Cin = lc + 1 + C, 3, 3], features [B, K, C], same dtype (F32/BF16/F16).
 *   blobsplat_conv_in_hoisted — every step:
 *        out[b, o] = bias[o] + sum_{c < lc} weight[o, c] (*) latents[b, c] + sum_{j < J} weff[b, o, j] (*) cond2[b, j]
 *      latents [B, lc, h, halves*w] (left = reference-image latents, right = noisy latents), cond [B, J = 1 + K, h, w]
 *      (plane 0 = the canvas's score plane, planes 1.. = the K blob score planes; cond2 = cond repeated in every width
 *      half, as the canvas holds it), out [B, O, h, halves*w].  fp32 accumulation, one rounding to the output dtype.
 *      Equals conv_in(construct_blobnet_input(...)) up to summation order.
 */
BLOBSPLAT_API int blobsplat_conv_in_weights(const void* weight, const void* features, float* weff, int B, int O, int Cin,
                              int lc, int C, int K, int dtype, int device, void* stream);
BLOBSPLAT_API int blobsplat_conv_in_hoisted(const void* latents, const void* cond, const void* weight, const void* bias,
                              const float* weff, void* out, int B, int O, int Cin, int lc, int J, int h, int w,
                              int halves, int dtype, int device, void* stream);

/*
 * (4) fused render — stages 1+2+3 in ONE launch: blob parameters + features -> composed score maps
 *     and the feature grid at the same resolution, with the per-pixel weights never leaving the SM
 *     (tcgen05 MMA with the weights as the TMEM A operand).  Replaces the whole of splat_features
 *     (utils.py:80-241) for interp_size == score_size.
 *   features [N, K, C] (dtype feat_dtype: F32, BF16 or F16); grid [N, C, H, W] (dtype out_dtype);
 *   composed [N, K, H, W] (dtype out_dtype) or NULL.  param dtype is F32.
 *   F32 out: 3xTF32 split-precision MMA (error ~2^-21, inside the 1e-5 parity bar);
 *   BF16/F16 out: single kind::f16 MMA with fp32 accumulation.
 */
BLOBSPLAT_API int blobsplat_render(const float* xs, const float* ys, const float* covs, const float* sizes,
                     const void* features, int feat_dtype, int N, int M, int H, int W, int C,
                     void* composed, void* grid, int out_dtype, int device, void* stream);

/*
 * (4a) the same render as a one-launch LATENCY kernel on CUDA cores, float32 only: for the single small renders the
 *      reference's scripts and UI issue (scripts/blobctrl_inference.py:112-117, scripts/blobctrl_app.py:637-650; BASELINE
 *      config 2).  No tensor memory, no operand staging, full-fp32 products and sums like the reference's einsum
 *      (utils.py:77).  K = M + 1 <= 33 (BLOBSPLAT_E_UNSUPPORTED otherwise); composed may be NULL.
 */
BLOBSPLAT_API int blobsplat_render_small(const float* xs, const float* ys, const float* covs, const float* sizes,
                           const float* features, int N, int M, int H, int W, int C,
                           float* composed, float* grid, int device, void* stream);

/*
 * (4b) the multi-resolution conditioning in ONE call (BASELINE configs[2]): level l has size S >> l.  Level 0 is (4);
 *      levels 1.. are pyramid_resize of its composed maps (utils.py:280-294, exact 2x2 means) followed by
 *      splat_features_from_scores per level (utils.py:57-77).  HOST arrays of n_levels (1..4) entries:
 *      features[l] [N, M+1, C[l]] (NULL = level wanted as maps only), composed[l] [N, M+1, S>>l, S>>l] (required: the
 *      pyramid is an output, utils.py:235-241), grids[l] [N, C[l], S>>l, S>>l].  float32 parameters; float32 /
 *      bfloat16 / float16 features and maps of one dtype.  Same results, bit for bit, as calling (4), (2b), (3b) in sequence.
 *      Fewer launches: for S == 64 in 16 bits the pyramid leaves the render launch itself (render_tc2.cuh, kPyr) and the
 *      lower levels are one launch of the TMA engine — two launches for BlobNet's four resolutions.
 */
BLOBSPLAT_API int blobsplat_render_multiscale(const float* xs, const float* ys, const float* covs, const float* sizes,
                              int N, int M, int S, int n_levels, const void* const* features, const int* C,
                              void* const* composed, void* const* grids, int dtype, int device, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BLOBSPLAT_H_ */
