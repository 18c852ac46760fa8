// extern "C" surface of libblobsplat.so — argument validation, device guard, dispatch.
// Declarations and the reference interfaces each entry replaces: include/blobsplat.h.
#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace blobsplat {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int cuda_fail(cudaError_t e, const char* what) {
  set_error("CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
  return BLOBSPLAT_E_CUDA;
}

// kernels' host-side dispatchers (scores.cu, resize.cu, feature_splat.cu, render_tc.cu)
int scores_dispatch(const void*, const void*, const void*, const float*, int, int, int, int, int, int, void*, int,
                    void*, int, int, cudaStream_t);
int composite_dispatch(const void*, void*, int, int, int, int, int, cudaStream_t);
int scores_ellipse_dispatch(const float*, const float*, float, float, int, int, int, int, int, void*, int, void*, int,
                            cudaStream_t);
int resize_dispatch(const void*, void*, int, int, int, int, int, int, cudaStream_t);
int pyramid_dispatch(const void*, void* const*, int, int, int, int, cudaStream_t);
int feature_splat_fma_dispatch(const void*, int64_t, int64_t, int64_t, const void*, void*, int, int, int, int, int,
                               int, cudaStream_t);
int feature_splat_tc_dispatch(const void*, int64_t, int64_t, int64_t, const void*, void*, int, int, int, int, int, int,
                              cudaStream_t);
bool splat_tma_usable(int, const void* const*, const int64_t*, const int64_t*, const int64_t*, const void* const*, void* const*,
                      int, int, const int*, const int*, const int*, int);
int splat_tma_dispatch(int, const void* const*, const int64_t*, const int64_t*, const void* const*, void* const*, int, int,
                       const int*, const int*, const int*, int, cudaStream_t);
int render_tc_pyramid_dispatch(const float*, const float*, const float*, const float*, const void*, int, int, int, int, void*,
                               void*, void* const*, int, int, cudaStream_t);
int feature_splat_levels_tc_dispatch(int, const void* const*, const int64_t*, const int64_t*, const int64_t*,
                                     const void* const*, void* const*, int, int, const int*, const int*, const int*, int,
                                     cudaStream_t);
int render_tc_dispatch(const float*, const float*, const float*, const float*, const void*, int, int, int, int, int,
                       int, void*, void*, int, cudaStream_t);
int conditioning_fill_dispatch(const void*, const void*, void*, int, int, int, int, int, int, int, int, int, int,
                               cudaStream_t);
int residual_inject_dispatch(void*, const void*, const float*, float, int, int, int, int, int, int, int, cudaStream_t);
int preview_dispatch(const void*, const void*, const void*, const float*, int, const void*, int, int, int, int, int, void*, void*,
                     unsigned char*, cudaStream_t);
bool render_small_supported(int K);
int render_small_dispatch(const float*, const float*, const float*, const float*, const float*, int, int, int, int, int, float*, float*,
                          cudaStream_t);
int conv_in_weights_dispatch(const void*, const void*, float*, int, int, int, int, int, int, int, cudaStream_t);
int conv_in_hoisted_dispatch(const void*, const void*, const void*, const void*, const float*, void*, int, int, int, int, int, int,
                             int, int, int, cudaStream_t);
int render_tc_supported(int K, int C, int H, int W, int feat_dtype, int out_dtype, const char** why);
void render_tc_limits(int* max_k, int* c_multiple, int* max_c);

constexpr int kMaxBlobs = 1 << 20;
// BLOBSPLAT_F32_EXACT=1 (read per call): float32 stage 3 stays on the CUDA-core FMA engine under AUTO — full fp32 products
// and sums like the reference's einsum — and the fused tensor-core render reports "unsupported", so splat_features takes
// its scores + FMA-splat path.  Default off: the split-precision tensor paths are within 1e-6 of scale (bar: 1e-5).
static inline bool f32_exact() {
  const char* e = std::getenv("BLOBSPLAT_F32_EXACT");
  return e && e[0] == '1';
}

#ifndef BS_TMA_LEVELS_AUTO
#define BS_TMA_LEVELS_AUTO 1     // AUTO runs 16-bit pyramids on the TMA engine (splat_tma.cu)
#endif
#ifndef BS_FUSE_LEVELS_AUTO
#define BS_FUSE_LEVELS_AUTO 0
#endif

static bool valid_dtype(int dt) { return dt >= BLOBSPLAT_F32 && dt <= BLOBSPLAT_F16; }

}  // namespace blobsplat

using namespace blobsplat;

extern "C" {

int blobsplat_abi_version(void) { return BLOBSPLAT_ABI_VERSION; }

int blobsplat_get_caps(blobsplat_caps* out) {
  BS_CHECK_ARG(out != nullptr, "caps pointer is NULL");
  out->abi_version = BLOBSPLAT_ABI_VERSION;
  out->sm_arch = 100;
  out->max_blobs = kMaxBlobs;
  render_tc_limits(&out->tensor_max_k, &out->tensor_c_multiple, &out->tensor_max_c);
  return BLOBSPLAT_OK;
}

int blobsplat_last_error(char* buf, size_t cap) {
  const size_t len = strlen(g_err);
  if (buf && cap) {
    const size_t n = len < cap - 1 ? len : cap - 1;
    memcpy(buf, g_err, n);
    buf[n] = 0;
  }
  return (int)len;
}

int blobsplat_scores(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype, int N,
                     int M, int H, int W, int select, void* composed, int composed_dtype, void* raw, int raw_dtype,
                     int composite_mode, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1, "bad shape N=%d M=%d H=%d W=%d", N, M, H, W);
  BS_CHECK_ARG(M <= kMaxBlobs, "M=%d exceeds max_blobs=%d", M, kMaxBlobs);
  BS_CHECK_ARG((long long)H * W < (1ll << 31), "H*W must fit in int32");
  BS_CHECK_ARG(N <= 65535, "N=%d exceeds the grid limit 65535; split the batch", N);
  BS_CHECK_ARG(select >= BLOBSPLAT_SELECT_ALL && select <= BLOBSPLAT_SELECT_BG, "bad select %d", select);
  BS_CHECK_ARG(composite_mode >= BLOBSPLAT_COMPOSITE_AUTO && composite_mode <= BLOBSPLAT_COMPOSITE_WARP_SCAN,
               "bad composite_mode %d", composite_mode);
  if (N == 0) return BLOBSPLAT_OK;   // empty batch: nothing to render (buffers of an empty tensor are NULL)
  BS_CHECK_ARG(composed || raw, "both outputs are NULL");
  BS_CHECK_ARG(!composed || valid_dtype(composed_dtype), "bad composed dtype %d", composed_dtype);
  BS_CHECK_ARG(!raw || valid_dtype(raw_dtype), "bad raw dtype %d", raw_dtype);
  BS_CHECK_ARG(M == 0 || (xs && ys && covs && sizes), "NULL blob parameter pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return scores_dispatch(xs, ys, covs, sizes, param_dtype, N, M, H, W, select, composed, composed_dtype, raw,
                         raw_dtype, composite_mode, (cudaStream_t)stream);
}

int blobsplat_scores_ellipse(const float* ellipses, const float* sizes, float img_w, float img_h, int N, int M, int H,
                             int W, int select, void* composed, int composed_dtype, void* raw, int raw_dtype, int device,
                             void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1, "bad shape N=%d M=%d H=%d W=%d", N, M, H, W);
  BS_CHECK_ARG(M <= kMaxBlobs && N <= 65535 && (long long)H * W < (1ll << 31), "shape too large");
  BS_CHECK_ARG(img_w > 0.f && img_h > 0.f, "bad image size %g x %g", (double)img_w, (double)img_h);
  BS_CHECK_ARG(select >= BLOBSPLAT_SELECT_ALL && select <= BLOBSPLAT_SELECT_BG, "bad select %d", select);
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(composed || raw, "both outputs are NULL");
  BS_CHECK_ARG(!composed || valid_dtype(composed_dtype), "bad composed dtype %d", composed_dtype);
  BS_CHECK_ARG(!raw || valid_dtype(raw_dtype), "bad raw dtype %d", raw_dtype);
  BS_CHECK_ARG(M == 0 || (ellipses && sizes), "NULL ellipse parameter pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return scores_ellipse_dispatch(ellipses, sizes, img_w, img_h, N, M, H, W, select, composed, composed_dtype, raw, raw_dtype,
                                 (cudaStream_t)stream);
}

int blobsplat_preview(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype, const void* colors,
                      int colors_per_image, int N, int M, int H, int W, void* image, void* composed, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1, "bad shape N=%d M=%d H=%d W=%d", N, M, H, W);
  BS_CHECK_ARG(M <= kMaxBlobs && N <= 65535 && (long long)H * W < (1ll << 31), "shape too large");
  BS_CHECK_ARG(param_dtype == BLOBSPLAT_F32 || param_dtype == BLOBSPLAT_F64, "preview renders float32 or float64 (got %d)", param_dtype);
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(image && colors, "NULL image / colour pointer");
  BS_CHECK_ARG(M == 0 || (xs && ys && covs && sizes), "NULL blob parameter pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return preview_dispatch(xs, ys, covs, sizes, param_dtype, colors, colors_per_image, N, M, H, W, image, composed, nullptr,
                          (cudaStream_t)stream);
}

int blobsplat_preview_u8(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype, const void* colors,
                         int colors_per_image, int N, int M, int H, int W, unsigned char* image_hwc, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1, "bad shape N=%d M=%d H=%d W=%d", N, M, H, W);
  BS_CHECK_ARG(M <= kMaxBlobs && N <= 65535 && (long long)H * W < (1ll << 31), "shape too large");
  BS_CHECK_ARG(param_dtype == BLOBSPLAT_F32 || param_dtype == BLOBSPLAT_F64, "preview renders float32 or float64 (got %d)", param_dtype);
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(image_hwc && colors, "NULL image / colour pointer");
  BS_CHECK_ARG(M == 0 || (xs && ys && covs && sizes), "NULL blob parameter pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return preview_dispatch(xs, ys, covs, sizes, param_dtype, colors, colors_per_image, N, M, H, W, nullptr, nullptr, image_hwc,
                          (cudaStream_t)stream);
}

int blobsplat_composite(const void* scores_in, void* composed, int N, int K, int H, int W, int dtype, int device,
                        void* stream) {
  BS_CHECK_ARG(N >= 0 && K >= 1 && H >= 1 && W >= 1, "bad shape N=%d K=%d H=%d W=%d", N, K, H, W);
  BS_CHECK_ARG((long long)H * W < (1ll << 31) && N <= 65535, "shape too large");
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(scores_in && composed, "NULL pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return composite_dispatch(scores_in, composed, N, K, H, W, dtype, (cudaStream_t)stream);
}

int blobsplat_resize_bilinear(const void* in, void* out, int B, int Hin, int Win, int Hout, int Wout, int dtype,
                              int device, void* stream) {
  BS_CHECK_ARG(B >= 0 && Hin >= 1 && Win >= 1 && Hout >= 1 && Wout >= 1, "bad shape");
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  if (B == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(in && out, "NULL pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return resize_dispatch(in, out, B, Hin, Win, Hout, Wout, dtype, (cudaStream_t)stream);
}

int blobsplat_pyramid(const void* in, void* const* outs, int n_levels, int B, int S, int dtype, int device,
                      void* stream) {
  BS_CHECK_ARG(B >= 0 && S >= 1 && n_levels >= 1, "bad shape");
  BS_CHECK_ARG(n_levels < 31 && (S % (1 << n_levels)) == 0, "S=%d is not divisible by 2^%d", S, n_levels);
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  if (B == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(in && outs, "NULL pointer");
  for (int l = 0; l < n_levels; ++l) BS_CHECK_ARG(outs[l] != nullptr, "level %d output is NULL", l);
  DeviceGuard g(device);
  if (g.status) return g.status;
  return pyramid_dispatch(in, outs, n_levels, B, S, dtype, (cudaStream_t)stream);
}

int blobsplat_feature_splat(const void* scores, int64_t stride_n, int64_t stride_k, int64_t stride_p,
                            const void* features, void* out, int N, int K, int C, int H, int W, int dtype,
                            int engine, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && K >= 1 && C >= 1 && H >= 1 && W >= 1, "bad shape N=%d K=%d C=%d H=%d W=%d", N, K, C, H, W);
  BS_CHECK_ARG((long long)H * W < (1ll << 31) && N <= 65535, "shape too large");
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  BS_CHECK_ARG(engine >= BLOBSPLAT_ENGINE_AUTO && engine <= BLOBSPLAT_ENGINE_TMA, "bad engine %d", engine);
  BS_CHECK_ARG(stride_k >= 0 && stride_p >= 1 && stride_n >= 0, "bad strides");
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(scores && features && out, "NULL pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  // Stage 3 is a dense contraction [P x K] x [K x C]: on tensor cores (tcgen05, weights staged into TMEM) when
  // K and C make it one, on CUDA-core FMA tiles otherwise (tiny K: the K = 1 pipeline splat, C = 3 previews).
  if (engine == BLOBSPLAT_ENGINE_TMA) {
    if (!splat_tma_usable(1, &scores, &stride_n, &stride_k, &stride_p, &features, &out, N, K, &C, &H, &W, dtype))
      BS_UNSUPPORTED("TMA feature splat: needs 16-bit, 16-byte aligned, pixel-contiguous maps with H*W and C multiples of 8");
    return splat_tma_dispatch(1, &scores, &stride_n, &stride_k, &features, &out, N, K, &C, &H, &W, dtype, (cudaStream_t)stream);
  }
  // AUTO, 16-bit maps, a large output (>= 2^28 elements): the TMA engine's store stream wins (cfg5c's stage 3 alone,
  // 1024 images: 0.65 ms vs 0.77 ms); smaller launches stay on the thread-staged tensor engine (64 images: 37 vs 43 us)
  if (engine == BLOBSPLAT_ENGINE_AUTO && BS_TMA_LEVELS_AUTO && C >= 64 && K >= 12 && (long long)N * C * H * W >= (1ll << 28) &&
      splat_tma_usable(1, &scores, &stride_n, &stride_k, &stride_p, &features, &out, N, K, &C, &H, &W, dtype))
    return splat_tma_dispatch(1, &scores, &stride_n, &stride_k, &features, &out, N, K, &C, &H, &W, dtype, (cudaStream_t)stream);
  const char* why = nullptr;
  const bool tc_ok = dtype != BLOBSPLAT_F64 && render_tc_supported(K, C, H, W, dtype, dtype, &why);
  if (engine == BLOBSPLAT_ENGINE_TENSOR && !tc_ok) BS_UNSUPPORTED("tensor-core feature splat: %s", why ? why : "float64");
  // AUTO: tensor engine for a real contraction (K >= 12) and, for 16-bit maps, whatever K once the output is large
  // enough that its streaming drain beats the FMA tiles (>= 16 M elements: 2.7x at the pipeline's K = 1, C = 1024
  // splat of 16 images; 16-bit products are exact in the fp32 accumulator, so K = 1 stays bit-identical).  float32
  // keeps small K on the FMA engine: a single product is exact there, 3xTF32 is not.
  const bool big = dtype != BLOBSPLAT_F32 && (long long)N * C * H * W >= (1ll << 24);
  const bool exact = dtype == BLOBSPLAT_F32 && f32_exact();
  if (tc_ok && (engine == BLOBSPLAT_ENGINE_TENSOR || (engine == BLOBSPLAT_ENGINE_AUTO && !exact && C >= 64 && (K >= 12 || big))))
    return feature_splat_tc_dispatch(scores, stride_n, stride_k, stride_p, features, out, N, K, C, H, W, dtype,
                                     (cudaStream_t)stream);
  return feature_splat_fma_dispatch(scores, stride_n, stride_k, stride_p, features, out, N, K, C, H, W, dtype,
                                    (cudaStream_t)stream);
}

int blobsplat_feature_splat_levels(int n_levels, const void* const* scores, const int64_t* stride_n,
                                   const int64_t* stride_k, const int64_t* stride_p, const void* const* features,
                                   void* const* outs, int N, int K, const int* C, const int* H, const int* W, int dtype,
                                   int engine, int device, void* stream) {
  BS_CHECK_ARG(n_levels >= 0 && n_levels <= 16, "bad level count %d", n_levels);
  if (n_levels == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(scores && stride_n && stride_k && stride_p && features && outs && C && H && W, "NULL level array");
  BS_CHECK_ARG(engine >= BLOBSPLAT_ENGINE_AUTO && engine <= BLOBSPLAT_ENGINE_TMA, "bad engine %d", engine);
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  // One launch for the whole pyramid when every level is a dense contraction with the same operand tiling
  // (BlobNet's 640/1280-channel levels); otherwise level by level.  BS_FUSE_LEVELS_AUTO: whether AUTO fuses too —
  // measured on cfg3 (profiles/render_tc_r1.md) the per-level launches with their compile-time plane strides win
  // until operand staging overlaps across units, so AUTO stays level by level and TENSOR asks for the single launch.
  // 16-bit pyramids: the TMA engine takes all levels in one launch (AUTO for a real contraction, or on request)
  if (N > 0 && n_levels <= 4 && (engine == BLOBSPLAT_ENGINE_TMA || (engine == BLOBSPLAT_ENGINE_AUTO && BS_TMA_LEVELS_AUTO && n_levels >= 2 && K >= 12))) {
    bool ok = true;
    for (int i = 0; i < n_levels && ok; ++i) ok = scores[i] && features[i] && outs[i] && (C[i] >= 64 || engine == BLOBSPLAT_ENGINE_TMA) && H[i] >= 1 && W[i] >= 1;
    ok = ok && splat_tma_usable(n_levels, scores, stride_n, stride_k, stride_p, features, outs, N, K, C, H, W, dtype);
    if (ok) {
      DeviceGuard g(device);
      if (g.status) return g.status;
      return splat_tma_dispatch(n_levels, scores, stride_n, stride_k, features, outs, N, K, C, H, W, dtype, (cudaStream_t)stream);
    }
    if (engine == BLOBSPLAT_ENGINE_TMA)
      BS_UNSUPPORTED("TMA feature splat: needs 16-bit, 16-byte aligned, pixel-contiguous maps with H*W and C multiples of 8");
  }
  bool fuse = N > 0 && n_levels >= 2 && n_levels <= 4 && dtype != BLOBSPLAT_F64 &&
              (engine == BLOBSPLAT_ENGINE_TENSOR || (BS_FUSE_LEVELS_AUTO && engine == BLOBSPLAT_ENGINE_AUTO && K >= 12));
  for (int i = 0; i < n_levels && fuse; ++i) {
    const char* why = nullptr;
    fuse = scores[i] && features[i] && outs[i] && C[i] >= 64 && H[i] >= 1 && W[i] >= 1 &&
           (long long)H[i] * W[i] < (1ll << 31) && stride_k[i] >= 0 && stride_p[i] >= 1 && stride_n[i] >= 0 &&
           render_tc_supported(K, C[i], H[i], W[i], dtype, dtype, &why);
  }
  if (fuse && N <= 65535 && K >= 1) {
    DeviceGuard g(device);
    if (g.status) return g.status;
    const int rc = feature_splat_levels_tc_dispatch(n_levels, scores, stride_n, stride_k, stride_p, features, outs, N, K,
                                                    C, H, W, dtype, (cudaStream_t)stream);
    if (rc <= 0) return rc;     // 1 = the levels cannot share a launch
  }
  for (int i = 0; i < n_levels; ++i) {
    const int rc = blobsplat_feature_splat(scores[i], stride_n[i], stride_k[i], stride_p[i], features[i], outs[i], N, K,
                                           C[i], H[i], W[i], dtype, engine, device, stream);
    if (rc != BLOBSPLAT_OK) return rc;
  }
  return BLOBSPLAT_OK;
}

int blobsplat_conditioning_fill(const void* scores, const void* features, void* out, int B, int K, int C, int h, int w,
                                int c_total, int c_off, int halves, int write_scores, int dtype, int device,
                                void* stream) {
  BS_CHECK_ARG(B >= 0 && K >= 1 && C >= 0 && h >= 1 && w >= 4 && (w % 4) == 0, "bad shape B=%d K=%d C=%d h=%d w=%d (w %% 4 == 0)", B, K, C, h, w);
  BS_CHECK_ARG(halves == 1 || halves == 2, "halves must be 1 or 2 (got %d)", halves);
  const int planes = (write_scores ? K : 0) + C;
  BS_CHECK_ARG(c_off >= 0 && c_off + planes <= c_total, "channels [%d, %d) outside the buffer's %d", c_off, c_off + planes, c_total);
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  if (B == 0 || planes == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(scores && out && (C == 0 || features), "NULL pointer");
  BS_CHECK_ARG(aligned_to(out, 16), "output buffer must be 16-byte aligned");
  BS_CHECK_ARG(B <= 65535, "batch too large");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return conditioning_fill_dispatch(scores, features, out, B, K, C, h, w, c_total, c_off, halves, write_scores, dtype,
                                    (cudaStream_t)stream);
}

int blobsplat_residual_inject(void* hidden, const void* residual, const float* scale_per_sample, float scale, int B, int C,
                              int H, int Wh, int Wr, int cols, int dtype, int device, void* stream) {
  BS_CHECK_ARG(B >= 0 && C >= 1 && H >= 1 && Wh >= 1 && Wr >= 1, "bad shape");
  BS_CHECK_ARG(cols >= 1 && cols <= Wh && cols <= Wr, "cols=%d must fit both widths (%d, %d)", cols, Wh, Wr);
  BS_CHECK_ARG(valid_dtype(dtype), "bad dtype %d", dtype);
  if (B == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(hidden && residual, "NULL pointer");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return residual_inject_dispatch(hidden, residual, scale_per_sample, scale, B, C, H, Wh, Wr, cols, dtype, (cudaStream_t)stream);
}

int blobsplat_conv_in_weights(const void* weight, const void* features, float* weff, int B, int O, int Cin, int lc, int C, int K,
                              int dtype, int device, void* stream) {
  BS_CHECK_ARG(B >= 0 && O >= 1 && lc >= 0 && C >= 1 && K >= 1, "bad shape B=%d O=%d lc=%d C=%d K=%d", B, O, lc, C, K);
  BS_CHECK_ARG(Cin == lc + 1 + C, "weight has %d input planes, expected lc + 1 + C = %d", Cin, lc + 1 + C);
  BS_CHECK_ARG(valid_dtype(dtype) && dtype != BLOBSPLAT_F64, "float32 / bfloat16 / float16 only");
  if (B == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(weight && features && weff, "NULL pointer");
  BS_CHECK_ARG(B <= 65535, "batch too large");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return conv_in_weights_dispatch(weight, features, weff, B, O, Cin, lc, C, K, dtype, (cudaStream_t)stream);
}

int blobsplat_conv_in_hoisted(const void* latents, const void* cond, const void* weight, const void* bias, const float* weff,
                              void* out, int B, int O, int Cin, int lc, int J, int h, int w, int halves, int dtype, int device,
                              void* stream) {
  BS_CHECK_ARG(B >= 0 && O >= 1 && lc >= 1 && J >= 1 && h >= 1 && w >= 1, "bad shape B=%d O=%d lc=%d J=%d h=%d w=%d", B, O, lc, J, h, w);
  BS_CHECK_ARG(halves == 1 || halves == 2, "halves must be 1 or 2 (got %d)", halves);
  BS_CHECK_ARG(Cin >= lc, "weight has %d input planes < lc = %d", Cin, lc);
  BS_CHECK_ARG(valid_dtype(dtype) && dtype != BLOBSPLAT_F64, "float32 / bfloat16 / float16 only");
  if (B == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(latents && cond && weight && weff && out, "NULL pointer");
  BS_CHECK_ARG(B <= 65535 && (O + 31) / 32 <= 65535, "batch / channel count too large");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return conv_in_hoisted_dispatch(latents, cond, weight, bias, weff, out, B, O, Cin, lc, J, h, w, halves * w, dtype,
                                  (cudaStream_t)stream);
}

int blobsplat_render(const float* xs, const float* ys, const float* covs, const float* sizes, const void* features,
                     int feat_dtype, int N, int M, int H, int W, int C, void* composed, void* grid, int out_dtype,
                     int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1 && C >= 1, "bad shape N=%d M=%d H=%d W=%d C=%d", N, M, H, W, C);
  BS_CHECK_ARG((long long)H * W < (1ll << 31), "H*W must fit in int32");
  BS_CHECK_ARG(valid_dtype(feat_dtype) && valid_dtype(out_dtype), "bad dtype");
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(M == 0 || (xs && ys && covs && sizes), "NULL blob parameter pointer");   // no blobs: background only
  BS_CHECK_ARG(features && grid, "NULL pointer");
  const char* why = nullptr;
  if (!render_tc_supported(M + 1, C, H, W, feat_dtype, out_dtype, &why)) BS_UNSUPPORTED("fused render: %s", why);
  if (out_dtype == BLOBSPLAT_F32 && f32_exact()) BS_UNSUPPORTED("fused render: BLOBSPLAT_F32_EXACT=1 keeps float32 on the FMA engine");
  DeviceGuard g(device);
  if (g.status) return g.status;
  return render_tc_dispatch(xs, ys, covs, sizes, features, feat_dtype, N, M, H, W, C, composed, grid, out_dtype,
                            (cudaStream_t)stream);
}

int blobsplat_render_small(const float* xs, const float* ys, const float* covs, const float* sizes, const float* features,
                           int N, int M, int H, int W, int C, float* composed, float* grid, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && H >= 1 && W >= 1 && C >= 1, "bad shape N=%d M=%d H=%d W=%d C=%d", N, M, H, W, C);
  BS_CHECK_ARG((long long)H * W < (1ll << 31) && N <= 65535 && (C + 15) / 16 <= 65535, "shape too large");
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(M == 0 || (xs && ys && covs && sizes), "NULL blob parameter pointer");
  BS_CHECK_ARG(features && grid, "NULL pointer");
  if (!render_small_supported(M + 1)) BS_UNSUPPORTED("latency render: K = %d planes (the pixel's weights live in registers: K <= 33)", M + 1);
  DeviceGuard g(device);
  if (g.status) return g.status;
  return render_small_dispatch(xs, ys, covs, sizes, features, N, M, H, W, C, composed, grid, (cudaStream_t)stream);
}

#ifndef BS_FUSED_PYRAMID_LEVELS
#define BS_FUSED_PYRAMID_LEVELS 3
#endif
#ifndef BS_FUSED_PYRAMID
#define BS_FUSED_PYRAMID 1      // multi-scale renders of 64 x 64 16-bit maps: the pyramid comes out of the render launch
#endif
int blobsplat_render_multiscale(const float* xs, const float* ys, const float* covs, const float* sizes, int N, int M,
                                int S, int n_levels, const void* const* features, const int* C, void* const* composed,
                                void* const* grids, int dtype, int device, void* stream) {
  BS_CHECK_ARG(N >= 0 && M >= 0 && S >= 1, "bad shape N=%d M=%d S=%d", N, M, S);
  BS_CHECK_ARG(n_levels >= 1 && n_levels <= 4, "1..4 levels per call (got %d)", n_levels);
  BS_CHECK_ARG((S % (1 << (n_levels - 1))) == 0, "S=%d is not divisible by 2^%d", S, n_levels - 1);
  BS_CHECK_ARG(valid_dtype(dtype) && dtype != BLOBSPLAT_F64, "float32 / bfloat16 / float16 maps only");
  BS_CHECK_ARG(features && C && composed && grids, "NULL level array");
  if (N == 0) return BLOBSPLAT_OK;
  BS_CHECK_ARG(features[0] && grids[0] && composed[0], "level 0 needs features, a grid and a composed-map buffer");
  for (int l = 1; l < n_levels; ++l) BS_CHECK_ARG(composed[l] != nullptr, "level %d composed-map buffer is NULL", l);
  // level 0: stages 1+2+3 fused; levels 1..: exact 2x2 means of the composed maps, then stage 3 per level
  int rc = 1;
  static const bool fused_pyr = [] { const char* e = std::getenv("BLOBSPLAT_FUSED_PYRAMID"); return !(e && e[0] == '0'); }();   // A/B knob
  if (n_levels > 1 && BS_FUSED_PYRAMID && fused_pyr && dtype != BLOBSPLAT_F32 && S == 64 && (M > 0 ? (xs && ys && covs && sizes) : true)) {
    // BlobNet's latent size in 16 bits: the pyramid leaves the render launch itself (render_tc2.cuh, kPyr)
    const char* why = nullptr;
    if (render_tc_supported(M + 1, C[0], S, S, dtype, dtype, &why)) {
      DeviceGuard g(device);
      if (g.status) return g.status;
      const int fused_levels = std::min(BS_FUSED_PYRAMID_LEVELS, n_levels - 1);
      rc = render_tc_pyramid_dispatch(xs, ys, covs, sizes, features[0], N, M, S, C[0], composed[0], grids[0], composed + 1,
                                      fused_levels, dtype, (cudaStream_t)stream);
      if (rc == 0 && n_levels - 1 > fused_levels)
        rc = blobsplat_pyramid(composed[fused_levels], composed + 1 + fused_levels, n_levels - 1 - fused_levels, N * (M + 1),
                               S >> fused_levels, dtype, device, stream);
      if (rc < 0) return rc;
    }
  }
  if (rc == 1) {
    rc = blobsplat_render(xs, ys, covs, sizes, features[0], dtype, N, M, S, S, C[0], composed[0], grids[0], dtype, device, stream);
    if (rc != BLOBSPLAT_OK || n_levels == 1) return rc;
    rc = blobsplat_pyramid(composed[0], composed + 1, n_levels - 1, N * (M + 1), S, dtype, device, stream);
    if (rc != BLOBSPLAT_OK) return rc;
  }
  if (n_levels == 1) return BLOBSPLAT_OK;
  const void* sc[4]; const void* ft[4]; void* out[4];
  int64_t sn[4], sk[4], sp[4];
  int Cs[4], Hs[4], Ws[4], n = 0;
  for (int l = 1; l < n_levels; ++l) {
    if (!features[l]) continue;                         // level wanted as maps only
    BS_CHECK_ARG(grids[l] != nullptr, "level %d has features but no grid buffer", l);
    const int s = S >> l;
    sc[n] = composed[l]; ft[n] = features[l]; out[n] = grids[l];
    sn[n] = (int64_t)(M + 1) * s * s; sk[n] = (int64_t)s * s; sp[n] = 1;
    Cs[n] = C[l]; Hs[n] = s; Ws[n] = s; ++n;
  }
  return blobsplat_feature_splat_levels(n, sc, sn, sk, sp, ft, out, N, M + 1, Cs, Hs, Ws, dtype, BLOBSPLAT_ENGINE_AUTO, device,
                                        stream);
}

}  // extern "C"
