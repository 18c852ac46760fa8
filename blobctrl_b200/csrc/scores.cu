// Stages 1+2 of the blob renderer: blob parameters -> raw opacity -> depth-ordered composite.
//
// Replaces blobctrl/utils/utils.py:120-194 of the reference (grid/delta/linalg.solve/sigmoid/clamp/
// where/cat/flip/cumprod/flip/roll/mul/setitem: ~18 ATen launches and six [N,M,2,P] / [N,P,K]
// HBM round trips) with one launch whose only HBM traffic is the 28-byte blob records in and the
// [N,K,H,W] planes out.
//
// Two stage-2 mappings are provided and measured against each other:
//   lane = pixel : each thread owns V horizontally adjacent pixels, walks k = M..1 with the
//                  transmittance T in a register and emits d_k = s_k*T as one 128-bit store per
//                  plane.  Serial order identical to the reference's cumprod.
//   lane = blob  : a warp owns 32 pixels; lanes hold 32 consecutive blobs of one pixel, an
//                  exclusive multiplicative suffix scan over __shfl_down_sync gives
//                  prod_{j>k}(1-s_j), chunks are chained from the high-index end with a broadcast
//                  carry, and the [K][32] tile goes through shared memory so the HBM stores stay
//                  pixel-contiguous per plane.
#include "common.cuh"

namespace blobsplat {

constexpr int kScoreThreads = 128;
constexpr int kBlobChunk = 128;  // blobs staged in shared memory at a time (4 KB)

__device__ __forceinline__ void plane_of(int k, int select, int& plane, bool& write) {
  // utils.py:183-191: ALL keeps k, FG drops the background, BG keeps only the background
  if (select == BLOBSPLAT_SELECT_ALL) { plane = k; write = true; }
  else if (select == BLOBSPLAT_SELECT_FG) { plane = k - 1; write = k >= 1; }
  else { plane = 0; write = k == 0; }
}

template <typename O, int V>
__global__ void __launch_bounds__(kScoreThreads)
scores_lane_pixel_f32(const float* __restrict__ xs, const float* __restrict__ ys,
                      const float* __restrict__ covs, const float* __restrict__ sizes, int M, int H, int W,
                      int select, int ksel, O* __restrict__ composed, O* __restrict__ raw,
                      const float* __restrict__ ell, float img_w, float img_h) {
  __shared__ BlobCoef coef[kBlobChunk];
  const int n = blockIdx.y;
  const int P = H * W;
  const int pix = (blockIdx.x * kScoreThreads + threadIdx.x) * V;
  const bool active = pix < P;
  const int y = active ? pix / W : 0;
  const int x = active ? pix - y * W : 0;
  const float xf = (float)x, yf = (float)y;

  float T[V];
#pragma unroll
  for (int j = 0; j < V; ++j) T[j] = 1.0f;

  O* cbase = composed ? composed + (size_t)n * ksel * P + pix : nullptr;
  O* rbase = raw ? raw + (size_t)n * (M + 1) * P + pix : nullptr;

  for (int hi = M; hi > 0; hi -= kBlobChunk) {
    const int lo = hi > kBlobChunk ? hi - kBlobChunk : 0;
    const int cnt = hi - lo;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kScoreThreads) {
      const size_t b = (size_t)n * M + lo + i;
      if (ell != nullptr) {   // ellipse front end: xs/ys/covs unused
        const float* e = ell + 5 * b;
        coef[i] = make_blob_coef_ellipse((double)e[0], (double)e[1], (double)e[2], (double)e[3], (double)e[4], sizes[b],
                                         (double)img_w, (double)img_h, H, W);
      } else {
        const float* c = covs + 4 * b;
        coef[i] = make_blob_coef((double)xs[b], (double)ys[b], (double)c[0], (double)c[1], (double)c[2],
                                 (double)c[3], sizes[b], H, W);
      }
    }
    __syncthreads();
    if (!active) continue;
#pragma unroll 2
    for (int i = cnt - 1; i >= 0; --i) {
      const BlobCoef c = coef[i];
      const int k = lo + i + 1;
      float s[V];
      if (coef_gated(c)) {
#pragma unroll
        for (int j = 0; j < V; ++j) s[j] = 1e-6f;
      } else {
        const float dyh = yf - c.cy_hi, dxh = xf - c.cx_hi;
        if (!coef_general(c)) {
          const float u0 = fmaf(c.p, dxh, c.u0);
          const float v0 = fmaf(c.r, dxh, fmaf(c.t, dyh, c.v0));
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float u = fmaf((float)j, c.p, u0);
            const float v = fmaf((float)j, c.r, v0);
            s[j] = opacity_from_q2m1(fmaf(u, u, fmaf(v, v, -1.0f)));
          }
        } else {
          const float dx0 = dxh - c.u0, dy = dyh - c.v0;     // general form: (u0, v0) = (cx_lo, cy_lo)
#pragma unroll
          for (int j = 0; j < V; ++j) {
            const float dx = dx0 + (float)j;
            s[j] = opacity_from_q2(fmaf(dx, fmaf(c.p, dx, c.r * dy), c.t * dy * dy));
          }
        }
      }
      float d[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        d[j] = s[j] * T[j];               // d_k = s_k * prod_{j>k}(1 - s_j)   (utils.py:180-181)
        T[j] = fmaf(-s[j], T[j], T[j]);   // T <- T * (1 - s_k)
      }
      int plane; bool wr;
      plane_of(k, select, plane, wr);
      if (cbase && wr) VecStore<O, V>::st(cbase + (size_t)plane * P, d);
      if (rbase) VecStore<O, V>::st(rbase + (size_t)k * P, s);
    }
  }
  if (!active) return;
  // background: alpha 1 (utils.py:175-176) -> d_0 = prod_{j>=1}(1 - s_j)
  if (cbase && select != BLOBSPLAT_SELECT_FG) VecStore<O, V>::st(cbase, T);
  if (rbase) {
    float one[V];
#pragma unroll
    for (int j = 0; j < V; ++j) one[j] = 1.0f;
    VecStore<O, V>::st(rbase, one);
  }
}

// ---- float64 parameters (what the reference's scripts actually feed: blobctrl_inference.py:104-106) --
struct BlobCoefD {
  double cx, cy, A, B2, C;
  uint32_t gated;
};

template <int V>
__global__ void __launch_bounds__(kScoreThreads)
scores_lane_pixel_f64(const double* __restrict__ xs, const double* __restrict__ ys,
                      const double* __restrict__ covs, const float* __restrict__ sizes, int M, int H, int W,
                      int select, int ksel, double* __restrict__ composed, double* __restrict__ raw) {
  __shared__ BlobCoefD coef[kBlobChunk];
  const int n = blockIdx.y;
  const int P = H * W;
  const int pix = (blockIdx.x * kScoreThreads + threadIdx.x) * V;
  const bool active = pix < P;
  const int y = active ? pix / W : 0;
  const int x = active ? pix - y * W : 0;
  double T[V];
#pragma unroll
  for (int j = 0; j < V; ++j) T[j] = 1.0;
  double* cbase = composed ? composed + (size_t)n * ksel * P + pix : nullptr;
  double* rbase = raw ? raw + (size_t)n * (M + 1) * P + pix : nullptr;

  for (int hi = M; hi > 0; hi -= kBlobChunk) {
    const int lo = hi > kBlobChunk ? hi - kBlobChunk : 0;
    const int cnt = hi - lo;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kScoreThreads) {
      const size_t b = (size_t)n * M + lo + i;
      const double* c = covs + 4 * b;
      const double det = c[0] * c[3] - c[1] * c[2];
      BlobCoefD o;
      o.cx = xs[b] * (double)W; o.cy = ys[b] * (double)H;
      o.A = (c[3] / det) / ((double)W * (double)W);
      o.B2 = (-(c[1] + c[2]) / det) / ((double)W * (double)H);
      o.C = (c[0] / det) / ((double)H * (double)H);
      o.gated = sizes[b] < 0.5f;
      coef[i] = o;
    }
    __syncthreads();
    if (!active) continue;
    for (int i = cnt - 1; i >= 0; --i) {
      const BlobCoefD c = coef[i];
      const int k = lo + i + 1;
      double s[V], d[V];
      const double dy = (double)y - c.cy;
#pragma unroll
      for (int j = 0; j < V; ++j) {
        const double dx = (double)(x + j) - c.cx;
        const double q = dx * (c.A * dx + c.B2 * dy) + c.C * dy * dy;
        const double sv = fmin(2.0 / (1.0 + exp(q)), 1.0);
        s[j] = c.gated ? (double)1e-6f : sv;   // 1e-6 is a float32 scalar in the reference (utils.py:172)
        d[j] = s[j] * T[j];
        T[j] = T[j] * (1.0 - s[j]);
      }
      int plane; bool wr;
      plane_of(k, select, plane, wr);
      if (cbase && wr) VecStore<double, V>::st(cbase + (size_t)plane * P, d);
      if (rbase) VecStore<double, V>::st(rbase + (size_t)k * P, s);
    }
  }
  if (!active) return;
  if (cbase && select != BLOBSPLAT_SELECT_FG) VecStore<double, V>::st(cbase, T);
  if (rbase) {
    double one[V];
#pragma unroll
    for (int j = 0; j < V; ++j) one[j] = 1.0;
    VecStore<double, V>::st(rbase, one);
  }
}

// ---- lane = blob: warp-level multiplicative suffix scan across blobs ---------------------------------
constexpr int kScanWarps = 4;
constexpr int kScanMaxBlobs = 256;
constexpr int kTileStride = 33;  // [K][32] tile padded to 33 floats: conflict-free column writes

template <typename O>
__global__ void __launch_bounds__(kScanWarps * 32)
scores_warp_scan_f32(const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ covs,
                     const float* __restrict__ sizes, int M, int H, int W, int select, int ksel,
                     O* __restrict__ composed) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  BlobCoef* coef = reinterpret_cast<BlobCoef*>(smem_raw);
  float* tiles = reinterpret_cast<float*>(smem_raw + sizeof(BlobCoef) * M);
  const int n = blockIdx.y;
  const int P = H * W;
  const int K = M + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    const size_t b = (size_t)n * M + i;
    const float* c = covs + 4 * b;
    coef[i] = make_blob_coef((double)xs[b], (double)ys[b], (double)c[0], (double)c[1], (double)c[2],
                             (double)c[3], sizes[b], H, W);
  }
  __syncthreads();
  float* tile = tiles + (size_t)warp * K * kTileStride;
  const int base = (blockIdx.x * kScanWarps + warp) * 32;
  if (base >= P) return;
  const int npx = min(32, P - base);
  const int nchunks = (M + 31) >> 5;

  for (int px = 0; px < npx; ++px) {
    const int pix = base + px;
    const int y = pix / W;
    const float xf = (float)(pix - y * W), yf = (float)y;
    float carry = 1.0f;  // prod of (1 - s_j) over all blobs in higher chunks
    for (int ch = nchunks - 1; ch >= 0; --ch) {
      const int m = ch * 32 + lane;
      const bool valid = m < M;
      const float s = valid ? blob_opacity(coef[m], xf, yf) : 0.0f;
      float suf = 1.0f - s;  // inclusive suffix product over lanes >= lane
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const float o = __shfl_down_sync(0xffffffffu, suf, off);
        if (lane + off < 32) suf *= o;
      }
      float excl = __shfl_down_sync(0xffffffffu, suf, 1);  // prod over lanes > lane
      if (lane == 31) excl = 1.0f;
      if (valid) tile[(m + 1) * kTileStride + px] = s * (excl * carry);
      carry *= __shfl_sync(0xffffffffu, suf, 0);
    }
    if (lane == 0) tile[px] = carry;  // background plane: alpha 1 * total transmittance
  }
  __syncwarp();
  if (lane < npx) {
    O* out = composed + (size_t)n * ksel * P + base + lane;
    for (int k = 0; k < K; ++k) {
      int plane; bool wr;
      plane_of(k, select, plane, wr);
      if (wr) out[(size_t)plane * P] = Cvt<O>::from(tile[k * kTileStride + lane]);
    }
  }
}

// ---- preview: stages 1+2 + the C = 3 colour splat in ONE launch ----------------------------------------
// The UI preview (scripts/blobctrl_app.py:637-650 -> utils.py:198-223 with only_vis=True) needs only
// feature_img[n, ch, y, x] = sum_k d_k * colour[k, ch] (utils.py:244-270 with an identity viz_score_fn): the
// composed maps never have to exist.  One thread = V adjacent pixels, front-to-back walk with the
// transmittance in a register (as above), the colour table in shared memory.  PT = float: whitened
// coefficients (common.cuh); PT = double: the reference's fp64 arithmetic, as scores_lane_pixel_f64.
template <typename PT, int V>
__global__ void __launch_bounds__(kScoreThreads)
preview_kernel(const PT* __restrict__ xs, const PT* __restrict__ ys, const PT* __restrict__ covs,
               const float* __restrict__ sizes, const PT* __restrict__ colors, int colors_per_image, int M, int H, int W,
               PT* __restrict__ image, PT* __restrict__ composed, unsigned char* __restrict__ image_u8) {
  constexpr bool kF64 = sizeof(PT) == 8;
  using Coef = typename std::conditional<kF64, BlobCoefD, BlobCoef>::type;
  __shared__ Coef coef[kBlobChunk];
  __shared__ PT col[kBlobChunk + 1][3];
  const int n = blockIdx.y;
  const int P = H * W;
  const int pix = (blockIdx.x * kScoreThreads + threadIdx.x) * V;
  const bool active = pix < P;
  const int y = active ? pix / W : 0;
  const int x = active ? pix - y * W : 0;
  const PT* cb = colors + (colors_per_image ? (size_t)n * (M + 1) * 3 : 0);
  PT T[V], rgb[3][V];
#pragma unroll
  for (int j = 0; j < V; ++j) { T[j] = (PT)1; rgb[0][j] = rgb[1][j] = rgb[2][j] = (PT)0; }
  PT* cbase = composed ? composed + (size_t)n * (M + 1) * P + pix : nullptr;

  for (int hi = M; hi > 0; hi -= kBlobChunk) {
    const int lo = hi > kBlobChunk ? hi - kBlobChunk : 0;
    const int cnt = hi - lo;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += kScoreThreads) {
      const size_t b = (size_t)n * M + lo + i;
      const PT* c = covs + 4 * b;
      if constexpr (kF64) {
        const double det = c[0] * c[3] - c[1] * c[2];
        BlobCoefD o;
        o.cx = xs[b] * (double)W; o.cy = ys[b] * (double)H;
        o.A = (c[3] / det) / ((double)W * (double)W);
        o.B2 = (-(c[1] + c[2]) / det) / ((double)W * (double)H);
        o.C = (c[0] / det) / ((double)H * (double)H);
        o.gated = sizes[b] < 0.5f;
        coef[i] = o;
      } else {
        coef[i] = make_blob_coef((double)xs[b], (double)ys[b], (double)c[0], (double)c[1], (double)c[2], (double)c[3],
                                 sizes[b], H, W);
      }
      for (int ch = 0; ch < 3; ++ch) col[i][ch] = cb[(size_t)(lo + i + 1) * 3 + ch];
    }
    __syncthreads();
    if (!active) continue;
    for (int i = cnt - 1; i >= 0; --i) {
      const Coef c = coef[i];
      PT d[V];
#pragma unroll
      for (int j = 0; j < V; ++j) {
        PT s;
        if constexpr (kF64) {
          const double dx = (double)(x + j) - c.cx, dy = (double)y - c.cy;
          const double q = dx * (c.A * dx + c.B2 * dy) + c.C * dy * dy;
          s = c.gated ? (double)1e-6f : fmin(2.0 / (1.0 + exp(q)), 1.0);
          d[j] = s * T[j];
          T[j] = T[j] * (1.0 - s);
        } else {
          s = blob_opacity(c, (float)(x + j), (float)y);
          d[j] = s * T[j];
          T[j] = fmaf(-s, T[j], T[j]);
        }
#pragma unroll
        for (int ch = 0; ch < 3; ++ch) rgb[ch][j] += d[j] * col[i][ch];
      }
      if (cbase) VecStore<PT, V>::st(cbase + (size_t)(lo + i + 1) * P, d);
    }
  }
  if (!active) return;
  if (cbase) VecStore<PT, V>::st(cbase, T);
  unsigned char px8[3 * V];
#pragma unroll
  for (int ch = 0; ch < 3; ++ch) {
    PT v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) {
      v[j] = rgb[ch][j] + T[j] * cb[ch];                                    // background: alpha 1 * transmittance
      // the app's conversion (blobctrl_app.py:645-646): (img * 255).astype(np.uint8) = truncation of the product in PT
      px8[3 * j + ch] = (unsigned char)(int)(v[j] * (PT)255);
    }
    if (image) VecStore<PT, V>::st(image + ((size_t)n * 3 + ch) * P + pix, v);
  }
  if (image_u8) {                                                           // [N, H, W, 3]: this thread's 3 V bytes are contiguous
    unsigned char* const o = image_u8 + ((size_t)n * P + pix) * 3;
    if constexpr ((3 * V) % 4 == 0) {
#pragma unroll
      for (int w = 0; w < 3 * V / 4; ++w)
        reinterpret_cast<unsigned int*>(o)[w] = px8[4 * w] | (px8[4 * w + 1] << 8) | (px8[4 * w + 2] << 16) | ((unsigned)px8[4 * w + 3] << 24);
    } else if constexpr ((3 * V) % 2 == 0) {
#pragma unroll
      for (int w = 0; w < 3 * V / 2; ++w) reinterpret_cast<unsigned short*>(o)[w] = (unsigned short)(px8[2 * w] | (px8[2 * w + 1] << 8));
    } else {
#pragma unroll
      for (int w = 0; w < 3 * V; ++w) o[w] = px8[w];
    }
  }
}

// ---- composite only (viz_score_fn branch, utils.py:205-206) -------------------------------------------
template <typename T, typename A>
__global__ void __launch_bounds__(256)
composite_kernel(const T* __restrict__ in, T* __restrict__ out, int K, int P) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= P) return;
  const T* s = in + (size_t)n * K * P + pix;
  T* d = out + (size_t)n * K * P + pix;
  A trans = (A)1;
  for (int k = K - 1; k >= 0; --k) {
    const A sv = (A)Cvt<T>::to(s[(size_t)k * P]);
    d[(size_t)k * P] = Cvt<T>::from(k == K - 1 ? sv : sv * trans);
    trans = trans * ((A)1 - sv);
  }
}

// ---- host launchers ----------------------------------------------------------------------------------
template <typename O>
static int launch_lane_pixel_f32(const float* xs, const float* ys, const float* covs, const float* sizes, int N,
                                 int M, int H, int W, int select, int ksel, void* composed, void* raw,
                                 cudaStream_t st, const float* ell = nullptr, float img_w = 0.f, float img_h = 0.f) {
  constexpr int VMAX = Vec128<O>::n;
  const int P = H * W;
  const bool vec = (W % VMAX == 0) && aligned_to(composed, 16) && aligned_to(raw, 16);
  const int V = vec ? VMAX : 1;
  dim3 grid((unsigned)((P / V + kScoreThreads - 1) / kScoreThreads), (unsigned)N);
  if (vec)
    scores_lane_pixel_f32<O, VMAX><<<grid, kScoreThreads, 0, st>>>(xs, ys, covs, sizes, M, H, W, select, ksel,
                                                                  (O*)composed, (O*)raw, ell, img_w, img_h);
  else
    scores_lane_pixel_f32<O, 1><<<grid, kScoreThreads, 0, st>>>(xs, ys, covs, sizes, M, H, W, select, ksel,
                                                               (O*)composed, (O*)raw, ell, img_w, img_h);
  BS_CUDA(cudaGetLastError());
  return 0;
}

template <typename O>
static int launch_warp_scan_f32(const float* xs, const float* ys, const float* covs, const float* sizes, int N,
                                int M, int H, int W, int select, int ksel, void* composed, cudaStream_t st) {
  const int P = H * W;
  const size_t smem = sizeof(BlobCoef) * M + sizeof(float) * (size_t)kScanWarps * (M + 1) * kTileStride;
  static thread_local int configured_dev = -1;  // opt in to >48 KB dynamic smem once per device/thread
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(scores_warp_scan_f32<O>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    configured_dev = dev;
  }
  dim3 grid((unsigned)((P + kScanWarps * 32 - 1) / (kScanWarps * 32)), (unsigned)N);
  scores_warp_scan_f32<O><<<grid, kScanWarps * 32, smem, st>>>(xs, ys, covs, sizes, M, H, W, select, ksel,
                                                              (O*)composed);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int scores_dispatch(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype, int N,
                    int M, int H, int W, int select, void* composed, int composed_dtype, void* raw, int raw_dtype,
                    int mode, cudaStream_t st) {
  const int ksel = select == BLOBSPLAT_SELECT_ALL ? M + 1 : (select == BLOBSPLAT_SELECT_FG ? M : 1);
  if (param_dtype == BLOBSPLAT_F64) {
    if ((composed && composed_dtype != BLOBSPLAT_F64) || (raw && raw_dtype != BLOBSPLAT_F64))
      BS_UNSUPPORTED("float64 parameters produce float64 maps (the reference's dtype rule); got out dtype %d/%d",
                     composed_dtype, raw_dtype);
    if (mode == BLOBSPLAT_COMPOSITE_WARP_SCAN) BS_UNSUPPORTED("warp-scan composite is float32-only");
    const int P = H * W;
    const bool vec = (W % 2 == 0) && aligned_to(composed, 16) && aligned_to(raw, 16);
    dim3 grid((unsigned)((P / (vec ? 2 : 1) + kScoreThreads - 1) / kScoreThreads), (unsigned)N);
    if (vec)
      scores_lane_pixel_f64<2><<<grid, kScoreThreads, 0, st>>>((const double*)xs, (const double*)ys,
                                                               (const double*)covs, sizes, M, H, W, select, ksel,
                                                               (double*)composed, (double*)raw);
    else
      scores_lane_pixel_f64<1><<<grid, kScoreThreads, 0, st>>>((const double*)xs, (const double*)ys,
                                                               (const double*)covs, sizes, M, H, W, select, ksel,
                                                               (double*)composed, (double*)raw);
    BS_CUDA(cudaGetLastError());
    return 0;
  }
  if (param_dtype != BLOBSPLAT_F32) BS_UNSUPPORTED("blob parameters must be float32 or float64 (got %d)", param_dtype);
  if (composed && raw && composed_dtype != raw_dtype)
    BS_UNSUPPORTED("composed and raw maps must share a dtype (got %d and %d)", composed_dtype, raw_dtype);
  const int odt = composed ? composed_dtype : raw_dtype;
  if (odt == BLOBSPLAT_F64) BS_UNSUPPORTED("float64 maps need float64 parameters");
  const float* fx = (const float*)xs; const float* fy = (const float*)ys; const float* fc = (const float*)covs;

  if (mode == BLOBSPLAT_COMPOSITE_WARP_SCAN) {
    if (raw) BS_UNSUPPORTED("warp-scan composite does not emit raw scores; use LANE_PIXEL or AUTO");
    if (!composed) BS_CHECK_ARG(false, "warp-scan composite needs a composed output");
    if (M < 1 || M > kScanMaxBlobs) BS_UNSUPPORTED("warp-scan composite supports 1 <= M <= %d (got %d)", kScanMaxBlobs, M);
    switch (odt) {
      case BLOBSPLAT_F32: return launch_warp_scan_f32<float>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, st);
      case BLOBSPLAT_BF16: return launch_warp_scan_f32<__nv_bfloat16>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, st);
      case BLOBSPLAT_F16: return launch_warp_scan_f32<__half>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, st);
    }
  }
  // AUTO: lane = pixel.  Measured faster at every benchmark shape (profiles/): stage 1 already walks the
  // blobs per pixel, so the scan adds shuffles and a shared-memory transpose without removing work.
  switch (odt) {
    case BLOBSPLAT_F32: return launch_lane_pixel_f32<float>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, raw, st);
    case BLOBSPLAT_BF16: return launch_lane_pixel_f32<__nv_bfloat16>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, raw, st);
    case BLOBSPLAT_F16: return launch_lane_pixel_f32<__half>(fx, fy, fc, sizes, N, M, H, W, select, ksel, composed, raw, st);
  }
  BS_UNSUPPORTED("unknown output dtype %d", odt);
}

int scores_ellipse_dispatch(const float* ell, const float* sizes, float img_w, float img_h, int N, int M, int H, int W,
                            int select, void* composed, int composed_dtype, void* raw, int raw_dtype, cudaStream_t st) {
  const int ksel = select == BLOBSPLAT_SELECT_ALL ? M + 1 : (select == BLOBSPLAT_SELECT_FG ? M : 1);
  if (composed && raw && composed_dtype != raw_dtype) BS_UNSUPPORTED("composed and raw maps must share a dtype");
  const int odt = composed ? composed_dtype : raw_dtype;
  switch (odt) {
    case BLOBSPLAT_F32: return launch_lane_pixel_f32<float>(nullptr, nullptr, nullptr, sizes, N, M, H, W, select, ksel, composed, raw, st, ell, img_w, img_h);
    case BLOBSPLAT_BF16: return launch_lane_pixel_f32<__nv_bfloat16>(nullptr, nullptr, nullptr, sizes, N, M, H, W, select, ksel, composed, raw, st, ell, img_w, img_h);
    case BLOBSPLAT_F16: return launch_lane_pixel_f32<__half>(nullptr, nullptr, nullptr, sizes, N, M, H, W, select, ksel, composed, raw, st, ell, img_w, img_h);
  }
  BS_UNSUPPORTED("ellipse front end renders float32/bfloat16/float16 maps (got dtype %d)", odt);
}

int preview_dispatch(const void* xs, const void* ys, const void* covs, const float* sizes, int param_dtype, const void* colors,
                     int colors_per_image, int N, int M, int H, int W, void* image, void* composed, unsigned char* image_u8,
                     cudaStream_t st) {
  const int P = H * W;
  if (param_dtype == BLOBSPLAT_F64) {
    const bool vec = (W % 2 == 0) && aligned_to(image, 16) && aligned_to(composed, 16) && aligned_to(image_u8, 2);
    dim3 grid((unsigned)((P / (vec ? 2 : 1) + kScoreThreads - 1) / kScoreThreads), (unsigned)N);
    if (vec)
      preview_kernel<double, 2><<<grid, kScoreThreads, 0, st>>>((const double*)xs, (const double*)ys, (const double*)covs, sizes,
                                                                (const double*)colors, colors_per_image, M, H, W, (double*)image, (double*)composed, image_u8);
    else
      preview_kernel<double, 1><<<grid, kScoreThreads, 0, st>>>((const double*)xs, (const double*)ys, (const double*)covs, sizes,
                                                                (const double*)colors, colors_per_image, M, H, W, (double*)image, (double*)composed, image_u8);
  } else if (param_dtype == BLOBSPLAT_F32) {
    const bool vec = (W % 4 == 0) && aligned_to(image, 16) && aligned_to(composed, 16) && aligned_to(image_u8, 4);
    dim3 grid((unsigned)((P / (vec ? 4 : 1) + kScoreThreads - 1) / kScoreThreads), (unsigned)N);
    if (vec)
      preview_kernel<float, 4><<<grid, kScoreThreads, 0, st>>>((const float*)xs, (const float*)ys, (const float*)covs, sizes,
                                                               (const float*)colors, colors_per_image, M, H, W, (float*)image, (float*)composed, image_u8);
    else
      preview_kernel<float, 1><<<grid, kScoreThreads, 0, st>>>((const float*)xs, (const float*)ys, (const float*)covs, sizes,
                                                               (const float*)colors, colors_per_image, M, H, W, (float*)image, (float*)composed, image_u8);
  } else {
    BS_UNSUPPORTED("preview renders float32 or float64 (got %d)", param_dtype);
  }
  BS_CUDA(cudaGetLastError());
  return 0;
}

int composite_dispatch(const void* in, void* out, int N, int K, int H, int W, int dtype, cudaStream_t st) {
  const int P = H * W;
  dim3 grid((unsigned)((P + 255) / 256), (unsigned)N);
  switch (dtype) {
    case BLOBSPLAT_F32: composite_kernel<float, float><<<grid, 256, 0, st>>>((const float*)in, (float*)out, K, P); break;
    case BLOBSPLAT_F64: composite_kernel<double, double><<<grid, 256, 0, st>>>((const double*)in, (double*)out, K, P); break;
    case BLOBSPLAT_BF16: composite_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>((const __nv_bfloat16*)in, (__nv_bfloat16*)out, K, P); break;
    case BLOBSPLAT_F16: composite_kernel<__half, float><<<grid, 256, 0, st>>>((const __half*)in, (__half*)out, K, P); break;
    default: BS_UNSUPPORTED("unknown dtype %d", dtype);
  }
  BS_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace blobsplat
