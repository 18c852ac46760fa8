// One-image latency kernel: stages 1+2+3 of splat_features (blobctrl/utils/utils.py:80-241 at interp_size == score_size)
// for the single small renders the reference's scripts and UI issue (BASELINE config 2: 16 blobs x 64x64 x 320 channels,
// batch 1, float32) — on CUDA cores, with no tensor memory, no operand transposition and no split precision.
//
// At this size the fused tcgen05 render is a serial chain of small steps (profiles/latency_phases_r2.txt: 7.2 us in the
// kernel, 4.5 of them before the first opacity is evaluated: TMEM allocation, K-major 2 x fp16 operand staging, MMA commit
// round trips), while the arithmetic itself is 22 MFMA — a microsecond of the FP32 pipe.  Here
//   CTA = 128 pixels (thread = pixel) x CT channels of one image; grid = (pixel tiles, channel tiles, images); CT = 16, 32 or
//         64, the narrowest that keeps the grid at <= 1024 CTAs (cfg2: 32 x 20 = 640 CTAs of 16 channels — 8.2 us as a graph
//         replay against 10.3 / 12.3 us with 32 / 64 channels; the K = 2, C = 1024 splat prefers 32: profiles/ab_small_r2.txt)
//   blob coefficients: one thread per blob, float64 whitening (common.cuh::make_blob_coef), in shared memory
//   features [K, CT] of the channel tile: a coalesced copy into shared memory, in flight during the coefficient math
//   stages 1+2: front-to-back walk, the composed weights d_k of the pixel stay in REGISTERS (K <= KMAX, unrolled)
//   stage 3: acc[4 channels] += d_k * F[k][c..c+3] (one broadcast LDS.128 per 4 FFMA), full-fp32 products and sums like
//            the reference's einsum; plane stores are 128-byte lines per warp
// The composed maps are written by the CTAs of channel tile 0.  Same coefficients and opacity form as scores_lane_pixel_f32
// (which steps u, v along its four pixels: the maps agree to the last place, 5e-7).
#include "common.cuh"

namespace blobsplat {

constexpr int kRsThreads = 128;
constexpr int kRsMaxCtas = 1024;

template <int KMAX, int kRsCT>
__global__ void __launch_bounds__(kRsThreads)
render_small_kernel(const float* __restrict__ xs, const float* __restrict__ ys, const float* __restrict__ covs,
                    const float* __restrict__ sizes, const float* __restrict__ feats, int M, int H, int W, int C,
                    float* __restrict__ composed, float* __restrict__ grid) {
  __shared__ BlobCoef coef[KMAX];
  __shared__ __align__(16) float F[KMAX][kRsCT];
  const int n = blockIdx.z, c0 = blockIdx.y * kRsCT, K = M + 1, P = H * W;
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  // features of this channel tile (rows k < K, channels c0 .. c0 + kRsCT): issued first, they land during the float64 math
  const float* f = feats + (size_t)n * K * C;
  for (int i = tid; i < K * kRsCT; i += kRsThreads) {
    const int k = i / kRsCT, c = i - k * kRsCT;
    F[k][c] = (c0 + c < C) ? __ldg(f + (size_t)k * C + c0 + c) : 0.0f;
  }
  for (int i = tid; i < M; i += kRsThreads) {
    const size_t b = (size_t)n * M + i;
    const float* c = covs + 4 * b;
    coef[i] = make_blob_coef((double)xs[b], (double)ys[b], (double)c[0], (double)c[1], (double)c[2], (double)c[3],
                             sizes[b], H, W);
  }
  __syncthreads();
  const int pix = blockIdx.x * kRsThreads + tid;
  if (pix >= P) return;
  const int y = pix / W, x = pix - y * W;
  const float xf = (float)x, yf = (float)y;
  // stages 1+2 (utils.py:120-181): d_k = s_k * prod_{j > k} (1 - s_j), blob m is plane m + 1, plane 0 the background
  float d[KMAX];
  float T = 1.0f;
#pragma unroll
  for (int i = KMAX - 2; i >= 0; --i) {
    d[i + 1] = 0.0f;
    if (i < M) {
      const float s = blob_opacity(coef[i], xf, yf);
      d[i + 1] = s * T;
      T = fmaf(-s, T, T);
    }
  }
  d[0] = T;
  if (composed != nullptr && blockIdx.y == 0) {
    float* cp = composed + (size_t)n * K * P + pix;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
      if (k < K) __stcs(cp + (size_t)k * P, d[k]);
  }
  // stage 3 (utils.py:57-77): out[c, pixel] = sum_k d_k * f[k, c]
  float* gp = grid + ((size_t)n * C + c0) * P + pix;
#pragma unroll 2
  for (int c = 0; c < kRsCT; c += 4) {
    if (c0 + c >= C) break;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      if (k < K) {
        const float4 fv = *reinterpret_cast<const float4*>(&F[k][c]);
        acc.x = fmaf(d[k], fv.x, acc.x); acc.y = fmaf(d[k], fv.y, acc.y);
        acc.z = fmaf(d[k], fv.z, acc.z); acc.w = fmaf(d[k], fv.w, acc.w);
      }
    }
    __stcs(gp + (size_t)c * P, acc.x);
    if (c0 + c + 1 < C) __stcs(gp + (size_t)(c + 1) * P, acc.y);
    if (c0 + c + 2 < C) __stcs(gp + (size_t)(c + 2) * P, acc.z);
    if (c0 + c + 3 < C) __stcs(gp + (size_t)(c + 3) * P, acc.w);
  }
}

// Envelope of the latency kernel: float32 in and out, K = M + 1 <= 33 planes (the weights of a pixel live in registers).
bool render_small_supported(int K) { return K >= 1 && K <= 33; }

template <int KMAX>
static int launch_small(int ct, dim3 g, cudaStream_t st, const float* xs, const float* ys, const float* covs, const float* sizes,
                        const float* feats, int M, int H, int W, int C, float* composed, float* grid) {
  if (ct == 16) BS_CUDA(launch_pdl(render_small_kernel<KMAX, 16>, g, dim3(kRsThreads), 0, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid));
  else if (ct == 32) BS_CUDA(launch_pdl(render_small_kernel<KMAX, 32>, g, dim3(kRsThreads), 0, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid));
  else BS_CUDA(launch_pdl(render_small_kernel<KMAX, 64>, g, dim3(kRsThreads), 0, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid));
  return 0;
}

int render_small_dispatch(const float* xs, const float* ys, const float* covs, const float* sizes, const float* feats, int N,
                          int M, int H, int W, int C, float* composed, float* grid, cudaStream_t st) {
  const int K = M + 1, P = H * W;
  const long long px_tiles = (P + kRsThreads - 1) / kRsThreads;
  int ct = 16;                                                   // narrowest channel tile that keeps the grid small
  while (ct < 64 && px_tiles * ((C + ct - 1) / ct) * N > kRsMaxCtas) ct <<= 1;
  const dim3 g((unsigned)px_tiles, (unsigned)((C + ct - 1) / ct), (unsigned)N);
  if (K <= 9) return launch_small<9>(ct, g, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid);
  if (K <= 17) return launch_small<17>(ct, g, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid);
  return launch_small<33>(ct, g, st, xs, ys, covs, sizes, feats, M, H, W, C, composed, grid);
}

}  // namespace blobsplat
