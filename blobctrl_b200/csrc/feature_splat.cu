// Stage 3 on CUDA cores: out[n,c,p] = sum_k w[n,k,p] * f[n,k,c]  (NCHW, contiguous).
//
// Replaces splat_features_from_scores (blobctrl/utils/utils.py:57-77; duplicate method at
// blobctrl/pipelines/pipeline_blobnet.py:706-721): einsum('nmhw,nmc->nchw') -> bmm + .contiguous().
//
// This is the general engine: any K, any C, any of f32/f64/bf16/f16, arbitrary score strides
// ([N,K,H,W] or [N,H,W,K]).  Each CTA owns 32*PX pixels x 64 channels; a warp = one 8-channel group,
// a lane = PX adjacent pixels, so every output plane row is written as one 128-bit store per lane
// (512 B contiguous per warp per channel).  The contraction runs out of shared memory: per k one
// conflict-free LDS.128 of weights + two broadcast LDS.128 of features feed 8*PX FFMAs.
// It is FP32-FMA-bound for K >~ 16; the tensor-core render in render_tc.cu is the
// engine for the large-K benchmark shapes.
#include "common.cuh"

namespace blobsplat {

constexpr int kFsThreads = 256;
constexpr int kFsTC = 64;  // channels per CTA (8 warps x 8)
constexpr int kFsKC = 16;  // k-chunk staged in shared memory

template <typename T, typename A, int PX>
__global__ void __launch_bounds__(kFsThreads)
feature_splat_fma(const T* __restrict__ scores, long long sn, long long sk, long long sp,
                  const T* __restrict__ feats, T* __restrict__ out, int K, int C, int P, bool vec_ok) {
  constexpr int TP = 32 * PX;
  __shared__ __align__(16) A w_s[kFsKC][TP];
  __shared__ __align__(16) A f_s[kFsKC][kFsTC];
  const int n = blockIdx.z;
  const int p0 = blockIdx.x * TP;
  const int c0 = blockIdx.y * kFsTC;
  const int lane = threadIdx.x & 31, cg = threadIdx.x >> 5;

  A acc[8][PX];
#pragma unroll
  for (int c = 0; c < 8; ++c)
#pragma unroll
    for (int j = 0; j < PX; ++j) acc[c][j] = (A)0;

  const T* sbase = scores + (size_t)n * sn;
  const T* fbase = feats + (size_t)n * K * C;

  for (int k0 = 0; k0 < K; k0 += kFsKC) {
    __syncthreads();
    for (int i = threadIdx.x; i < kFsKC * TP; i += kFsThreads) {
      const int kk = i / TP, pp = i - kk * TP;
      const int k = k0 + kk, p = p0 + pp;
      w_s[kk][pp] = (k < K && p < P) ? (A)Cvt<T>::to(sbase[(size_t)k * sk + (size_t)p * sp]) : (A)0;
    }
    for (int i = threadIdx.x; i < kFsKC * kFsTC; i += kFsThreads) {
      const int kk = i / kFsTC, cc = i - kk * kFsTC;
      const int k = k0 + kk, c = c0 + cc;
      f_s[kk][cc] = (k < K && c < C) ? (A)Cvt<T>::to(fbase[(size_t)k * C + c]) : (A)0;
    }
    __syncthreads();
    const int kcnt = min(kFsKC, K - k0);      // small K (the pipeline's K = 1 splat): no work on the zero padding
#pragma unroll 4
    for (int kk = 0; kk < kcnt; ++kk) {
      A wv[PX], fv[8];
#pragma unroll
      for (int j = 0; j < PX; ++j) wv[j] = w_s[kk][lane * PX + j];
#pragma unroll
      for (int c = 0; c < 8; ++c) fv[c] = f_s[kk][cg * 8 + c];
#pragma unroll
      for (int c = 0; c < 8; ++c)
#pragma unroll
        for (int j = 0; j < PX; ++j) acc[c][j] = fma(fv[c], wv[j], acc[c][j]);
    }
  }

  const int p = p0 + lane * PX;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ch = c0 + cg * 8 + c;
    if (ch >= C) break;
    T* dst = out + ((size_t)n * C + ch) * P + p;
    if (vec_ok && p + PX <= P) {
      if constexpr (sizeof(T) == 8) {  // two 128-bit stores for double
#pragma unroll
        for (int j = 0; j < PX; j += 2) VecStore<T, 2>::st(dst + j, &acc[c][j]);
      } else {
        VecStore<T, PX>::st(dst, acc[c]);
      }
    } else {
#pragma unroll
      for (int j = 0; j < PX; ++j)
        if (p + j < P) dst[j] = Cvt<T>::from(acc[c][j]);
    }
  }
}

template <typename T, typename A, int PX>
static int launch_fs(const void* scores, int64_t sn, int64_t sk, int64_t sp, const void* feats, void* out, int N,
                     int K, int C, int P, cudaStream_t st) {
  constexpr int TP = 32 * PX;
  const bool vec_ok = (P % PX == 0) && aligned_to(out, 16);
  dim3 grid((unsigned)((P + TP - 1) / TP), (unsigned)((C + kFsTC - 1) / kFsTC), (unsigned)N);
  feature_splat_fma<T, A, PX><<<grid, kFsThreads, 0, st>>>((const T*)scores, sn, sk, sp, (const T*)feats, (T*)out,
                                                          K, C, P, vec_ok);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int feature_splat_fma_dispatch(const void* scores, int64_t sn, int64_t sk, int64_t sp, const void* feats, void* out,
                               int N, int K, int C, int H, int W, int dtype, cudaStream_t st) {
  const int P = H * W;
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_fs<float, float, 4>(scores, sn, sk, sp, feats, out, N, K, C, P, st);
    case BLOBSPLAT_F64: return launch_fs<double, double, 2>(scores, sn, sk, sp, feats, out, N, K, C, P, st);
    case BLOBSPLAT_BF16: return launch_fs<__nv_bfloat16, float, 8>(scores, sn, sk, sp, feats, out, N, K, C, P, st);
    case BLOBSPLAT_F16: return launch_fs<__half, float, 8>(scores, sn, sk, sp, feats, out, N, K, C, P, st);
  }
  BS_UNSUPPORTED("unknown dtype %d", dtype);
}

}  // namespace blobsplat
