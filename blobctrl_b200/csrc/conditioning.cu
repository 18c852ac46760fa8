// N2: fused construct_blobnet_input.
//
// Replaces, for the loop-invariant part of BlobNet's input, the per-step
//   right = cat([latent_model_input, gs_scores, gs_feats], 1); left = cat([image_latents, gs_scores, gs_feats], 1);
//   blobnet_model_input = cat([left, right], -1)           (blobctrl/pipelines/pipeline_blobnet.py:724-739, called
//   twice per denoising step at :1043-1049 and :1071-1076) and the stage-3 splat that produced gs_feats (:984).
// The score plane and the C feature planes (sum_k s_k * f[k,c]) are written ONCE, straight into both width halves of
// a persistent [B, c_total, h, 2w] buffer at a channel offset; per step only the 4 latent channels are refreshed.
// Pure streaming stores (K is 1 in the pipeline): one thread = 4 adjacent pixels of one output channel, 128-bit
// (64-bit for 16-bit dtypes) stores to the left and the right half.
#include "common.cuh"

namespace blobsplat {

// V = pixels per thread: 4, or 8 for 16-bit dtypes when w % 8 == 0 (16-byte loads and stores)
template <typename T, int V>
__global__ void __launch_bounds__(256)
conditioning_fill_kernel(const T* __restrict__ scores, const T* __restrict__ feats, T* __restrict__ out, int K, int C,
                         int h, int w, int c_total, int c_off, int halves, bool write_scores) {
  // out[b, c_off + j, y, x + half*w]:  j < K_s (score planes, when write_scores) then C feature planes
  const int b = blockIdx.z;
  const int planes = (write_scores ? K : 0) + C;
  const int P = h * w, w4 = w / V;
  const bool vec_in = (reinterpret_cast<uintptr_t>(scores) & 15) == 0;    // rows are V-aligned (w % V == 0)
  const int items_per_plane = h * w4;
  const long long total = (long long)planes * items_per_plane;
  const int Wt = w * halves;
  const T* sb = scores + (size_t)b * K * P;
  const T* fb = feats ? feats + (size_t)b * K * C : nullptr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int plane = (int)(i / items_per_plane);
    const int r = (int)(i - (long long)plane * items_per_plane);
    const int y = r / w4, x = (r - y * w4) * V;
    using VecT = typename std::conditional<sizeof(T) * V == 16, uint4, uint2>::type;   // V pixels of T
    float v[V];
#pragma unroll
    for (int j = 0; j < V; ++j) v[j] = 0.f;
    const bool is_score = write_scores && plane < K;
    const int c = plane - (write_scores ? K : 0);
    for (int k = is_score ? plane : 0; k < (is_score ? plane + 1 : K); ++k) {
      const float f = is_score ? 1.0f : (float)Cvt<T>::to(fb[(size_t)k * C + c]);
      const T* sp = sb + (size_t)k * P + y * w + x;      // V adjacent pixels of score plane k
      T px[V];
      if (vec_in) {
        *reinterpret_cast<VecT*>(px) = __ldg(reinterpret_cast<const VecT*>(sp));
      } else {
#pragma unroll
        for (int j = 0; j < V; ++j) px[j] = sp[j];
      }
#pragma unroll
      for (int j = 0; j < V; ++j) v[j] = is_score ? (float)Cvt<T>::to(px[j]) : fmaf((float)Cvt<T>::to(px[j]), f, v[j]);
    }
    T* o = out + (((size_t)b * c_total + c_off + plane) * h + y) * Wt + x;
    T t[V];
#pragma unroll
    for (int j = 0; j < V; ++j) t[j] = Cvt<T>::from(v[j]);
    for (int hf = 0; hf < halves; ++hf) *reinterpret_cast<VecT*>(o + hf * w) = *reinterpret_cast<const VecT*>(t);
  }
}

template <typename T>
static int launch_fill(const void* scores, const void* feats, void* out, int B, int K, int C, int h, int w, int c_total,
                       int c_off, int halves, bool write_scores, cudaStream_t st) {
  // scores rows must be aligned to the vector: capi checks w % 4 == 0 and a 16-byte aligned output
  const bool wide = sizeof(T) == 2 && (w % 8) == 0;
  const int V = wide ? 8 : 4;
  const long long total = (long long)((write_scores ? K : 0) + C) * h * (w / V);
  const unsigned bx = (unsigned)std::min<long long>((total + 255) / 256, 148 * 8);
  dim3 grid(bx, 1, (unsigned)B);
  if (wide) {
    if constexpr (sizeof(T) == 2)
      conditioning_fill_kernel<T, 8><<<grid, 256, 0, st>>>((const T*)scores, (const T*)feats, (T*)out, K, C, h, w, c_total,
                                                          c_off, halves, write_scores);
  } else {
    conditioning_fill_kernel<T, 4><<<grid, 256, 0, st>>>((const T*)scores, (const T*)feats, (T*)out, K, C, h, w, c_total,
                                                        c_off, halves, write_scores);
  }
  BS_CUDA(cudaGetLastError());
  return 0;
}

int conditioning_fill_dispatch(const void* scores, const void* feats, void* out, int B, int K, int C, int h, int w,
                               int c_total, int c_off, int halves, int write_scores, int dtype, cudaStream_t st) {
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_fill<float>(scores, feats, out, B, K, C, h, w, c_total, c_off, halves, write_scores, st);
    case BLOBSPLAT_BF16: return launch_fill<__nv_bfloat16>(scores, feats, out, B, K, C, h, w, c_total, c_off, halves, write_scores, st);
    case BLOBSPLAT_F16: return launch_fill<__half>(scores, feats, out, B, K, C, h, w, c_total, c_off, halves, write_scores, st);
  }
  BS_UNSUPPORTED("conditioning fill supports float32/bfloat16/float16 (got %d)", dtype);
}

}  // namespace blobsplat

// N4: fused BlobNet residual injection.
// Replaces, per residual, the three elementwise passes of the reference loop:
//   residual * conditioning_scale                      (blobctrl/models/blobnet.py:936-938)
//   residual[..., -h:]                                 (blobctrl/pipelines/pipeline_blobnet.py:1085-1087, a slice copy)
//   sample[..., -h:] = sample[..., -h:] + residual     (diffusers/.../unet_2d_condition.py:1215-1219 and the block
//                                                       variants in unet_2d_blocks.py:1303-1319, 2598-2615)
// with one in-place pass over the right-hand `cols` columns: hidden[b,c,y,Wh-cols+x] += scale_b * residual[b,c,y,Wr-cols+x].
// Rounding follows the reference op by op (scaled residual rounded to the tensor dtype, then the sum rounded).
namespace blobsplat {

template <typename T>
__global__ void __launch_bounds__(256)
residual_inject_kernel(T* __restrict__ hidden, const T* __restrict__ residual, const float* __restrict__ scale_b, float scale,
                       long long rows_per_sample, long long rows, int Wh, int Wr, int cols) {
  const long long total = rows * cols;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / cols;
    const int x = (int)(i - row * cols);
    const float s = scale_b ? scale_b[row / rows_per_sample] : scale;
    const T r = residual[row * Wr + (Wr - cols) + x];
    T* h = hidden + row * Wh + (Wh - cols) + x;
    if constexpr (sizeof(T) == 8) {
      *h = __dadd_rn(*h, __dmul_rn(r, (double)s));
    } else {
      // explicit _rn intrinsics: no FMA contraction, the reference rounds the product before the sum
      const T scaled = Cvt<T>::from(__fmul_rn((float)Cvt<T>::to(r), s));
      *h = Cvt<T>::from(__fadd_rn((float)Cvt<T>::to(*h), (float)Cvt<T>::to(scaled)));
    }
  }
}

// Same op on 16-byte vectors (V elements): the right-half columns of a row are one contiguous, 16-byte aligned run
// whenever Wh, Wr, cols are multiples of V — every BlobNet / UNet resolution (64 .. 8 pixels wide).  Same per-element
// rounding as the scalar kernel.
template <typename T, int V>
__global__ void __launch_bounds__(256)
residual_inject_vec_kernel(T* __restrict__ hidden, const T* __restrict__ residual, const float* __restrict__ scale_b,
                           float scale, int rows_per_sample, long long rows, int Wh, int Wr, int cols) {
  const int vpr = cols / V;                                   // vectors per row
  const long long total = rows * vpr;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / vpr;
    const int x = (int)(i - row * vpr) * V;
    const float s = scale_b ? scale_b[row / rows_per_sample] : scale;
    uint4* hp = reinterpret_cast<uint4*>(hidden + row * Wh + (Wh - cols) + x);
    const uint4 rv = __ldg(reinterpret_cast<const uint4*>(residual + row * Wr + (Wr - cols) + x));
    uint4 hv = *hp;
    const T* r = reinterpret_cast<const T*>(&rv);
    T* h = reinterpret_cast<T*>(&hv);
#pragma unroll
    for (int j = 0; j < V; ++j) {
      const T scaled = Cvt<T>::from(__fmul_rn((float)Cvt<T>::to(r[j]), s));
      h[j] = Cvt<T>::from(__fadd_rn((float)Cvt<T>::to(h[j]), (float)Cvt<T>::to(scaled)));
    }
    *hp = hv;
  }
}

template <typename T>
static int launch_inject(void* hidden, const void* residual, const float* scale_b, float scale, int B, int C, int H, int Wh,
                         int Wr, int cols, cudaStream_t st) {
  const long long rows = (long long)B * C * H;
  if constexpr (sizeof(T) <= 4) {
    constexpr int V = 16 / sizeof(T);
    if (Wh % V == 0 && Wr % V == 0 && cols % V == 0 && (long long)C * H < (1ll << 31) &&
        ((reinterpret_cast<uintptr_t>(hidden) | reinterpret_cast<uintptr_t>(residual)) & 15) == 0) {
      const long long vecs = rows * (cols / V);
      const unsigned vb = (unsigned)std::min<long long>((vecs + 255) / 256, 148 * 32);
      residual_inject_vec_kernel<T, V><<<vb, 256, 0, st>>>((T*)hidden, (const T*)residual, scale_b, scale, C * H, rows, Wh, Wr, cols);
      BS_CUDA(cudaGetLastError());
      return 0;
    }
  }
  const long long total = rows * cols;
  const unsigned blocks = (unsigned)std::min<long long>((total + 255) / 256, 148 * 16);
  residual_inject_kernel<T><<<blocks, 256, 0, st>>>((T*)hidden, (const T*)residual, scale_b, scale, (long long)C * H, rows, Wh,
                                                   Wr, cols);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int residual_inject_dispatch(void* hidden, const void* residual, const float* scale_b, float scale, int B, int C, int H, int Wh,
                             int Wr, int cols, int dtype, cudaStream_t st) {
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_inject<float>(hidden, residual, scale_b, scale, B, C, H, Wh, Wr, cols, st);
    case BLOBSPLAT_F64: return launch_inject<double>(hidden, residual, scale_b, scale, B, C, H, Wh, Wr, cols, st);
    case BLOBSPLAT_BF16: return launch_inject<__nv_bfloat16>(hidden, residual, scale_b, scale, B, C, H, Wh, Wr, cols, st);
    case BLOBSPLAT_F16: return launch_inject<__half>(hidden, residual, scale_b, scale, B, C, H, Wh, Wr, cols, st);
  }
  BS_UNSUPPORTED("unknown dtype %d", dtype);
}

}  // namespace blobsplat

namespace blobsplat {

// N1: loop-invariant hoist of BlobNet's conv_in.
//
// BlobNet's first layer is Conv2d(lc + 1 + C, O, 3, padding = 1) (blobctrl/models/blobnet.py:241-245, applied at :840) over
// the canvas construct_blobnet_input builds every step (blobctrl/pipelines/pipeline_blobnet.py:724-739, :1043-1049):
// lc = 4 latent planes, one score plane, C = 1024 feature planes, left half = reference-image latents, right half =
// noisy latents, the conditioning planes identical in both halves.  The feature planes are rank-K
// (feats[c] = sum_k s_k * f[k, c], :984) and convolution is linear, so
//
//   conv_in(canvas)[b, o] = bias[o] + sum_{c < lc} W[o, c] * lat[b, c]                       <- changes every step
//                         + W[o, lc] * score[b]  +  sum_k W_eff[b, o, k] * s_k[b]             <- loop-invariant planes
//   W_eff[b, o, k, tap] = sum_c W[o, lc + 1 + c, tap] * f[b, k, c]                            <- once per edit
//
// i.e. a (lc + 1 + K) -> O convolution with per-sample kernels for the conditioning planes instead of a 1029 -> 320 one
// (~48.6 GFLOP per sample and step), and neither the 1024 feature planes nor the 1029-plane canvas are ever built.
//
// conv_in_weights_kernel: one block per (sample, output channel): the 9 * K dot products over C.
template <typename T, int KMAX>
__global__ void __launch_bounds__(256)
conv_in_weights_kernel(const T* __restrict__ weight, const T* __restrict__ feats, float* __restrict__ weff, int O, int Cin,
                       int lc, int C, int K, int Jp) {
  // weff [B, O, Jp, 12]: plane 0 = the score plane's own kernel W[o, lc]; plane 1 + k = W_eff[b, o, k]; 9 taps padded to 12
  const int o = blockIdx.x, b = blockIdx.y;
  const T* w = weight + ((size_t)o * Cin + lc + 1) * 9;       // [C][9] contiguous
  const T* f = feats + (size_t)b * K * C;
  __shared__ float red[8][KMAX * 9];
  for (int k0 = 0; k0 < K; k0 += KMAX) {
    const int kn = min(KMAX, K - k0);
    float acc[KMAX][9];
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) acc[k][t] = 0.f;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      float wv[9];
#pragma unroll
      for (int t = 0; t < 9; ++t) wv[t] = (float)Cvt<T>::to(w[(size_t)c * 9 + t]);
#pragma unroll
      for (int k = 0; k < KMAX; ++k) {
        if (k < kn) {
          const float fv = (float)Cvt<T>::to(f[(size_t)(k0 + k) * C + c]);
#pragma unroll
          for (int t = 0; t < 9; ++t) acc[k][t] = fmaf(wv[t], fv, acc[k][t]);
        }
      }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 0; k < KMAX; ++k)
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        float v = acc[k][t];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (lane == 0) red[warp][k * 9 + t] = v;
      }
    __syncthreads();
    if ((int)threadIdx.x < kn * 9) {
      float v = 0.f;
      for (int wi = 0; wi < (int)(blockDim.x >> 5); ++wi) v += red[wi][threadIdx.x];
      const int k = threadIdx.x / 9, t = threadIdx.x - k * 9;
      weff[(((size_t)b * O + o) * Jp + 1 + k0 + k) * 12 + t] = v;
    }
    __syncthreads();
  }
  if (threadIdx.x < 9) weff[(((size_t)b * O + o) * Jp) * 12 + threadIdx.x] = (float)Cvt<T>::to(weight[((size_t)o * Cin + lc) * 9 + threadIdx.x]);
}

// conv_in_hoisted_kernel: the per-step layer.  Block = 64 threads x 2 pixels (x and x + 64 of a 128-pixel row segment)
// x kOcT output channels; the (lc + J) input planes' 3-row halo and the block's kernels live in shared memory, the kernels
// padded to 12 floats so a thread fetches one (output channel, plane) kernel with three broadcast LDS.128 for its 18 FMAs.
constexpr int kOcT = 32;
template <typename T>
__global__ void __launch_bounds__(64)
conv_in_hoisted_kernel(const T* __restrict__ lat, const T* __restrict__ cond, const T* __restrict__ weight,
                       const T* __restrict__ bias, const float* __restrict__ weff, T* __restrict__ out, int O, int Cin, int lc,
                       int J, int h, int w, int Wt) {
  extern __shared__ __align__(16) float sm[];
  const int planes = lc + J;
  float* wsm = sm;                                   // [kOcT][planes][12]
  float* in = sm + (size_t)kOcT * planes * 12;       // [planes][3][130]
  const int xt = blockIdx.x % ((Wt + 127) / 128), y = blockIdx.x / ((Wt + 127) / 128);
  const int o0 = blockIdx.y * kOcT, b = blockIdx.z;
  const int x0 = xt * 128;
  for (int i = threadIdx.x; i < kOcT * planes * 12; i += 64) {
    const int t = i % 12, pl = (i / 12) % planes, oc = i / (12 * planes);
    float v = 0.f;
    if (t < 9 && o0 + oc < O)
      v = pl < lc ? (float)Cvt<T>::to(weight[((size_t)(o0 + oc) * Cin + pl) * 9 + t])
                  : weff[(((size_t)b * O + o0 + oc) * J + (pl - lc)) * 12 + t];
    wsm[i] = v;
  }
  for (int i = threadIdx.x; i < planes * 3 * 130; i += 64) {
    const int xx = i % 130, r = (i / 130) % 3, pl = i / 390;
    const int gx = x0 + xx - 1, gy = y + r - 1;
    float v = 0.f;
    if (gx >= 0 && gx < Wt && gy >= 0 && gy < h)
      v = pl < lc ? (float)Cvt<T>::to(lat[(((size_t)b * lc + pl) * h + gy) * Wt + gx])
                  : (float)Cvt<T>::to(cond[(((size_t)b * J + (pl - lc)) * h + gy) * w + (gx % w)]);   // same planes in every half
    in[i] = v;
  }
  __syncthreads();
  float acc0[kOcT], acc1[kOcT];
#pragma unroll
  for (int oc = 0; oc < kOcT; ++oc) {
    const float bv = (bias && o0 + oc < O) ? (float)Cvt<T>::to(bias[o0 + oc]) : 0.f;
    acc0[oc] = bv; acc1[oc] = bv;
  }
  const int tx = threadIdx.x;
  for (int pl = 0; pl < planes; ++pl) {
    float a[9], c[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        a[r * 3 + d] = in[(pl * 3 + r) * 130 + tx + d];
        c[r * 3 + d] = in[(pl * 3 + r) * 130 + tx + 64 + d];
      }
#pragma unroll
    for (int oc = 0; oc < kOcT; ++oc) {
      const float4* wp = reinterpret_cast<const float4*>(wsm + ((size_t)oc * planes + pl) * 12);
      const float4 w0 = wp[0], w1 = wp[1], w2 = wp[2];
      const float wv[9] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w, w2.x};
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        acc0[oc] = fmaf(wv[t], a[t], acc0[oc]);
        acc1[oc] = fmaf(wv[t], c[t], acc1[oc]);
      }
    }
  }
  const int gx0 = x0 + tx, gx1 = x0 + tx + 64;
#pragma unroll
  for (int oc = 0; oc < kOcT; ++oc) {
    if (o0 + oc >= O) break;
    T* orow = out + (((size_t)b * O + o0 + oc) * h + y) * Wt;
    if (gx0 < Wt) orow[gx0] = Cvt<T>::from(acc0[oc]);
    if (gx1 < Wt) orow[gx1] = Cvt<T>::from(acc1[oc]);
  }
}

template <typename T>
static int launch_conv_in_weights(const void* weight, const void* feats, float* weff, int B, int O, int Cin, int lc, int C, int K,
                                  cudaStream_t st) {
  conv_in_weights_kernel<T, 4><<<dim3((unsigned)O, (unsigned)B), 256, 0, st>>>((const T*)weight, (const T*)feats, weff, O, Cin, lc,
                                                                              C, K, 1 + K);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int conv_in_weights_dispatch(const void* weight, const void* feats, float* weff, int B, int O, int Cin, int lc, int C, int K,
                             int dtype, cudaStream_t st) {
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_conv_in_weights<float>(weight, feats, weff, B, O, Cin, lc, C, K, st);
    case BLOBSPLAT_BF16: return launch_conv_in_weights<__nv_bfloat16>(weight, feats, weff, B, O, Cin, lc, C, K, st);
    case BLOBSPLAT_F16: return launch_conv_in_weights<__half>(weight, feats, weff, B, O, Cin, lc, C, K, st);
  }
  BS_UNSUPPORTED("conv_in hoist supports float32/bfloat16/float16 (got %d)", dtype);
}

template <typename T>
static int launch_conv_in_hoisted(const void* lat, const void* cond, const void* weight, const void* bias, const float* weff,
                                  void* out, int B, int O, int Cin, int lc, int J, int h, int w, int Wt, cudaStream_t st) {
  const int planes = lc + J;
  const size_t smem = ((size_t)kOcT * planes * 12 + (size_t)planes * 3 * 130) * sizeof(float);
  if (smem > 200 * 1024) BS_UNSUPPORTED("conv_in hoist: %d input planes do not fit in shared memory", planes);
  static thread_local int configured_dev = -1;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(conv_in_hoisted_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured_dev = dev;
  }
  dim3 grid((unsigned)(((Wt + 127) / 128) * h), (unsigned)((O + kOcT - 1) / kOcT), (unsigned)B);
  conv_in_hoisted_kernel<T><<<grid, 64, smem, st>>>((const T*)lat, (const T*)cond, (const T*)weight, (const T*)bias, weff, (T*)out,
                                                   O, Cin, lc, J, h, w, Wt);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int conv_in_hoisted_dispatch(const void* lat, const void* cond, const void* weight, const void* bias, const float* weff, void* out,
                             int B, int O, int Cin, int lc, int J, int h, int w, int Wt, int dtype, cudaStream_t st) {
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_conv_in_hoisted<float>(lat, cond, weight, bias, weff, out, B, O, Cin, lc, J, h, w, Wt, st);
    case BLOBSPLAT_BF16: return launch_conv_in_hoisted<__nv_bfloat16>(lat, cond, weight, bias, weff, out, B, O, Cin, lc, J, h, w, Wt, st);
    case BLOBSPLAT_F16: return launch_conv_in_hoisted<__half>(lat, cond, weight, bias, weff, out, B, O, Cin, lc, J, h, w, Wt, st);
  }
  BS_UNSUPPORTED("conv_in hoist supports float32/bfloat16/float16 (got %d)", dtype);
}

}  // namespace blobsplat
