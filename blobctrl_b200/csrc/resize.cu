// Bilinear resize (align_corners=False) and the halving pyramid.
//
// Replaces torch.nn.functional.interpolate(mode='bilinear', align_corners=False) at the two places
// the renderer uses it: pyramid_resize (blobctrl/utils/utils.py:280-294) and the size != H branch of
// splat_features_from_scores (utils.py:70-73).  Source-index rule restated from ATen's
// area_pixel_compute_source_index: src = max(0, (dst + 0.5) * in/out - 0.5), i1 = min(i0 + 1, in-1).
// For an exact halving of an even size this is the 2x2 mean (weights 0.5), so the whole pyramid is
// produced by one launch that keeps each 2^L x 2^L block in registers.
#include "common.cuh"

namespace blobsplat {

template <typename T, typename A>
__global__ void __launch_bounds__(256)
resize_bilinear_kernel(const T* __restrict__ in, T* __restrict__ out, int B, int Hin, int Win, int Hout, int Wout,
                       A scale_h, A scale_w) {
  const size_t total = (size_t)B * Hout * Wout;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ox = (int)(i % Wout);
    const int oy = (int)((i / Wout) % Hout);
    const size_t b = i / ((size_t)Wout * Hout);
    A sy = ((A)oy + (A)0.5) * scale_h - (A)0.5; sy = sy < (A)0 ? (A)0 : sy;
    A sx = ((A)ox + (A)0.5) * scale_w - (A)0.5; sx = sx < (A)0 ? (A)0 : sx;
    int y0 = (int)sy; y0 = y0 > Hin - 1 ? Hin - 1 : y0;
    int x0 = (int)sx; x0 = x0 > Win - 1 ? Win - 1 : x0;
    const int y1 = y0 + (y0 < Hin - 1), x1 = x0 + (x0 < Win - 1);
    const A ly = sy - (A)y0, lx = sx - (A)x0;
    const A hy = (A)1 - ly, hx = (A)1 - lx;
    const T* src = in + b * (size_t)Hin * Win;
    const A v00 = (A)Cvt<T>::to(src[(size_t)y0 * Win + x0]), v01 = (A)Cvt<T>::to(src[(size_t)y0 * Win + x1]);
    const A v10 = (A)Cvt<T>::to(src[(size_t)y1 * Win + x0]), v11 = (A)Cvt<T>::to(src[(size_t)y1 * Win + x1]);
    // same association as ATen: h0*(w0*a + w1*b) + h1*(w0*c + w1*d)
    out[i] = Cvt<T>::from(hy * (hx * v00 + lx * v01) + ly * (hx * v10 + lx * v11));
  }
}

// One thread owns a 2^L x 2^L block of the finest level (L = n_levels <= 3) and emits every coarser level
// from registers.  Level l value = 0.5*(0.5*a + 0.5*b) + 0.5*(0.5*c + 0.5*d) of level l-1, rounded to T
// between levels exactly as the reference's repeated interpolate does.
constexpr int kMaxPyramidLevels = 3;
struct PyramidPtrs { void* p[kMaxPyramidLevels]; };

template <typename T, typename A, int L>
__global__ void __launch_bounds__(256)
pyramid_kernel(const T* __restrict__ in, PyramidPtrs outs, int B, int S) {
  constexpr int E = 1 << L;  // block edge at the finest level
  const int bs = S / E;      // blocks per row
  const size_t total = (size_t)B * bs * bs;
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  if (i >= total) return;
  const int bx = (int)(i % bs), by = (int)((i / bs) % bs);
  const size_t b = i / ((size_t)bs * bs);
  A v[E][E];
  const T* src = in + b * (size_t)S * S + (size_t)by * E * S + (size_t)bx * E;
  // a block row is E contiguous elements: 16-byte loads where that is a whole number of vectors and aligned
  constexpr int VL = 16 / sizeof(T);
  const bool vec = (E % VL == 0) && ((reinterpret_cast<uintptr_t>(in) & 15) == 0) && (S % VL == 0);
  if (vec) {
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
      for (int c0 = 0; c0 < E; c0 += VL) {
        if constexpr (E % VL == 0) {
          const uint4 raw = __ldg(reinterpret_cast<const uint4*>(src + (size_t)r * S + c0));
          const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
          for (int j = 0; j < VL; ++j) v[r][c0 + j] = (A)Cvt<T>::to(e[j]);
        }
      }
  } else {
#pragma unroll
    for (int r = 0; r < E; ++r)
#pragma unroll
      for (int c = 0; c < E; ++c) v[r][c] = (A)Cvt<T>::to(src[(size_t)r * S + c]);
  }
#pragma unroll
  for (int l = 1; l <= L; ++l) {
    const int e = E >> l;   // block edge at this level
    const int s = S >> l;   // image size at this level
    T* dst = (T*)outs.p[l - 1] + b * (size_t)s * s + (size_t)by * e * s + (size_t)bx * e;
#pragma unroll
    for (int r = 0; r < e; ++r) {
      T q[E];
#pragma unroll
      for (int c = 0; c < e; ++c) {
        const A top = (A)0.5 * v[2 * r][2 * c] + (A)0.5 * v[2 * r][2 * c + 1];
        const A bot = (A)0.5 * v[2 * r + 1][2 * c] + (A)0.5 * v[2 * r + 1][2 * c + 1];
        q[c] = Cvt<T>::from((A)0.5 * top + (A)0.5 * bot);
        v[r][c] = (A)Cvt<T>::to(q[c]);
      }
      // one store per block row where its e elements make an aligned 4 / 8 / 16-byte word
      T* drow = dst + (size_t)r * s;
      const int bytes = e * (int)sizeof(T);
      const bool al = (reinterpret_cast<uintptr_t>(outs.p[l - 1]) & 15) == 0 && (s * (int)sizeof(T)) % bytes == 0;
      if (al && bytes == 16) *reinterpret_cast<uint4*>(drow) = *reinterpret_cast<const uint4*>(q);
      else if (al && bytes == 8) *reinterpret_cast<uint2*>(drow) = *reinterpret_cast<const uint2*>(q);
      else if (al && bytes == 4) *reinterpret_cast<uint32_t*>(drow) = *reinterpret_cast<const uint32_t*>(q);
      else {
#pragma unroll
        for (int c = 0; c < e; ++c) drow[c] = q[c];
      }
    }
  }
}

template <typename T, typename A>
static int launch_resize(const void* in, void* out, int B, int Hin, int Win, int Hout, int Wout, cudaStream_t st) {
  const size_t total = (size_t)B * Hout * Wout;
  const unsigned blocks = (unsigned)((total + 255) / 256 < 148u * 64u ? (total + 255) / 256 : 148u * 64u);
  resize_bilinear_kernel<T, A><<<blocks, 256, 0, st>>>((const T*)in, (T*)out, B, Hin, Win, Hout, Wout,
                                                       (A)Hin / (A)Hout, (A)Win / (A)Wout);
  BS_CUDA(cudaGetLastError());
  return 0;
}

int resize_dispatch(const void* in, void* out, int B, int Hin, int Win, int Hout, int Wout, int dtype,
                    cudaStream_t st) {
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_resize<float, float>(in, out, B, Hin, Win, Hout, Wout, st);
    case BLOBSPLAT_F64: return launch_resize<double, double>(in, out, B, Hin, Win, Hout, Wout, st);
    case BLOBSPLAT_BF16: return launch_resize<__nv_bfloat16, float>(in, out, B, Hin, Win, Hout, Wout, st);
    case BLOBSPLAT_F16: return launch_resize<__half, float>(in, out, B, Hin, Win, Hout, Wout, st);
  }
  BS_UNSUPPORTED("unknown dtype %d", dtype);
}

template <typename T, typename A>
static int launch_pyramid(const void* in, void* const* outs, int n_levels, int B, int S, cudaStream_t st) {
  PyramidPtrs p{};
  for (int l = 0; l < n_levels; ++l) p.p[l] = outs[l];
  const size_t total = (size_t)B * (S >> n_levels) * (S >> n_levels);
  const unsigned blocks = (unsigned)((total + 255) / 256);
  switch (n_levels) {
    case 1: BS_CUDA(launch_pdl(pyramid_kernel<T, A, 1>, dim3(blocks), dim3(256), 0, st, (const T*)in, p, B, S)); break;
    case 2: BS_CUDA(launch_pdl(pyramid_kernel<T, A, 2>, dim3(blocks), dim3(256), 0, st, (const T*)in, p, B, S)); break;
    case 3: BS_CUDA(launch_pdl(pyramid_kernel<T, A, 3>, dim3(blocks), dim3(256), 0, st, (const T*)in, p, B, S)); break;
  }
  return 0;
}

int pyramid_dispatch(const void* in, void* const* outs, int n_levels, int B, int S, int dtype, cudaStream_t st) {
  if (n_levels < 1 || n_levels > kMaxPyramidLevels)
    BS_UNSUPPORTED("pyramid supports 1..%d levels per launch (got %d); chain calls for more", kMaxPyramidLevels, n_levels);
  switch (dtype) {
    case BLOBSPLAT_F32: return launch_pyramid<float, float>(in, outs, n_levels, B, S, st);
    case BLOBSPLAT_F64: return launch_pyramid<double, double>(in, outs, n_levels, B, S, st);
    case BLOBSPLAT_BF16: return launch_pyramid<__nv_bfloat16, float>(in, outs, n_levels, B, S, st);
    case BLOBSPLAT_F16: return launch_pyramid<__half, float>(in, outs, n_levels, B, S, st);
  }
  BS_UNSUPPORTED("unknown dtype %d", dtype);
}

}  // namespace blobsplat
