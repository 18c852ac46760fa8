// Shared device/host helpers for the blob-splat kernels (sm_100a only).
#pragma once
#include <utility>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <type_traits>

#include "../../include/blobsplat.h"

namespace blobsplat {

// ---- error plumbing (thread-local message, C-ABI status codes) ----------------------------------
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);

#define BS_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      ::blobsplat::set_error(__VA_ARGS__); \
      return BLOBSPLAT_E_INVALID;          \
    }                                      \
  } while (0)

#define BS_UNSUPPORTED(...)              \
  do {                                   \
    ::blobsplat::set_error(__VA_ARGS__); \
    return BLOBSPLAT_E_UNSUPPORTED;      \
  } while (0)

#define BS_CUDA(call)                                                 \
  do {                                                                \
    cudaError_t _e = (call);                                          \
    if (_e != cudaSuccess) return ::blobsplat::cuda_fail(_e, #call);  \
  } while (0)

// RAII device guard for the `device` argument of the C ABI.
struct DeviceGuard {
  int prev = -1;
  bool switched = false;
  int status = 0;
  explicit DeviceGuard(int device) {
    if (device < 0) return;
    cudaError_t e = cudaGetDevice(&prev);
    if (e != cudaSuccess) { status = cuda_fail(e, "cudaGetDevice"); return; }
    if (prev != device) {
      e = cudaSetDevice(device);
      if (e != cudaSuccess) { status = cuda_fail(e, "cudaSetDevice"); return; }
      switched = true;
    }
  }
  ~DeviceGuard() {
    if (switched) cudaSetDevice(prev);
  }
};

inline size_t dtype_size(int dt) {
  switch (dt) {
    case BLOBSPLAT_F32: return 4;
    case BLOBSPLAT_F64: return 8;
    case BLOBSPLAT_BF16: return 2;
    case BLOBSPLAT_F16: return 2;
    default: return 0;
  }
}

inline bool aligned_to(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

// ---- scalar conversion -----------------------------------------------------------------------------
template <typename T> struct Cvt;
template <> struct Cvt<float> {
  __device__ __forceinline__ static float from(float v) { return v; }
  __device__ __forceinline__ static float to(float v) { return v; }
};
template <> struct Cvt<double> {
  __device__ __forceinline__ static double from(double v) { return v; }
  __device__ __forceinline__ static double to(double v) { return v; }
};
template <> struct Cvt<__nv_bfloat16> {
  __device__ __forceinline__ static __nv_bfloat16 from(float v) { return __float2bfloat16_rn(v); }
  __device__ __forceinline__ static float to(__nv_bfloat16 v) { return __bfloat162float(v); }
};
template <> struct Cvt<__half> {
  __device__ __forceinline__ static __half from(float v) { return __float2half_rn(v); }
  __device__ __forceinline__ static float to(__half v) { return __half2float(v); }
};

// ---- 128-bit (or narrower) vector stores of V consecutive elements, streaming (evict-first) --------
// Outputs are written once and never re-read by the same kernel: st.global.cs keeps them from
// displacing the small, hot inputs (blob table, features) in L1/L2.
template <typename T, int V> struct VecStore;

template <> struct VecStore<float, 4> {
  __device__ __forceinline__ static void st(float* p, const float* v) {
    __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
  }
};
template <> struct VecStore<float, 1> {
  __device__ __forceinline__ static void st(float* p, const float* v) { __stcs(p, v[0]); }
};
template <> struct VecStore<double, 2> {
  __device__ __forceinline__ static void st(double* p, const double* v) {
    __stcs(reinterpret_cast<double2*>(p), make_double2(v[0], v[1]));
  }
};
template <> struct VecStore<double, 1> {
  __device__ __forceinline__ static void st(double* p, const double* v) { __stcs(p, v[0]); }
};
template <> struct VecStore<__nv_bfloat16, 8> {
  __device__ __forceinline__ static void st(__nv_bfloat16* p, const float* v) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]), b = __floats2bfloat162_rn(v[2], v[3]);
    __nv_bfloat162 c = __floats2bfloat162_rn(v[4], v[5]), d = __floats2bfloat162_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    __stcs(reinterpret_cast<uint4*>(p), u);
  }
};
template <> struct VecStore<__nv_bfloat16, 1> {
  __device__ __forceinline__ static void st(__nv_bfloat16* p, const float* v) { *p = __float2bfloat16_rn(v[0]); }
};
template <> struct VecStore<__half, 8> {
  __device__ __forceinline__ static void st(__half* p, const float* v) {
    __half2 a = __floats2half2_rn(v[0], v[1]), b = __floats2half2_rn(v[2], v[3]);
    __half2 c = __floats2half2_rn(v[4], v[5]), d = __floats2half2_rn(v[6], v[7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&a); u.y = *reinterpret_cast<uint32_t*>(&b);
    u.z = *reinterpret_cast<uint32_t*>(&c); u.w = *reinterpret_cast<uint32_t*>(&d);
    __stcs(reinterpret_cast<uint4*>(p), u);
  }
};
template <> struct VecStore<__half, 1> {
  __device__ __forceinline__ static void st(__half* p, const float* v) { *p = __float2half_rn(v[0]); }
};

// number of elements in one 128-bit store
template <typename T> struct Vec128 { static constexpr int n = 16 / sizeof(T); };

// ---- programmatic dependent launch (PDL) --------------------------------------------------------------------
// Kernels launched with launch_pdl may begin (block scheduling, TMEM / barrier set-up) while the previous kernel in
// the stream is still draining; they call pdl_wait() before their first global-memory access, which blocks until
// every prerequisite grid has completed and its writes are visible.  pdl_launch_dependents() lets the NEXT kernel's
// blocks be scheduled as soon as this grid's blocks start to retire.  Saves the launch gap between the back-to-back
// launches of a multi-resolution render (BS_PDL=0 turns it off).
#ifndef BS_PDL
#define BS_PDL 1
#endif
__device__ __forceinline__ void pdl_wait() {
#if BS_PDL
  asm volatile("griddepcontrol.wait;" ::: "memory");
#endif
}
__device__ __forceinline__ void pdl_launch_dependents() {
#if BS_PDL
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
#endif
}
template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = BS_PDL;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

constexpr float kLog2e = 1.4426950408889634f;

// ---- per-blob coefficients, built once per (image, blob) in fp64 and staged in shared memory -------
// Whitened form: q*log2(e) = u^2 + v^2, u = p*dx, v = r*dx + t*dy, dx/dy in pixels
// relative to the centre.  The centre is a hi+lo float pair so (x - cx) carries no fp32 cancellation error; the
// lo parts are folded into per-blob constants: with dxh = x - cx_hi, dyh = y - cy_hi (exact-ish differences)
//   u = fma(p, dxh, u0),  v = fma(r, dxh, fma(t, dyh, v0)),  u0 = -p*cx_lo,  v0 = -(r*cx_lo + t*cy_lo)
// — 2 FADD + 3 FFMA per pixel and blob.  Non-positive-definite covariances (never produced by the reference's
// callers) take the general quadratic form qa*dx^2 + qb*dx*dy + qc*dy^2 instead, with (u0, v0) = (cx_lo, cy_lo).
struct __align__(16) BlobCoef {
  float cx_hi, cy_hi;
  float p, r, t;      // whitened (or qa, qb, qc when c0 == kC0General)
  float u0, v0;       // folded centre residuals (or cx_lo, cy_lo when c0 == kC0General)
  float c0;           // exponent offset AND blob kind: see below
};
// c0 is the constant folded into the exponent, q' - 1 = u^2 + v^2 + c0, and doubles as the blob's kind:
//   kC0Normal  (-1)      positive-definite blob
//   kC0Gated   (> 0)     non-existent blob (sizes < 0.5 -> score 1e-6, utils.py:165-172): p = r = t = 0 and
//                        c0 = log2(1e6 - 0.5), so the branch-free form yields 1/(0.5 + 2^c0) = 1e-6 (to ~3e-7 relative)
//                        with no select; the branching forms test c0 > 0 and return the exact constant
//   kC0General (-2)      covariance not positive definite: p, r, t hold the plain quadratic form
constexpr float kC0Normal = -1.0f, kC0General = -2.0f, kC0Gated = 19.931567847f;
__device__ __forceinline__ bool coef_gated(const BlobCoef& c) { return c.c0 > 0.0f; }
__device__ __forceinline__ bool coef_general(const BlobCoef& c) { return c.c0 == kC0General; }

__device__ __forceinline__ BlobCoef make_blob_coef(double xs, double ys, double c00, double c01, double c10,
                                                   double c11, float size, int H, int W) {
  BlobCoef o;
  // centre in pixels: utils.py:138 (square) / :147 (tuple) — xs*W, ys*H
  const double cx = xs * (double)W, cy = ys * (double)H;
  o.cx_hi = (float)cx; o.cy_hi = (float)cy;
  const double cx_lo = cx - (double)o.cx_hi, cy_lo = cy - (double)o.cy_hi;
  // q = delta^T Sigma^-1 delta with delta = (dx/W, dy/H): only the symmetric part of Sigma^-1 matters.
  const double det = c00 * c11 - c01 * c10;
  const double l2e = 1.4426950408889634;
  const double A = l2e * (c11 / det) / ((double)W * (double)W);
  const double B = l2e * (-0.5 * (c01 + c10) / det) / ((double)W * (double)H);
  const double C = l2e * (c00 / det) / ((double)H * (double)H);
  const double schur = A - (B * B) / C;
  if (size < 0.5f) {
    o.p = o.r = o.t = o.u0 = o.v0 = 0.0f; o.c0 = kC0Gated;
  } else if (C > 0.0 && schur > 0.0) {
    const double t = sqrt(C);
    o.t = (float)t; o.r = (float)(B / t); o.p = (float)sqrt(schur); o.c0 = kC0Normal;
    o.u0 = (float)(-(double)o.p * cx_lo); o.v0 = (float)(-((double)o.r * cx_lo + (double)o.t * cy_lo));
  } else {
    o.p = (float)A; o.r = (float)(2.0 * B); o.t = (float)C; o.c0 = kC0General;
    o.u0 = (float)cx_lo; o.v0 = (float)cy_lo;
  }
  return o;
}

// Same coefficients straight from an OpenCV ellipse ((xc,yc),(d1,d2),angle_deg) in pixels of an img_w x img_h image
// — the reference's host recipe (scripts/blobctrl_inference.py:71-109 + utils.py:297-341) folded into the per-blob
// prologue: theta = rad(((180-angle) mod 180 + 90) mod 180), Sigma = Q diag(b^2, a^2) Q^T / (img_w^2 + img_h^2) with
// Q = [[c, s], [-s, c]] (the rotation with the off-diagonals negated), a = d1/2, b = d2/2.  Then
// q = D^2 * ((c*dx' - s*dy')^2 / b^2 + (s*dx' + c*dy')^2 / a^2), dx' = dx/W, dy' = dy/H: no covariance, no inverse, no
// cancellation — exact whitening factors even for the degenerate 1e-5-pixel ellipses of the demo states.
__device__ __forceinline__ BlobCoef make_blob_coef_ellipse(double xc, double yc, double d1, double d2, double angle_deg,
                                                           float size, double img_w, double img_h, int H, int W) {
  BlobCoef o;
  const double cx = xc / img_w * (double)W, cy = yc / img_h * (double)H;
  o.cx_hi = (float)cx; o.cy_hi = (float)cy;
  const double cx_lo = cx - (double)o.cx_hi, cy_lo = cy - (double)o.cy_hi;
  double a1 = 180.0 - angle_deg; a1 -= 180.0 * floor(a1 / 180.0);            // python's % 180
  double a2 = a1 + 90.0; a2 -= 180.0 * floor(a2 / 180.0);
  const double th = a2 * 0.017453292519943295;
  const double cs = cos(th), sn = sin(th);
  const double a = 0.5 * d1, b = 0.5 * d2;
  const double D = sqrt((img_w * img_w + img_h * img_h) * 1.4426950408889634);  // diagonal, log2(e) folded in
  const double m00 = D * cs / (b * (double)W), m01 = -D * sn / (b * (double)H);
  const double m10 = D * sn / (a * (double)W), m11 = D * cs / (a * (double)H);
  const double Bq = m00 * m01 + m10 * m11, Cq = m01 * m01 + m11 * m11;
  const double t = sqrt(Cq);
  o.t = (float)t; o.r = (float)(Bq / t); o.p = (float)(fabs(m00 * m11 - m01 * m10) / t); o.c0 = kC0Normal;
  o.u0 = (float)(-(double)o.p * cx_lo); o.v0 = (float)(-((double)o.r * cx_lo + (double)o.t * cy_lo));
  if (size < 0.5f) { o.p = o.r = o.t = o.u0 = o.v0 = 0.0f; o.c0 = kC0Gated; }
  return o;
}

// s = min(1, 2*sigmoid(-q)) = min(1, 2 / (1 + 2^(q*log2e)))   (utils.py:162-163)
__device__ __forceinline__ float opacity_from_q2(float q2) {
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q2));
  return fminf(__fdividef(2.0f, 1.0f + e), 1.0f);
}

// same with the "-1" already folded into the exponent: s = 1 / (0.5 + 2^(q' - 1)) — one MUFU.EX2, one FADD,
// one MUFU.RCP (no multiply by 2)
__device__ __forceinline__ float opacity_from_q2m1(float q2m1) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q2m1));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(0.5f + e));
  return fminf(r, 1.0f);
}

// branch-free form for positive-definite blobs (the only kind the reference's callers produce)
__device__ __forceinline__ float blob_opacity_pd(const BlobCoef& c, float xf, float yf) {
  const float dyh = yf - c.cy_hi, dxh = xf - c.cx_hi;
  const float u = fmaf(c.p, dxh, c.u0);
  const float v = fmaf(c.r, dxh, fmaf(c.t, dyh, c.v0));
  return opacity_from_q2m1(fmaf(u, u, fmaf(v, v, c.c0)));   // gated blobs: u = v = 0, c0 = log2(1e6 - 0.5)
}

// raw opacity of one blob at one pixel (stage 1 + gate)
__device__ __forceinline__ float blob_opacity(const BlobCoef& c, float xf, float yf) {
  if (coef_gated(c)) return 1e-6f;
  const float dyh = yf - c.cy_hi, dxh = xf - c.cx_hi;
  if (!coef_general(c)) {
    const float u = fmaf(c.p, dxh, c.u0);
    const float v = fmaf(c.r, dxh, fmaf(c.t, dyh, c.v0));
    return opacity_from_q2(fmaf(u, u, v * v));
  }
  const float dx = dxh - c.u0, dy = dyh - c.v0;          // general form: (u0, v0) = (cx_lo, cy_lo)
  return opacity_from_q2(fmaf(dx, fmaf(c.p, dx, c.r * dy), c.t * dy * dy));
}

}  // namespace blobsplat
