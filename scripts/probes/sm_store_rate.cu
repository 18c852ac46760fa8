// Probe: what can ONE SM push into L2 when DRAM is not the limit?  A few CTAs (one per SM) rewrite a small, L2-resident
// region with 4 / 8 / 16-byte stores per lane, from 4..16 warps.  Tells whether the ~20-24 B/clk/SM the drains reach is a
// property of the SM's store path or of how the drains issue.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o sm_store_rate sm_store_rate.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int W>   // bytes per lane
__global__ void __launch_bounds__(1024) hammer(unsigned char* buf, size_t region, int iters, int plane_stride) {
  unsigned char* base = buf + (size_t)blockIdx.x * region;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  // warp w writes "planes" like the drains: consecutive store instructions of a warp go to addresses plane_stride apart
  for (int it = 0; it < iters; ++it) {
    for (size_t off = (size_t)warp * 32 * W; off + 32 * W <= region; off += (size_t)warps * 32 * W) {
      size_t o = plane_stride ? ((off / (32 * W)) * (size_t)plane_stride + (it * 32 * W) % plane_stride) % (region - 32 * W) : off;
      o = o / (32 * W) * (32 * W);
      unsigned char* p = base + o + lane * W;
      if (W == 4) __stcs(reinterpret_cast<unsigned int*>(p), 1u);
      else if (W == 8) __stcs(reinterpret_cast<uint2*>(p), make_uint2(1u, 2u));
      else __stcs(reinterpret_cast<uint4*>(p), make_uint4(1u, 2u, 3u, 4u));
    }
  }
}

int main() {
  const size_t region = 256 << 10;      // per CTA: 256 KB; 16 CTAs -> 4 MB, far inside L2
  int clk_khz; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int ctas : {1, 16, 148}) {
    unsigned char* buf; CK(cudaMalloc(&buf, region * ctas));
    for (int W : {4, 8, 16}) for (int threads : {128, 256, 512, 1024}) for (int ps : {0, 8192}) {
      const int iters = 200;
      auto launch = [&] {
        if (W == 4) hammer<4><<<ctas, threads>>>(buf, region, iters, ps);
        else if (W == 8) hammer<8><<<ctas, threads>>>(buf, region, iters, ps);
        else hammer<16><<<ctas, threads>>>(buf, region, iters, ps);
      };
      launch(); CK(cudaDeviceSynchronize());
      CK(cudaEventRecord(a)); launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b));
      const double bytes = (double)region * iters;      // per CTA
      printf("CTAs %3d  %2d B/lane  %4d threads  %s: %6.1f GB/s per SM = %5.1f B/clk @%.2f GHz   (%7.1f GB/s total)\n", ctas, W, threads,
             ps ? "planes 8 KB apart" : "contiguous       ", bytes / ms / 1e6, bytes / (ms * 1e-3) / (clk_khz * 1e3), clk_khz / 1e6, bytes * ctas / ms / 1e6);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
