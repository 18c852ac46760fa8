// Probe: what does a TMA TENSOR store stream reach when every box row is one 128-byte line of a different channel plane
// (the drain pattern of splat_tma.cu: out[n][c][p], bf16, P = 4096 -> planes 8 KB apart)?
//   box (64 px, ROWS channels, 1), SWIZZLE_128B, ROWS in {32, 64, 128};  W issuing warps per CTA, each with a ring of D boxes.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_store_bw tma_store_bw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int P = 4096, C = 320;

// item = (image, 64-px column, channel block of `rows`): linear index -> coordinates
__global__ void __launch_bounds__(256) store_stream(const __grid_constant__ CUtensorMap tm, int n_img, int rows, int warps, int depth) {
  extern __shared__ __align__(1024) unsigned char sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < warps * depth * rows * 128 / 4; i += blockDim.x) reinterpret_cast<float*>(sm)[i] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (warp >= warps || lane != 0) return;
  const int cblocks = (C + rows - 1) / rows, pxb = P / 64;
  const long long total = (long long)n_img * pxb * cblocks;
  const long long stride = (long long)gridDim.x * warps;
  int it = 0;
  for (long long i = (long long)blockIdx.x * warps + warp; i < total; i += stride, ++it) {
    const int cb = (int)(i % cblocks), px = (int)((i / cblocks) % pxb), n = (int)(i / ((long long)cblocks * pxb));
    if (it >= depth) {
      switch (depth) {
        case 1: asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); break;
        case 2: asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); break;
        case 3: asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory"); break;
        case 4: asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory"); break;
        case 6: asm volatile("cp.async.bulk.wait_group.read 5;" ::: "memory"); break;
        default: asm volatile("cp.async.bulk.wait_group.read 7;" ::: "memory"); break;
      }
    }
    const uint32_t src = (uint32_t)__cvta_generic_to_shared(sm + ((size_t)(warp * depth + it % depth)) * rows * 128);
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(&tm), "r"(src), "r"(px * 64), "r"(cb * rows), "r"(n) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main() {
  const int n_img = 1024;
  const size_t bytes = (size_t)n_img * C * P * 2;
  void* buf; CK(cudaMalloc(&buf, bytes));
  int sms; CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr; cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  CK(cudaFuncSetAttribute(store_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  for (int rows : {32, 64, 128}) for (int warps : {1, 2, 4, 8}) for (int depth : {2, 3, 4, 8}) {
    const size_t smem = (size_t)warps * depth * rows * 128;
    if (smem > 200 * 1024) continue;
    CUtensorMap tm;
    cuuint64_t dims[3] = {P, C, (cuuint64_t)n_img}; cuuint64_t strides[2] = {P * 2, (cuuint64_t)C * P * 2};
    cuuint32_t box[3] = {64, (cuuint32_t)rows, 1}, es[3] = {1, 1, 1};
    if (enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
            CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed\n"); return 1; }
    for (int i = 0; i < 2; ++i) store_stream<<<sms, 256, smem>>>(tm, n_img, rows, warps, depth);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
      CK(cudaEventRecord(a)); store_stream<<<sms, 256, smem>>>(tm, n_img, rows, warps, depth); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b)); if (ms < best) best = ms;
    }
    printf("box 64px x %3d ch  issuing warps %d  ring %d (%3zu KB in flight/SM): %7.1f GB/s\n", rows, warps, depth, smem / 1024, bytes / best / 1e6);
  }
  return 0;
}
