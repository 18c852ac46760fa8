// Probe: can a 4-D tiled tensor map with NON-monotonic strides deliver the feature matrix [K, C] (C contiguous, 16-bit)
// straight into the canonical no-swizzle N-major UMMA operand layout used by tc_stage_b — item q = (k-group, channel
// chunk, row) at byte 16*q — with the operand-row offset (rows -3..-1 zero) and the K tail zero-filled by the TMA unit?
//   dims (fastest first): d0 = 8 elements of a 16-byte chunk, d1 = k row (stride 2C bytes), d2 = channel chunk (stride
//   16 bytes), d3 = image (stride 2KC bytes);  box = (8, 8, c_tile/8, 1), one TMA per k-group of 8 rows.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tma_b_layout tma_b_layout.cu   (driver entry point via cudart)
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int K = 33, C = 640, N = 3, KOFF = 3, KP = 48, CT = 320;

__global__ void probe(const __grid_constant__ CUtensorMap tm, unsigned short* out, int n, int c0) {
  extern __shared__ __align__(1024) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar;
  const unsigned bar_a = (unsigned)__cvta_generic_to_shared(&bar), dst = (unsigned)__cvta_generic_to_shared(sm);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned bytes = KP * CT * 2;
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(bytes) : "memory");
    for (int kg = 0; kg < KP / 8; ++kg) {
      const int c_e = 0, c_k = kg * 8 - KOFF, c_ch = c0 / 8, c_n = n;
      asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
                   ::"r"(dst + kg * CT * 16), "l"(&tm), "r"(c_e), "r"(c_k), "r"(c_ch), "r"(c_n), "r"(bar_a) : "memory");
    }
  }
  unsigned ok = 0;
  while (!ok) asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\nselp.u32 %0, 1, 0, p;\n}" : "=r"(ok) : "r"(bar_a) : "memory");
  for (int i = threadIdx.x; i < KP * CT; i += blockDim.x) out[i] = reinterpret_cast<unsigned short*>(sm)[i];
}

int main() {
  std::vector<__nv_bfloat16> h((size_t)N * K * C);
  for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) for (int c = 0; c < C; ++c)
    h[((size_t)n * K + k) * C + c] = __float2bfloat16((float)(n * 1000 + k * 7 + (c % 97)));      // exact in bf16? small ints < 256 exact; use check by recompute
  __nv_bfloat16* d; CK(cudaMalloc(&d, h.size() * 2)); CK(cudaMemcpy(d, h.data(), h.size() * 2, cudaMemcpyHostToDevice));
  unsigned short* out; CK(cudaMalloc(&out, KP * CT * 2));
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  cudaDriverEntryPointQueryResult qr;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qr));
  if (!enc) { printf("no cuTensorMapEncodeTiled entry point\n"); return 1; }
  CUtensorMap tm;
  cuuint64_t dims[4] = {8, (cuuint64_t)K, (cuuint64_t)(C / 8), (cuuint64_t)N};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, 16, (cuuint64_t)K * C * 2};
  cuuint32_t box[4] = {8, 8, CT / 8, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, d, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  printf("cuTensorMapEncodeTiled (non-monotonic strides) -> %d\n", (int)r);
  if (r != CUDA_SUCCESS) return 2;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, KP * CT * 2));
  int bad = 0;
  for (int n : {0, 2}) for (int c0 : {0, 320}) {
    probe<<<1, 128, KP * CT * 2>>>(tm, out, n, c0);
    CK(cudaDeviceSynchronize());
    std::vector<unsigned short> o(KP * CT);
    CK(cudaMemcpy(o.data(), out, o.size() * 2, cudaMemcpyDeviceToHost));
    // expected: item q = (kg, chunk, row): 8 channels c0 + chunk*8 .. +7 of operand row kg*8 + row (k = that - KOFF)
    for (int kg = 0; kg < KP / 8; ++kg) for (int ch = 0; ch < CT / 8; ++ch) for (int row = 0; row < 8; ++row) for (int e = 0; e < 8; ++e) {
      const int k = kg * 8 + row - KOFF, c = c0 + ch * 8 + e;
      const __nv_bfloat16 want = (k >= 0 && k < K) ? h[((size_t)n * K + k) * C + c] : __float2bfloat16(0.f);
      const unsigned short got = o[(((size_t)kg * (CT / 8) + ch) * 8 + row) * 8 + e];
      if (got != *reinterpret_cast<const unsigned short*>(&want)) { if (bad < 5) printf("mismatch n=%d c0=%d kg=%d ch=%d row=%d e=%d got %04x\n", n, c0, kg, ch, row, e, got); ++bad; }
    }
  }
  printf(bad ? "LAYOUT MISMATCH (%d)\n" : "layout OK: TMA writes the no-swizzle N-major operand directly (%d mismatches)\n", bad);
  return bad ? 3 : 0;
}
