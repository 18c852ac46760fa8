// Probe: which (TMEM lane, column) does each thread's register get from tcgen05.ld.16x256b.xN ?
// Build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -o /tmp/p scripts/probes/tmem_ld_layout.cu && /tmp/p
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t base_s;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&base_s)), "r"(64u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = base_s;
  const uint32_t row = warp * 32 + lane;
  const uint32_t laddr = tmem + ((uint32_t)(warp * 32) << 16);
  for (int g = 0; g < 4; ++g) {
    uint32_t v[8];
    for (int j = 0; j < 8; ++j) v[j] = row * 1000 + g * 8 + j;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(laddr + g * 8), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  for (int h = 0; h < 2; ++h) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                   "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(laddr + ((uint32_t)(h * 16) << 16)) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) out[((warp * 2 + h) * 32 + lane) * 16 + j] = r[j];
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64u) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 4 * 2 * 32 * 16 * 4);
  probe<<<1, 128>>>(d);
  cudaError_t e = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(e));
  static uint32_t h[4 * 2 * 32 * 16];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int w = 0; w < 4; ++w) for (int hh = 0; hh < 2; ++hh) for (int t = 0; t < 32; ++t) for (int j = 0; j < 16; ++j) {
    const uint32_t v = h[((w * 2 + hh) * 32 + t) * 16 + j];
    const int i = j >> 2, e2 = j & 3;
    const uint32_t exp_row = w * 32 + hh * 16 + t / 4 + 8 * (e2 >> 1), exp_col = 8 * i + 2 * (t % 4) + (e2 & 1);
    if (v != exp_row * 1000 + exp_col) { if (bad < 40) printf("w%d h%d t%d r%d: got row %u col %u, expected row %u col %u\n", w, hh, t, j, v / 1000, v % 1000, exp_row, exp_col); ++bad; }
  }
  printf("mismatches vs hypothesis (row = t/4 + 8*(j&2?1:0), col = 8*(j/4) + 2*(t%%4) + (j&1)): %d\n", bad);
  for (int t = 0; t < 8; ++t) { printf("w0 h0 t%d:", t); for (int j = 0; j < 8; ++j) printf(" %u", h[t * 16 + j]); printf("\n"); }
  return 0;
}
