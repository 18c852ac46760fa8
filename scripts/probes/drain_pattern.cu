// Probe: which store pattern can the drain of a D[channel, pixel] stage-3 engine use?  No TMEM, no shared memory, no operands:
// persistent CTAs walk items (image, 256-pixel tile, 128-channel group) of a [N][C][P] 16-bit tensor exactly like
// splat_tma.cu and only issue the stores, in one of three shapes per warp instruction:
//   0  four 128-byte lines of four channel planes (what the staged drain does: 8 lanes per line)
//   1  one 512-byte run of one channel plane (32 lanes x 16 B)
//   2  thirty-two 32-byte sectors of 32 channel planes (256-bit store per lane; the un-staged "sector" drain)
//   3  two 256-byte runs of two channel planes
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o drain_pattern drain_pattern.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <string>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s line %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

template <int MODE>
__global__ void __launch_bounds__(512) drain(unsigned char* out, int N, int C, int P, int items, int tiles, int groups) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, warps = blockDim.x >> 5;
  const int per = (items + gridDim.x - 1) / gridDim.x;
  const int begin = blockIdx.x * per, end = min(items, begin + per);
  const uint4 v = make_uint4(lane, warp, 3u, 4u);
  for (int item = begin; item < end; ++item) {
    const int g = item % groups, t = (item / groups) % tiles, n = item / (groups * tiles);
    const int npx = min(256, P - t * 256);
    // 128 channels x npx pixels, split over the warps by channel quarter (and pixel half when there are more than 4 warps)
    const int q = warp & 3, part = warp >> 2, parts = warps >> 2;
    unsigned char* const base = out + (((size_t)n * C + g * 128 + q * 32) * P + t * 256) * 2;
    const size_t plane = (size_t)P * 2;
    const int ch0 = g * 128 + q * 32;
    for (int b = part; b * 64 < npx; b += parts) {              // 64-pixel boxes: 32 channels x 128 B
      unsigned char* const bx = base + b * 128;
      if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) if (ch0 + 4 * j + (lane >> 3) < C) __stcs(reinterpret_cast<uint4*>(bx + (size_t)(4 * j + (lane >> 3)) * plane + (lane & 7) * 16), v);
      } else if (MODE == 2) {
#pragma unroll
        for (int j = 0; j < 4; ++j) if (ch0 + lane < C)
          asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%1,%2,%3,%4};" ::"l"(bx + (size_t)lane * plane + j * 32), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
      }
    }
    if (MODE == 1) {                                            // 32 channels x npx pixels as whole 512-byte runs
      for (int r = part; r < 32 && ch0 + r < C; r += parts)
        for (int x = lane * 16; x < npx * 2; x += 512) __stcs(reinterpret_cast<uint4*>(base + (size_t)r * plane + x), v);
    } else if (MODE == 3) {
      for (int r = 2 * part; r < 32 && ch0 + r + 1 < C; r += 2 * parts)
        for (int x = (lane & 15) * 16; x < npx * 2; x += 256) __stcs(reinterpret_cast<uint4*>(base + (size_t)(r + (lane >> 4)) * plane + x), v);
    }
  }
}

// drain_pattern sustained <seconds> [mode]: the cfg5c-sized store stream (mode 1 by default) back to back for that long, one line per
// ~0.25 s — does the bare store stream hold its burst rate once the power state has ramped?  (Sample clocks beside it with nvidia-smi.)
static int sustained(double seconds, int mode) {
  int clk_khz, sms; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int N = 1024, C = 320, P = 4096;
  const size_t bytes = (size_t)N * C * P * 2;
  unsigned char* buf; CK(cudaMalloc(&buf, bytes));
  const int tiles = (P + 255) / 256, groups = (C + 127) / 128, items = N * tiles * groups;
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  double elapsed = 0;
  while (elapsed < seconds) {
    const int reps = 500;
    CK(cudaEventRecord(a));
    for (int i = 0; i < reps; ++i) {
      if (mode == 0) drain<0><<<sms, 512>>>(buf, N, C, P, items, tiles, groups);
      else drain<1><<<sms, 512>>>(buf, N, C, P, items, tiles, groups);
    }
    CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    elapsed += ms * 1e-3;
    printf("t = %5.2f s  mode %d: %7.1f us per 2.7 GB  %5.2f TB/s\n", elapsed, mode, ms / reps * 1e3, bytes / (ms / reps) / 1e9);
    fflush(stdout);
  }
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 2 && std::string(argv[1]) == "sustained") return sustained(atof(argv[2]), argc > 3 ? atoi(argv[3]) : 1);
  int clk_khz, sms; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0)); CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
  struct Shape { const char* name; int N, C, P; } shapes[] = {{"cfg3 level 32 (84 MB)", 64, 640, 1024}, {"cfg3 level 16 (42 MB)", 64, 1280, 256},
                                                             {"cfg3 level 64 (168 MB)", 64, 320, 4096}, {"cfg5c grid (2.7 GB)", 1024, 320, 4096}};
  for (const Shape& s : shapes) {
    const size_t bytes = (size_t)s.N * s.C * s.P * 2;
    unsigned char* buf; CK(cudaMalloc(&buf, bytes));
    const int tiles = (s.P + 255) / 256, groups = (s.C + 127) / 128, items = s.N * tiles * groups;
    for (int mode = 0; mode < 4; ++mode) for (int threads : {256, 512}) {
      auto launch = [&] {
        if (mode == 0) drain<0><<<sms, threads>>>(buf, s.N, s.C, s.P, items, tiles, groups);
        else if (mode == 1) drain<1><<<sms, threads>>>(buf, s.N, s.C, s.P, items, tiles, groups);
        else if (mode == 2) drain<2><<<sms, threads>>>(buf, s.N, s.C, s.P, items, tiles, groups);
        else drain<3><<<sms, threads>>>(buf, s.N, s.C, s.P, items, tiles, groups);
      };
      for (int i = 0; i < 3; ++i) launch();
      CK(cudaDeviceSynchronize());
      const int reps = 20;
      CK(cudaEventRecord(a)); for (int i = 0; i < reps; ++i) launch(); CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
      float ms; CK(cudaEventElapsedTime(&ms, a, b)); ms /= reps;
      // C not a multiple of 128 wastes nothing here: groups cover C exactly for these shapes except 320 = 2.5 groups (last half group writes past C? no: guard)
      printf("%-24s mode %d  %3d threads: %8.1f us  %6.2f TB/s  %5.1f B/clk/SM @%.2f GHz\n", s.name, mode, threads, ms * 1e3, bytes / ms / 1e9,
             bytes / (ms * 1e-3) / sms / (clk_khz * 1e3), clk_khz / 1e6);
    }
    CK(cudaFree(buf));
  }
  return 0;
}
