// Fused render — stages 1+2+3 in ONE persistent, warp-specialised kernel on tcgen05 tensor cores.
//
// Replaces the whole of splat_features (blobctrl/utils/utils.py:80-241) for interp_size ==
// score_size: blob parameters + features [N,K,C] -> composed maps [N,K,H,W] and feature grid
// [N,C,H,W].  The per-pixel weights never leave the SM:
//
//   D[pixel, channel] = sum_k A[pixel, k] * B[k, channel]         (M = 128 pixels, N = C tile, K = M_blobs+1)
//
//   A  the composed weights d_k, produced by stages 1+2 on CUDA cores (lane = pixel, the serial
//      front-to-back walk of scores.cu) and written with tcgen05.st straight into TENSOR MEMORY —
//      TMEM lane = pixel, column = k — i.e. the MMA's A operand lives in TMEM (".ts" form).
//   B  the image's features, staged once per work unit in shared memory in the canonical K-major
//      no-swizzle UMMA layout (8-row x 16-byte core matrices; SBO = 128 B, LBO = C_tile*16 B).
//   D  fp32 accumulators in TMEM (C_tile <= 320 columns), split in two channel halves so the MMA of
//      half h of tile t+1 overlaps the drain of the other half of tile t.
//
// Precision (kSplit): float32 maps are split-precision MMAs with fp32 accumulation, two forms:
//   kSplit = 2 (default)  2xFP16: x = x1 + x2, x1 = fp16(x), x2 = fp16(x - x1) for the weights and for the features (the
//              features of a unit are first scaled by a power of two so that max|f| lands in [0.5, 1); the drain undoes it
//              exactly).  D = A1*B1 + A1*B2 + A2*B1 under kind::f16: every product is exact in the fp32 accumulator, the
//              dropped A2*B2 term and the fp16 subnormal resolution are <= 2^-24 of the unit's scale (measured ~2e-7).
//              Three K16 MMAs per 16 blobs: 15 tensor instructions per half tile at K = 65 — 44% less tensor time than
//              3xTF32, half the B footprint in shared memory (92 KB) and half the A columns in tensor memory.
//   kSplit = 1            3xTF32: hi = rna_tf32(x), lo = rna_tf32(x - hi); D = Ahi*Bhi + Ahi*Blo + Alo*Bhi — product
//              error ~3*2^-22, 27 K8 MMAs per half tile.  Kept as the A/B partner (BLOBSPLAT_F32_SPLIT=tf32).
// bf16/f16 maps (kSplit = 0) use one kind::f16 MMA per k-step.
//
// Warp roles (512 threads, 1 CTA/SM, persistent over work units = (image, channel chunk, tile range)):
//   warps 0-7   stages 1+2 for one 128-pixel tile, two warps per TMEM lane quarter: each composites one
//               of two blob ranges (a two-level multiplicative suffix scan across blobs, carry through
//               shared memory), d_k -> global composed planes (coalesced) and a shared-memory stash;
//               then, once the MMA has released A, stash -> hi/lo -> tcgen05.st.  Also compute the unit's
//               blob coefficients, and stage its features (B) when shared memory holds only one B buffer.
//   warps 8-11  epilogue: tcgen05.ld D half -> registers -> coalesced global stores of [N,C,H,W].
//   warp  12    TMEM allocation + single-thread tcgen05.mma issue + tcgen05.commit to mbarriers.
//   warps 13-15 operand staging: when several B buffers fit (16-bit maps, small K) they fill a ring of up to 4
//               ahead of the units, so a unit's features are in place before its first tile reaches the MMA.
#pragma once
#include <cstdlib>

#include "common.cuh"

// compile-time tuning knobs (scripts/ab_variants.py builds variants and A/Bs them in one process)
// ablation switches (measurement only — they change the results): which role bounds the pipeline?
#ifndef BS_ABL_NO_EPI_STORE
#define BS_ABL_NO_EPI_STORE 0   // epilogue drains TMEM but skips the global stores
#endif
#ifndef BS_ABL_NO_COMP_STORE
#define BS_ABL_NO_COMP_STORE 0  // stages 1+2 skip the composed-map global stores
#endif
#ifndef BS_ABL_NO_COMPUTE
#define BS_ABL_NO_COMPUTE 0     // stages 1+2 skip the blob evaluation (opacity := 0.01)
#endif
#ifndef BS_ABL_NO_MMA
#define BS_ABL_NO_MMA 0         // the MMA warp issues no tcgen05.mma (only the commits)
#endif
#ifndef BS_STAGE_COST
#define BS_STAGE_COST 3        // cost of staging one unit's operands, in tiles (loads queue behind the saturated store stream)
#endif
#ifndef BS_RING_MIN_UNITS
#define BS_RING_MIN_UNITS 3   // use the B ring when a CTA processes at least this many units
#endif
#ifndef BS_B_NMAJOR
#define BS_B_NMAJOR 1         // 16-bit maps: N-major B operand (raw 16-byte copies) instead of K-major (transposed)
#endif
#ifndef BS_SCORE_LOADS
#define BS_SCORE_LOADS 20
#endif
#ifndef BS_TIMING
#define BS_TIMING 0
#endif
#ifndef BS_SPLIT_NUM
#define BS_SPLIT_NUM 7          // back range's share of the blobs, in 16ths (it also carries the rescale pass)
#endif
#ifndef BS_SPLIT_ALIGN
#define BS_SPLIT_ALIGN 8        // both blob ranges are whole groups of this many blobs when M allows
#endif
#ifndef BS_D_PARTS
#define BS_D_PARTS 1            // accumulate / drain the channel tile in 64-channel parts (else two halves)
#endif
#ifndef BS_A_BUFS
#define BS_A_BUFS 2             // double-buffer the A operand in tensor memory when the columns allow
#endif
#ifndef BS_ALIGNED_SPLIT
#define BS_ALIGNED_SPLIT 1      // split the blobs on an operand k-group boundary: one pair barrier per tile instead of three
#endif
#ifndef BS_SKIP_UNIT_SCALE
#define BS_SKIP_UNIT_SCALE 1    // kSplit = 2: no per-unit feature scale when max|f| is in [0.5, 4096)
#endif
#ifndef BS_MAX_B
#define BS_MAX_B 4             // B operand ring depth (1 = staged by the compute warps between units)
#endif
#ifndef BS_RNA_CUSTOM
#define BS_RNA_CUSTOM 1
#endif

#if BS_TIMING
__device__ unsigned long long g_tc_timing[16];   // clock64 stamps of CTA 0 (debug builds only)
#define TC_STAMP(i) do { if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) ::g_tc_timing[i] = clock64(); } while (0)
#else
#define TC_STAMP(i) do { } while (0)
#endif

namespace blobsplat {

// kHalves = 1: 4 compute warps (one per TMEM lane quarter); kHalves = 2: 8 compute warps, two per quarter, each
// compositing one of two blob ranges.  Threads = (4*kHalves compute + 4 epilogue + 1 MMA) warps = 288 / 416 (+ 3 staging warps = 512 in the kRing instantiations).
constexpr int kTcTileM = 128;
constexpr int kTcMaxCTile = 320;
// Plane k (0 = background, m+1 = blob m) lives at operand row / TMEM column / stash column k + kTcKOff: with the
// offset 3 every group of 8 blobs of a range that ends on a multiple of 8 occupies two 16-byte aligned float4s of the
// pixel-major stash, so stage 1/2, the rescale pass and the TMEM conversion all use 128-bit shared-memory accesses.
constexpr int kTcKOff = 3;
constexpr int kTcMaxB = 4;                       // ring of B operand buffers (when they fit)
constexpr int kTcMaxParts = 8;                   // accumulator parts per tile (ring of MMA -> drain hand-overs)
constexpr int kTcStageWarps = 3;                 // staging warps: with the 13 others a CTA is 16 warps (registers are
                                                 // allocated for warp counts rounded up to 4 anyway)
constexpr int kTcMaxBlobs = 127;                 // coefficient table: 127 * 32 B
constexpr size_t kTcSmemBudget = 227 * 1024 - 256;

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");   // suspend-time hint (ns)
  return ok != 0;
}
// Bounded wait: a protocol bug becomes a CUDA error (trap) after ~2 s instead of a hung GPU.  try_wait already
// suspends the warp in hardware for a while; the optional nanosleep trades wake-up latency for issue slots.
#ifndef BS_SPIN_SLEEP_NS
#define BS_SPIN_SLEEP_NS 0
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (BS_SPIN_SLEEP_NS > 0) __nanosleep(BS_SPIN_SLEEP_NS);
    if (((++spins) & 0x3ffu) == 0 && clock64() - t0 > 4000000000ll) __trap();
  }
}
// Busy wait (mbarrier.test_wait, no hardware suspend): for single-thread roles (TMA producer, MMA issue) whose hand-overs are on
// the critical path of every tile — a suspended try_wait wakes up several hundred cycles after the phase flips.
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t a = smem_u32(bar);
  uint32_t ok = 0, spins = 0;
  long long t0 = 0;
  while (true) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(a), "r"(parity) : "memory");
    if (ok) return;
    if (((++spins) & 0xffffu) == 0) {
      if (t0 == 0) t0 = clock64();
      else if (clock64() - t0 > 4000000000ll) __trap();
    }
  }
}
#ifndef BS_MMA_SPIN
#define BS_MMA_SPIN 0          // fused kernels: the MMA-issuing thread busy-waits on its barriers (measured at cfg5b, sustained: 1.358 vs 1.314 ms suspended — spinning costs power)
#endif
#ifndef BS_DRAIN_SPIN
#define BS_DRAIN_SPIN 0        // the drain warps busy-wait for their accumulators (1.417 ms: worse still)
#endif
__device__ __forceinline__ void mbar_wait_mma(uint64_t* bar, uint32_t parity) {
  if (BS_MMA_SPIN) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity);
}
__device__ __forceinline__ void mbar_wait_drain(uint64_t* bar, uint32_t parity) {
  if (BS_DRAIN_SPIN) mbar_wait_spin(bar, parity); else mbar_wait(bar, parity);
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
// 16 lanes x 32 bit x 2, 16 repeats: threads 0-15 get columns c .. c+15 of the 16 lanes at the address's lane,
// threads 16-31 columns c+16 .. c+31 of the same lanes (layout measured with scripts/probes/tmem_ld_layout.cu).
__device__ __forceinline__ void tmem_ld_16x32bx2_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x32bx2.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16], 16;"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem desc]
template <bool kTf32>
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                        uint32_t accumulate) {
  if constexpr (kTf32) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
  }
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 | [46,48) version = 1 | [61,64) layout = 0
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) |
         ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
// cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format @7/@10, A K-major, b_major @16 (1 = N-major), N>>3 @17,
// M>>4 @24
__device__ __forceinline__ uint32_t make_idesc(uint32_t ab_format, uint32_t n, uint32_t b_n_major) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | (b_n_major << 16) | ((n >> 3) << 17) |
         ((uint32_t)(kTcTileM >> 4) << 24);
}

// Round-to-nearest (ties away) to TF32's 10 explicit mantissa bits: (bits + 0x1000) & ~0x1fff.  Same result as
// cvt.rna.tf32.f32 for finite values (weights are in [0,1], features finite) at 2 integer ops instead of ~5.
__device__ __forceinline__ float rna_tf32(float x) {
#if BS_RNA_CUSTOM
  return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xffffe000u);
#else
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
#endif
}

// Two values bound for two different planes.  For 16-bit maps one packed convert (F2FP) serves both stores —
// F2FP + PRMT + 2 STG per pair instead of 2 x (F2FP + PRMT + STG).
template <typename OT>
__device__ __forceinline__ void store_pair(OT* p0, OT* p1, float a, float b, bool pred) {
  if constexpr (std::is_same<OT, __nv_bfloat16>::value) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    if (pred) { __stcs(p0, __low2bfloat16(t)); __stcs(p1, __high2bfloat16(t)); }
  } else if constexpr (std::is_same<OT, __half>::value) {
    const __half2 t = __floats2half2_rn(a, b);
    if (pred) { __stcs(p0, __low2half(t)); __stcs(p1, __high2half(t)); }
  } else {
    if (pred) { __stcs(p0, (OT)a); __stcs(p1, (OT)b); }
  }
}

// One 32-channel block of the pixel-pair epilogue (float maps): r[j] / r[16 + j] are channel j of this thread's two
// adjacent pixels, stored as one 64-bit word; a half-warp covers one whole 128-byte line of a channel plane.
template <bool kScaled>
__device__ __forceinline__ void store_pixel_pairs(float* oc, size_t P, const uint32_t* r, bool pred, float inv) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float2 v = make_float2(__uint_as_float(r[j]), __uint_as_float(r[16 + j]));
    if constexpr (kScaled) { v.x *= inv; v.y *= inv; }        // undo the unit's power-of-two feature scale (exact)
    if (pred) __stcs(reinterpret_cast<float2*>(oc + (size_t)j * P), v);
  }
}

// G (8 or 4) blobs m-G+1 .. m of one pixel, front to back: branch-free, their MUFU/FMA chains interleave and only the
// transmittance T is a serial dependence (one FFMA per blob).  d_k goes to the composed planes (when wr) and to the stash
// as aligned float4s — m must be a multiple of 4 (plane m sits at stash column m + kTcKOff... the caller passes my = row + kTcKOff).
template <typename OT, int G>
__device__ __forceinline__ void composite_group(const BlobCoef* coef, int m, float xf, float yf, float& T, float* my, OT* comp_px,
                                                size_t P, bool wr) {
  float s[G], d[G];
#pragma unroll
  for (int j = 0; j < G; ++j) s[j] = BS_ABL_NO_COMPUTE ? 0.01f : blob_opacity_pd(coef[m - 1 - j], xf, yf);
  OT* const cp = comp_px + (size_t)m * P;      // plane k = m; the group's other planes are immediates
#pragma unroll
  for (int j = 0; j < G; ++j) {
    d[j] = s[j] * T;
    T = fmaf(-s[j], T, T);
  }
#pragma unroll
  for (int j = 0; j < G; j += 2) store_pair<OT>(cp - (ptrdiff_t)j * P, cp - (ptrdiff_t)(j + 1) * P, d[j], d[j + 1], wr);
  if constexpr (G == 8) {
    *reinterpret_cast<float4*>(my + m - 7) = make_float4(d[7], d[6], d[5], d[4]);
    *reinterpret_cast<float4*>(my + m - 3) = make_float4(d[3], d[2], d[1], d[0]);
  } else {
    *reinterpret_cast<float4*>(my + m - 3) = make_float4(d[3], d[2], d[1], d[0]);
  }
}

struct RenderTcParams {
  const float* xs; const float* ys; const float* covs; const float* sizes;
  const void* scores; long long sn, sk, sp;   // kFromScores: precomputed weights [N,K,P] with element strides
  const void* feats; void* composed; void* grid;
  int N, M, H, W, C;
  int K, Kp;              // K = M + 1; Kp = K rounded up to the MMA k-step (8 tf32 / 16 f16)
  int c_tile, c_chunks;   // channels per work unit (multiple of 32, <= 320); ceil(C / c_tile)
  int tiles_per_image;    // 128-pixel tiles per image
  int cw;                 // render_tc2 only: channels per drain sub-step
  int nb;                 // B operand buffers in shared memory (ring); > 1: the staging warps run ahead of the units
  int smem_bytes;         // dynamic shared memory of the launch (fixed part + nb B slots)
  int whole_runs;         // schedule: whole (image, chunk) runs round-robin vs contiguous equal tile ranges
  int total_tiles;        // N * c_chunks * tiles_per_image, linear index ((n * c_chunks + chunk) * tiles_per_image + tile)
  int pair_ok;            // float maps: grid planes allow aligned 2-pixel stores (P even, base 8-byte aligned)
  int tmem_cols;          // tensor-memory columns to allocate: accumulator + A operands, rounded up to a power of two
  int n_parts;            // the channel tile is accumulated and drained in n_parts parts of c_tile / n_parts channels: the MMAs of
                          // part j of tile t + 1 start as soon as part j of tile t has been drained
  int a_bufs;             // A operand buffers in tensor memory (2: stages 1+2 and the conversion run a tile ahead of the MMAs)
  int f32_split;          // float32 maps: 1 = 3xTF32, 2 = 2xFP16 (chosen per launch, split_for_launch)
  // render_tc2 with the halving pyramid fused in (64 x 64 maps, utils.py:280-294): the composed maps of levels 1..pyr_levels
  // ([N,K,32,32], [N,K,16,16], [N,K,8,8]) leave the same launch
  void* pyr[3]; int pyr_levels; int pyr_dbg;
};

// First tile of CTA i's range under the equal-shares schedule.
__host__ __device__ __forceinline__ int tc_range_begin(int total_tiles, int i, int ctas) {
  return (int)((long long)total_tiles * i / ctas);
}

// Kernel argument: one problem (the fused render, a single stage-3 splat) or up to kTcMaxLevels stage-3 problems of one
// pyramid that share K, dtype and channel tile and run as ONE launch over the concatenated tile sequence.
constexpr int kTcMaxLevels = 4;
struct RenderTcLevels {
  RenderTcParams lv[kTcMaxLevels];
  int n_levels;
  int tile_start[kTcMaxLevels + 1];   // first linear tile of each level; [n_levels] = total
};

struct TcBarriers {
  uint64_t a_full[2], a_free[2], b_full[kTcMaxB], b_free[kTcMaxB], d_full[kTcMaxParts], d_empty[kTcMaxParts];
  uint32_t tmem_base;
  float unit_inv[8];           // kSplit = 2: 1 / (power-of-two feature scale) of unit u in slot u & 7, applied by the drain.  8 slots:
                               // a slot is rewritten 8 units later, by which time the drain has long read it (its d_empty arrivals
                               // gate the MMAs of every unit in between)
  float red[8];                // scratch of the staging threads' max reduction (ring mode)
  float umax[8][4];            // kSplit = 2 without the ring: max|f| of unit u, one partial per drain warp, in slot u & 7
  uint64_t s_full[2];          // ... unit u arrives on s_full[u & 1] (4 arrivals).  Two barriers: the drain warps run up to one unit
                               // ahead of the staging that waits, and a single barrier's parity cannot tell phase u from u + 2
};

// Stage one unit's B operand: features [K, C] (c contiguous) of image n, channels c0 .. c0 + c_tile - 1, into the K-major
// no-swizzle operand layout at b_dst (second copy = TF32 residuals at + b_bytes).  Called by `nthreads` threads.
// kSplit = 2 additionally needs `sc` (TcBarriers: reduction scratch + the slot's inverse scale) and a named barrier shared by
// exactly the calling threads.
// max |f| over one unit's features (rows k < K, channels c0 .. c0 + c_tile - 1), this thread's share
template <typename FT>
__device__ __forceinline__ float unit_abs_max(const RenderTcParams& p, const FT* f, int c0, int tid, int nthreads) {
  const int cw = min(p.c_tile, p.C - c0);
  float mx = 0.0f;
  if constexpr (sizeof(FT) == 4) {
    if ((reinterpret_cast<uintptr_t>(p.feats) & 15) == 0 && (p.C & 3) == 0 && (cw & 3) == 0) {
      const int q4 = cw >> 2, items = p.K * q4;
      constexpr int kB = 8;                                   // loads in flight per thread
      for (int i0 = tid; i0 < items; i0 += kB * nthreads) {
        float4 v[kB];
#pragma unroll
        for (int b = 0; b < kB; ++b) {
          const int i = i0 + b * nthreads;
          const int k = i / q4, c = (i - k * q4) << 2;
          v[b] = i < items ? __ldg(reinterpret_cast<const float4*>(f + (size_t)k * p.C + c0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int b = 0; b < kB; ++b)
          mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v[b].x), fabsf(v[b].y)), fmaxf(fabsf(v[b].z), fabsf(v[b].w))));
      }
      return mx;
    }
  }
  for (int i = tid; i < p.K * cw; i += nthreads) {
    const int k = i / cw, c = i - k * cw;
    mx = fmaxf(mx, fabsf((float)Cvt<FT>::to(__ldg(f + (size_t)k * p.C + c0 + c))));
  }
  return mx;
}

template <typename FT, typename OT, int kSplit, bool kEpiMax = false>
__device__ __forceinline__ void tc_stage_b(const RenderTcParams& p, int n, int c0, unsigned char* b_smem, size_t b_bytes,
                                           int tid, int nthreads, TcBarriers* sc = nullptr, int unit = 0, int bar_id = 0) {
  constexpr bool kTf32 = kSplit == 1;
  constexpr bool kH2 = kSplit == 2;
  using BT = typename std::conditional<kTf32, float, typename std::conditional<kH2, __half, OT>::type>::type;
  if constexpr (kH2) {
    // Power-of-two scale of this unit's features: max|f| * scale in [0.5, 1), so fp16's absolute resolution (2^-24) is
    // relative to the unit's own magnitude whatever the features' range.  The max comes from the drain warps, which
    // computed it one unit ahead while they were waiting for accumulators (kEpiMax), or — with the B ring, where the
    // staging warps run several units ahead — from a first pass over the features here.
    const FT* f = reinterpret_cast<const FT*>(p.feats) + (size_t)n * p.K * p.C;
    float mx = 0.0f;
    if (kEpiMax) {
      mbar_wait(&sc->s_full[unit & 1], (uint32_t)((unit >> 1) & 1));
      mx = fmaxf(fmaxf(sc->umax[unit & 7][0], sc->umax[unit & 7][1]), fmaxf(sc->umax[unit & 7][2], sc->umax[unit & 7][3]));
    } else {
      mx = unit_abs_max<FT>(p, f, c0, tid, nthreads);
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      if ((tid & 31) == 0) sc->red[tid >> 5] = mx;
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");
      mx = 0.0f;
      for (int w = 0; w < (nthreads >> 5); ++w) mx = fmaxf(mx, sc->red[w]);
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(nthreads) : "memory");      // scratch may be rewritten by the next unit
    }
    int e = 0;
    if (mx > 0.0f && mx < 3.0e38f) (void)frexpf(mx, &e);                            // mx = m * 2^e, m in [0.5, 1)
    e = max(-100, min(100, e));
    // Units whose magnitude already sits in [0.5, 4096) stay unscaled: fp16 holds them without overflow and the absolute
    // resolution 2^-24 is <= 2^-23 of the unit's magnitude either way — and the drain skips its 2 multiplies per stored pair
    // (1 280 FMUL of the 14 400 warp instructions of a BlobNet tile; typical features are O(1)).
    if (BS_SKIP_UNIT_SCALE && mx >= 0.5f && mx < 4096.0f) e = 0;
    const float scale = exp2f((float)-e);
    if (tid == 0) sc->unit_inv[unit & 7] = exp2f((float)e);
    // features [K, C] -> K-major operand rows, x = f * scale split into x1 = fp16(x), x2 = fp16(x - x1).  One thread moves
    // blocks of 8 k-rows x 4 channels (8 128-bit loads, transposed in registers into 4 items per operand); two blocks per
    // iteration keep 16 loads in flight per thread, so the 83 KB of a BlobNet unit cost two load round trips.
    constexpr int kBlk = 2;
    const int cq = p.c_tile >> 2;
    const int blocks = (p.Kp >> 3) * cq;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(p.feats) & 15) == 0 && (p.C & 3) == 0;
    for (int q0 = tid; q0 < blocks; q0 += kBlk * nthreads) {
      float4 v[kBlk][8];
#pragma unroll
      for (int b = 0; b < kBlk; ++b) {
        const int qi = q0 + b * nthreads;
        const int kc = qi / cq, ch = c0 + ((qi - kc * cq) << 2);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kc * 8 + j - kTcKOff;                          // operand row k' = k + kTcKOff
          const bool ok = qi < blocks && k >= 0 && k < p.K && ch < p.C;
          v[b][j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok && vec_ok) {
            v[b][j] = __ldg(reinterpret_cast<const float4*>(f + (size_t)k * p.C + ch));
          } else if (ok) {
            const FT* src = f + (size_t)k * p.C + ch;
            v[b][j].x = (float)__ldg(src);
            if (ch + 1 < p.C) v[b][j].y = (float)__ldg(src + 1);
            if (ch + 2 < p.C) v[b][j].z = (float)__ldg(src + 2);
            if (ch + 3 < p.C) v[b][j].w = (float)__ldg(src + 3);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < kBlk; ++b) {
        const int qi = q0 + b * nthreads;
        if (qi >= blocks) break;
        const int kc = qi / cq, c = (qi - kc * cq) << 2;
        unsigned char* dst = b_smem + ((size_t)kc * p.c_tile + c) * 16;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          __half h1[8], h2[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float raw = cc == 0 ? v[b][j].x : (cc == 1 ? v[b][j].y : (cc == 2 ? v[b][j].z : v[b][j].w));
            const float x = raw * scale;                               // exact: a power of two
            h1[j] = __float2half_rn(x);
            h2[j] = __float2half_rn(x - __half2float(h1[j]));
          }
          *reinterpret_cast<uint4*>(dst + cc * 16) = *reinterpret_cast<const uint4*>(h1);
          *reinterpret_cast<uint4*>(dst + cc * 16 + b_bytes) = *reinterpret_cast<const uint4*>(h2);
        }
      }
    }
    return;
  }
  if constexpr (kSplit == 0 && BS_B_NMAJOR) {
    // 16-bit maps: B is an N-MAJOR operand (channels contiguous, as the features are stored) in the no-swizzle
    // canonical layout ((8 ch, 1, n), (8 k, groups)) : 8 k-rows x 16 bytes per core matrix, channel chunks 128 B
    // apart (SBO), k-groups c_tile*16 B apart (LBO).  Staging is a pure 16-byte copy, no conversion, no transposition:
    // item q = (k-group, channel chunk, row) goes to byte 16*q; a warp reads 8 feature rows x 64 contiguous bytes.
    static_assert(sizeof(FT) == 2 && std::is_same<FT, OT>::value, "16-bit staging copies raw elements");
    const FT* f = reinterpret_cast<const FT*>(p.feats) + (size_t)n * p.K * p.C;
    const int cq8 = p.c_tile >> 3;
    const int items = p.Kp * cq8;
    const bool vec_ok = (reinterpret_cast<uintptr_t>(p.feats) & 15) == 0 && (p.C & 7) == 0;
    uint4* dst = reinterpret_cast<uint4*>(b_smem);
    constexpr int kBatch = 8;     // 128 bytes in flight per thread: one round trip for a 30 KB operand on 256 threads
    for (int q0 = tid; q0 < items; q0 += kBatch * nthreads) {
      uint4 v[kBatch];
#pragma unroll
      for (int b = 0; b < kBatch; ++b) {
        const int q = q0 + b * nthreads;
        const int g = q >> 3, kg = g / cq8;
        const int k = kg * 8 + (q & 7) - kTcKOff, ch = c0 + (g - kg * cq8) * 8;      // operand row k' = k + kTcKOff
        v[b] = make_uint4(0u, 0u, 0u, 0u);
        if (q < items && k >= 0 && k < p.K && ch < p.C) {
          const FT* src = f + (size_t)k * p.C + ch;
          if (vec_ok) {
            v[b] = __ldg(reinterpret_cast<const uint4*>(src));
          } else {
            FT e[8];
#pragma unroll
            for (int cc = 0; cc < 8; ++cc) e[cc] = ch + cc < p.C ? __ldg(src + cc) : Cvt<FT>::from(0.0f);
            v[b] = *reinterpret_cast<const uint4*>(e);
          }
        }
      }
#pragma unroll
      for (int b = 0; b < kBatch; ++b)
        if (q0 + b * nthreads < items) dst[q0 + b * nthreads] = v[b];
    }
    return;
  }
  // features [K, C] (c contiguous) -> K-major operand rows: item (kc, c) = the T k-values kc*T .. kc*T+T-1 of
  // channel c as one 16-byte chunk.  One thread moves a T x VC block: T 128-bit global loads (VC adjacent
  // channels of one feature row each, coalesced across the warp), transposed in registers into VC items.
  const FT* f = reinterpret_cast<const FT*>(p.feats) + (size_t)n * p.K * p.C;
  constexpr int T = (16 / (int)sizeof(BT));
  constexpr int VC = 16 / sizeof(FT);
  const int cq = p.c_tile / VC;
  const int blocks = (p.Kp / T) * cq;
  const bool vec_ok = (reinterpret_cast<uintptr_t>(p.feats) & 15) == 0 && (p.C % VC) == 0;
  for (int qi = tid; qi < blocks; qi += nthreads) {
    const int kc = qi / cq, c = (qi - kc * cq) * VC;
    const int ch = c0 + c;
    float v[T][VC];
#pragma unroll
    for (int j = 0; j < T; ++j) {
      const int k = kc * T + j - kTcKOff;                          // operand row k' = k + kTcKOff
      const bool ok = k >= 0 && k < p.K && ch < p.C;
      if (ok && vec_ok) {
        const uint4 raw = __ldg(reinterpret_cast<const uint4*>(f + (size_t)k * p.C + ch));
        const FT* e = reinterpret_cast<const FT*>(&raw);
#pragma unroll
        for (int cc = 0; cc < VC; ++cc) v[j][cc] = (float)Cvt<FT>::to(e[cc]);
      } else {
#pragma unroll
        for (int cc = 0; cc < VC; ++cc)
          v[j][cc] = (ok && ch + cc < p.C) ? (float)Cvt<FT>::to(__ldg(f + (size_t)k * p.C + ch + cc)) : 0.0f;
      }
    }
    unsigned char* dst = b_smem + ((size_t)kc * p.c_tile + c) * 16;
#pragma unroll
    for (int cc = 0; cc < VC; ++cc) {
      if constexpr (kTf32) {
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { hi[j] = rna_tf32(v[j][cc]); lo[j] = rna_tf32(v[j][cc] - hi[j]); }
        *reinterpret_cast<float4*>(dst + cc * 16) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(dst + cc * 16 + b_bytes) = make_float4(lo[0], lo[1], lo[2], lo[3]);
      } else {
        OT h[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) h[j] = Cvt<OT>::from(v[j][cc]);
        *reinterpret_cast<uint4*>(dst + cc * 16) = *reinterpret_cast<const uint4*>(h);
      }
    }
  }
}

// FT: feature dtype in global memory; OT: output dtype; kSplit: 0 = 16-bit maps (kind::f16), 1 = float maps by 3xTF32,
// 2 = float maps by 2xFP16 (see the header)
// kP: pixels per image when it is one of the common sizes (64^2, 32^2, 16^2), else 0 = runtime.  With a
// compile-time plane stride every store of an unrolled group is [base + immediate]: no address arithmetic.
// kP = -1: several pyramid levels in one launch (RenderTcLevels); every work unit reads its own level's shape.
// kFromScores: the A operand comes from precomputed score maps in global memory (stand-alone stage 3,
// splat_features_from_scores) instead of being rendered from blob parameters (stages 1+2).
// kRing: the CTA has the 3 staging warps and a ring of p.nb >= 2 B buffers; otherwise one buffer, staged by the compute
// warps between units (the float path at BlobNet's sizes, where B fills shared memory, and launches with few units).
template <typename FT, typename OT, int kSplit, int kHalves, int kP, bool kFromScores, bool kRing>
__global__ void __launch_bounds__((4 * kHalves + 5 + (kRing ? kTcStageWarps : 0)) * 32, 1) render_tc_kernel(const __grid_constant__ RenderTcLevels L) {
  constexpr bool kTf32 = kSplit == 1;        // 3xTF32
  constexpr bool kH2 = kSplit == 2;          // 2xFP16
  constexpr bool kFloatMaps = kSplit != 0;   // float outputs: adjacent-pixel lane map, 64-bit pair stores
  const RenderTcParams& p0 = L.lv[0];     // Kp, c_tile and the dtypes are the same for every level
  if (threadIdx.x == 0) TC_STAMP(0);
  constexpr int kTcComputeWarps = 4 * kHalves;
  constexpr int kTcComputeThreads = kTcComputeWarps * 32;
  constexpr int kTcMmaWarp = kTcComputeWarps + 4;
  extern __shared__ __align__(1024) unsigned char smem[];
  using BT = typename std::conditional<kTf32, float, typename std::conditional<kH2, __half, OT>::type>::type;   // B element in smem
  constexpr int kElemsPer16B = 16 / sizeof(BT);                    // T: 4 (tf32) or 8 (16-bit)
  constexpr int kKStep = 2 * kElemsPer16B;                         // K per MMA: 8 or 16
  constexpr int kNumB = kFloatMaps ? 2 : 1;                        // hi + lo / x1 + x2
  constexpr int kACols = kTf32 ? 1 : 2;                            // k elements per 32-bit TMEM column

  const int n_parts = p0.n_parts;
  const int c_half = p0.c_tile / n_parts;          // channels per accumulator part (historic name: there used to be two)
  const size_t b_bytes = (size_t)(p0.Kp / kElemsPer16B) * p0.c_tile * 16;   // one B copy
  unsigned char* b_smem = smem;                                            // [nb][kNumB][Kp/T][c_tile][16 B]
  const int nb = kRing ? p0.nb : 1;
  const size_t b_stride = kNumB * b_bytes;                                 // one ring slot
  const int srow = p0.Kp + 4;                                               // stash row stride (floats): conflict-free LDS/STS.128
  float* stash = reinterpret_cast<float*>(smem + nb * b_stride);         // [128 pixels][Kp + 4] composed weights of one tile
  float* carry = stash + (size_t)srow * kTcTileM;                         // [2][128] front-range transmittance per pixel, by tile parity
  BlobCoef* coef = reinterpret_cast<BlobCoef*>(carry + 2 * kTcTileM);
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(coef + kTcMaxBlobs + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == kTcMmaWarp) {
    if (lane == 0) {
      mbar_init(&bars->a_full[0], kTcComputeThreads); mbar_init(&bars->a_free[0], 1);
      mbar_init(&bars->a_full[1], kTcComputeThreads); mbar_init(&bars->a_free[1], 1);
      for (int i = 0; i < kTcMaxB; ++i) {   // B is staged by the staging warps when there is a ring, else by the compute warps
        mbar_init(&bars->b_full[i], kRing ? kTcStageWarps * 32 : kTcComputeThreads); mbar_init(&bars->b_free[i], 1);
      }
      for (int i = 0; i < kTcMaxParts; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], 128); }
      mbar_init(&bars->s_full[0], 4); mbar_init(&bars->s_full[1], 4);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"((uint32_t)p0.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  if (threadIdx.x == 0) TC_STAMP(1);
  pdl_launch_dependents();
  pdl_wait();                                      // set-up above overlapped the previous kernel's tail
  const int a_cols = p0.Kp / kACols;                               // columns of one A operand
  const int a_bufs = p0.a_bufs;                                    // tile t uses buffer t % a_bufs
  const uint32_t tmem_a0 = tmem + (uint32_t)p0.c_tile;             // buffer 0: A hi (x1); A lo (x2) follows at + a_cols columns
  const uint32_t a_buf_cols = (uint32_t)((kFloatMaps ? 2 : 1) * a_cols);

  int unit_it = 0;      // units processed by this CTA so far
  int tile_it = 0;      // tiles processed by this CTA so far (barrier phases)

  // Work distribution over the linear tile sequence; a work unit = a run of tiles of one (image, channel chunk), whose
  // operands (blob coefficients, B) are staged once.  Two schedules, chosen on the host (fill_tc_units):
  //   whole_runs  CTA i takes the whole runs i, i + gridDim.x, ... (large batches: fewest stagings, CTAs in lock-step)
  //   otherwise   the sequence is cut into gridDim.x contiguous, equally long ranges (+-1 tile), which a CTA walks as
  //               partial runs (small batches: balance matters more than the extra stagings)
  // Several levels (kP = -1): always equal ranges over the concatenated sequence.
  const int seq_tiles = kP == -1 ? L.tile_start[L.n_levels] : p0.total_tiles;
  const bool whole_runs = kP != -1 && p0.whole_runs;
  const int g_end = whole_runs ? seq_tiles : tc_range_begin(seq_tiles, (int)blockIdx.x + 1, (int)gridDim.x);
  int level = 0;
  for (int gs = whole_runs ? (int)blockIdx.x * p0.tiles_per_image : tc_range_begin(seq_tiles, (int)blockIdx.x, (int)gridDim.x);
       gs < g_end; ++unit_it) {
    if constexpr (kP == -1) {
      while (gs >= L.tile_start[level + 1]) ++level;
    }
    const RenderTcParams& p = L.lv[kP == -1 ? level : 0];
    const int g = gs - (kP == -1 ? L.tile_start[level] : 0);        // tile index within the level
    const int img_chunk = g / p.tiles_per_image;
    const int n = img_chunk / p.c_chunks;
    const int chunk = img_chunk - n * p.c_chunks;
    const int c0 = chunk * p.c_tile;
    const int t_lo = g - img_chunk * p.tiles_per_image;
    const int ntiles = min(p.tiles_per_image - t_lo, g_end - gs);
    gs += ntiles;
    if (whole_runs) gs += ((int)gridDim.x - 1) * p.tiles_per_image;

    // plane stride: a compile-time constant for the single-level specialisations, per level at run time otherwise
    const int P = kP > 0 ? kP : p.H * p.W;
    if (warp < kTcComputeWarps) {
      // =============================== stages 1+2 + operand staging ===============================
      // 8 warps: two per TMEM lane quarter.  Warp (half, q) owns pixels q*32..q*32+31 of the tile and one of
      // two contiguous blob ranges: half 0 the front (high-index) range, half 1 the back range + background.
      const int ctid = warp * 32 + lane;            // 0..255 over the compute warps
      const int half = warp >> 2, q = warp & 3;
      const int px = q * 32 + lane;                 // TMEM lane / stash row of this thread
      // Its pixel within the tile.  Float maps: lanes L and L+16 of a quarter hold ADJACENT pixels, so the epilogue's
      // 16x32bx2 TMEM loads hand one thread two consecutive pixels of a channel (64-bit stores, half the store
      // instructions).  A warp still covers the same 32 consecutive pixels, so the composed-map stores stay coalesced.
      const int ppx = kFloatMaps ? q * 32 + ((lane & 15) << 1) + (lane >> 4) : px;
      asm volatile("bar.sync 1, %0;" ::"n"(kTcComputeThreads) : "memory");   // previous unit's tiles are done with `coef`
      // stage 3 from score maps: the planes of this warp, and the first tile's loads issued BEFORE the operand staging
      // so that the two global round trips of a unit's start overlap
      constexpr int kLd = BS_SCORE_LOADS;
      const int k_split = kHalves == 2 ? (p.K >> 1) : 0;
      const int k_lo = half ? 0 : k_split, k_hi = half ? k_split : p.K;
      OT pre[kFromScores ? kLd : 1];
      if constexpr (kFromScores) {
        const int pix0 = t_lo * kTcTileM + ppx;
        const OT* sc0 = reinterpret_cast<const OT*>(p.scores) + (size_t)n * p.sn + (size_t)(pix0 < P ? pix0 : 0) * p.sp;
#pragma unroll
        for (int j = 0; j < kLd; ++j)
          pre[j] = (pix0 < P && k_lo + j < k_hi) ? __ldg(sc0 + (size_t)(k_lo + j) * p.sk) : Cvt<OT>::from(0.0f);
      }
      uint32_t my_general = 0;
      if constexpr (!kFromScores)
      for (int i = ctid; i < p.M; i += kTcComputeThreads) {
        const size_t b = (size_t)n * p.M + i;
        const float* c = p.covs + 4 * b;
        const BlobCoef bc = make_blob_coef((double)p.xs[b], (double)p.ys[b], (double)c[0], (double)c[1], (double)c[2],
                                           (double)c[3], p.sizes[b], p.H, p.W);
        my_general |= coef_general(bc) ? 1u : 0u;
        coef[i] = bc;
      }
      if (unit_it == 0) {   // stash columns that are never written (k' < kTcKOff, k' >= K + kTcKOff) stay zero for the whole kernel
        for (int i = ctid; i < kTcTileM * srow; i += kTcComputeThreads) stash[i] = 0.0f;
      }
      if constexpr (!kRing) {   // single B buffer: staged here, after the MMAs of the previous unit have read it
        if (unit_it > 0) mbar_wait(&bars->b_free[0], (unit_it - 1) & 1);
        tc_stage_b<FT, OT, kSplit, kH2>(p, n, c0, b_smem, b_bytes, ctid, kTcComputeThreads, bars, unit_it, 1);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> tensor-core reads
        mbar_arrive(&bars->b_full[0]);
      }
      if (ctid == 0 && unit_it == 0) TC_STAMP(2);      // operands staged
      uint32_t any_general;   // barrier + OR-reduce: coef visible to all compute threads; does any blob need the slow form?
      asm volatile(
          "{\n\t.reg .pred q, r;\n\tsetp.ne.u32 q, %1, 0;\n\t"
          "barrier.cta.red.or.pred r, 1, %2, q;\n\tselp.u32 %0, 1, 0, r;\n\t}"
          : "=r"(any_general) : "r"(my_general), "n"(kTcComputeThreads) : "memory");

      // Two-level multiplicative suffix scan across blobs: each range is composited with a local
      // transmittance; the back range is then scaled by the front range's total transmittance.
      // The back range carries the extra rescale pass, so it gets the smaller share (7/16) of the blobs.
      // Aligned split (g_split >= 0): m_split is chosen so that the front range starts on an operand k-group boundary
      // (m_split + 1 + kTcKOff a multiple of the MMA's k-step).  Each warp then converts exactly the stash columns it wrote
      // itself, and the only hand-over left inside a quarter is the front range's transmittance: one named barrier per
      // tile instead of three (the ncu source page had 21 % of the compute warps' time in those barriers).
      int m_split = 0, g_split = -1;
      if (kHalves == 2) {
        m_split = ((p.M * BS_SPLIT_NUM) >> 4) & ~(BS_SPLIT_ALIGN - 1);
        if (BS_SPLIT_ALIGN == 8 && ((p.M - m_split) & 7) != 0 && (p.M & 7) == 0) m_split = (p.M * BS_SPLIT_NUM >> 4) & ~7;
        if (!kFromScores && BS_ALIGNED_SPLIT) {
          const int cand = ((((p.M * BS_SPLIT_NUM) >> 4) + 1 + kTcKOff + kKStep / 2) / kKStep) * kKStep - 1 - kTcKOff;
          if (cand >= 4 && cand * 4 >= p.M && cand * 8 <= p.M * 5) { m_split = cand; g_split = (cand + 1 + kTcKOff) / kKStep; }
        }
      }
      const int m_lo = half ? 0 : m_split, m_hi = half ? m_split : p.M;
      const int pair_bar = 2 + q;                   // named barrier of this quarter's two warps (64 threads)
      OT* comp = (p.composed && chunk == 0) ? reinterpret_cast<OT*>(p.composed) + (size_t)n * p.K * P : nullptr;
      for (int t = 0; t < ntiles; ++t, ++tile_it) {
        const int pix = (t_lo + t) * kTcTileM + ppx;
        const bool live = pix < P;
        const int y = live ? pix / p.W : 0;
        const float xf = (float)(live ? pix - y * p.W : 0), yf = (float)y;
        float* my = stash + (size_t)px * srow + kTcKOff;          // my[k] = plane k of this pixel
        if constexpr (kFromScores) {
          // stand-alone stage 3: this pixel's K weights come from global memory (plane-contiguous, coalesced
          // across lanes for [N,K,H,W]); the two warps of a quarter split the planes
          const OT* sc = reinterpret_cast<const OT*>(p.scores) + (size_t)n * p.sn + (size_t)(live ? pix : 0) * p.sp;
          // up to kLd planes in flight per lane: the loads are the tile's latency chain (one round trip for the
          // 16-17 planes a warp owns at K = 33, two at K = 65)
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < kLd; ++j)
              if (k_lo + j < k_hi) my[k_lo + j] = (float)Cvt<OT>::to(pre[j]);
          }
          for (int k = k_lo + (t == 0 ? kLd : 0); k < k_hi; k += kLd) {
            OT v[kLd];
#pragma unroll
            for (int j = 0; j < kLd; ++j)
              v[j] = (live && k + j < k_hi) ? __ldg(sc + (size_t)(k + j) * p.sk) : Cvt<OT>::from(0.0f);
#pragma unroll
            for (int j = 0; j < kLd; ++j)
              if (k + j < k_hi) my[k + j] = (float)Cvt<OT>::to(v[j]);
          }
        } else {
          float T = 1.0f;
          const bool wr = comp != nullptr && live && !BS_ABL_NO_COMP_STORE;
          const bool wr_now = wr && half == 0;          // the front range's values are final in the first pass
          OT* const comp_px = comp + pix;                 // this pixel in plane 0; plane k is + k*P
          int m = m_hi;
          // serial head until the top of the range is float4-aligned in the stash (m a multiple of 4)
          for (; m >= m_lo + 1 && (any_general || (m & 3) != 0 || m < m_lo + 4); --m) {
            const float s = any_general ? blob_opacity(coef[m - 1], xf, yf) : blob_opacity_pd(coef[m - 1], xf, yf);
            const float d = s * T;
            T = fmaf(-s, T, T);
            my[m] = d;
            if (wr_now) __stcs(comp_px + (size_t)m * P, Cvt<OT>::from(d));
          }
          // branch-free groups of 8 blobs (then at most one of 4): planes m-7..m are two aligned float4s of the stash
          for (; m >= m_lo + 8; m -= 8) composite_group<OT, 8>(coef, m, xf, yf, T, my, comp_px, (size_t)P, wr_now);
          if (m >= m_lo + 4) { composite_group<OT, 4>(coef, m, xf, yf, T, my, comp_px, (size_t)P, wr_now); m -= 4; }
          for (; m >= m_lo + 1; --m) {               // (fewer than 4 blobs left)
            const float s = any_general ? blob_opacity(coef[m - 1], xf, yf) : blob_opacity_pd(coef[m - 1], xf, yf);
            const float d = s * T;
            T = fmaf(-s, T, T);
            my[m] = d;
            if (wr_now) __stcs(comp_px + (size_t)m * P, Cvt<OT>::from(d));
          }
          if constexpr (kHalves == 2) {
            float* const cr = carry + (tile_it & 1) * kTcTileM;     // double-buffered: the front warp may be a tile ahead
            if (half == 0) cr[px] = T;
            asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          }
          if (half == kHalves - 1) {
            float c = 1.0f;
            if constexpr (kHalves == 2) {
              c = carry[(tile_it & 1) * kTcTileM + px];
              int k = m_hi;
              for (; k >= 1 && ((k & 3) != 0 || k < 4); --k) {       // unaligned top of the range
                const float v = my[k] * c;
                my[k] = v;
                if (wr) __stcs(comp_px + (size_t)k * P, Cvt<OT>::from(v));
              }
              for (; k >= 4; k -= 4) {                                // planes k-3..k: one aligned float4
                float4 v4 = *reinterpret_cast<const float4*>(my + k - 3);
                v4.x *= c; v4.y *= c; v4.z *= c; v4.w *= c;
                *reinterpret_cast<float4*>(my + k - 3) = v4;
                OT* const cp = comp_px + (size_t)k * P;
                store_pair<OT>(cp, cp - (ptrdiff_t)P, v4.w, v4.z, wr);
                store_pair<OT>(cp - (ptrdiff_t)2 * P, cp - (ptrdiff_t)3 * P, v4.y, v4.x, wr);
              }
              for (; k >= 1; --k) {
                const float v = my[k] * c;
                my[k] = v;
                if (wr) __stcs(comp_px + (size_t)k * P, Cvt<OT>::from(v));
              }
            }
            const float bg = T * c;                     // background: alpha 1 * total transmittance
            my[0] = bg;
            if (wr) __stcs(comp_px, Cvt<OT>::from(bg));
          }
        }
        if constexpr (kHalves == 2) {
          if (g_split < 0) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // quarter's stash columns complete
        }
        // this warp's operand k-groups: its own range's (aligned split) or every kHalves-th one
        const int g_lo = g_split < 0 ? half : (half ? 0 : g_split);
        const int g_hi = g_split < 0 ? p.Kp / kKStep : (half ? g_split : p.Kp / kKStep);
        const int g_inc = g_split < 0 ? kHalves : 1;
        if (ctid == 0 && tile_it == 0) TC_STAMP(3);    // first tile's weights in the stash

        const int abuf = a_bufs == 2 ? (tile_it & 1) : 0, ause = a_bufs == 2 ? (tile_it >> 1) : tile_it;   // buffer, its use count
        if (ause > 0) mbar_wait(&bars->a_free[abuf], (ause - 1) & 1);   // the MMAs of the tile that last used this buffer have read it
        tc_fence_after();
        const uint32_t tmem_a = tmem_a0 + (uint32_t)abuf * a_buf_cols;
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        if constexpr (kTf32) {
          for (int g = g_lo; g < g_hi; g += g_inc) {
            uint32_t hi[8], lo[8];
            const float4 wa = *reinterpret_cast<const float4*>(my - kTcKOff + g * 8);      // operand rows 8g .. 8g+7
            const float4 wb = *reinterpret_cast<const float4*>(my - kTcKOff + g * 8 + 4);
            const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float w = wv[j];      // rows of pixels past the image end are never stored: no masking needed
              const float h = rna_tf32(w);
              hi[j] = __float_as_uint(h);
              lo[j] = __float_as_uint(rna_tf32(w - h));
            }
            tmem_st8(tmem_a + lane_addr + g * 8, hi);
            tmem_st8(tmem_a + lane_addr + a_cols + g * 8, lo);
          }
        } else if constexpr (kH2) {
          // 2xFP16: w = w1 + w2 exactly to 2^-24; two k per 32-bit column, 8 columns = 16 consecutive k per operand
          for (int g = g_lo; g < g_hi; g += g_inc) {
            uint32_t p1[8], p2[8];
            const float* row = my - kTcKOff + g * 16;
            const float4 q0 = *reinterpret_cast<const float4*>(row), q1 = *reinterpret_cast<const float4*>(row + 4);
            const float4 q2 = *reinterpret_cast<const float4*>(row + 8), q3 = *reinterpret_cast<const float4*>(row + 12);
            const float wv[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const __half2 h = __floats2half2_rn(wv[2 * j], wv[2 * j + 1]);
              const float2 f = __half22float2(h);
              const __half2 r = __floats2half2_rn(wv[2 * j] - f.x, wv[2 * j + 1] - f.y);
              p1[j] = *reinterpret_cast<const uint32_t*>(&h);
              p2[j] = *reinterpret_cast<const uint32_t*>(&r);
            }
            tmem_st8(tmem_a + lane_addr + g * 8, p1);
            tmem_st8(tmem_a + lane_addr + a_cols + g * 8, p2);
          }
        } else {
          // two k per 32-bit column (low half = even k): 8 columns = 16 consecutive k
          for (int g = g_lo; g < g_hi; g += g_inc) {
            uint32_t pk[8];
            const float* row = my - kTcKOff + g * 16;                                       // operand rows 16g .. 16g+15
            const float4 q0 = *reinterpret_cast<const float4*>(row), q1 = *reinterpret_cast<const float4*>(row + 4);
            const float4 q2 = *reinterpret_cast<const float4*>(row + 8), q3 = *reinterpret_cast<const float4*>(row + 12);
            const float wv[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // one packed convert per column (low half = even k)
              if constexpr (std::is_same<OT, __nv_bfloat16>::value) {
                const __nv_bfloat162 t = __floats2bfloat162_rn(wv[2 * j], wv[2 * j + 1]);
                pk[j] = *reinterpret_cast<const uint32_t*>(&t);
              } else {
                const __half2 t = __floats2half2_rn(wv[2 * j], wv[2 * j + 1]);
                pk[j] = *reinterpret_cast<const uint32_t*>(&t);
              }
            }
            tmem_st8(tmem_a + lane_addr + g * 8, pk);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->a_full[abuf]);
        if (ctid == 0 && tile_it == 0) TC_STAMP(4);    // first A in TMEM
        if constexpr (kHalves == 2) {
          if (g_split < 0) asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // partner has read the stash
        }
      }
    } else if (warp < kTcComputeWarps + 4) {
      // ========================================= epilogue ==========================================
      const int q = warp - kTcComputeWarps;        // TMEM lane quarter
      OT* out = reinterpret_cast<OT*>(p.grid) + ((size_t)n * p.C + c0) * P;
      float inv = 1.0f;                             // kSplit = 2: read once the unit's first accumulator has landed
      if constexpr (kH2 && !kRing) {
        // max|f| of the NEXT unit's features (and of the first one), computed here while this warp would only wait for
        // the unit's first accumulator: the compute warps' operand staging then converts in a single pass
        const int et = (warp - kTcComputeWarps) * 32 + lane;
        for (int which = (unit_it == 0 ? 0 : 1); which < 2; ++which) {
          int n2 = n, c02 = c0, lv2 = kP == -1 ? level : 0;
          if (which == 1) {
            if (gs >= g_end) break;
            if constexpr (kP == -1) { while (gs >= L.tile_start[lv2 + 1]) ++lv2; }
            const RenderTcParams& pn = L.lv[lv2];
            const int ic = (gs - (kP == -1 ? L.tile_start[lv2] : 0)) / pn.tiles_per_image;
            n2 = ic / pn.c_chunks; c02 = (ic - n2 * pn.c_chunks) * pn.c_tile;
          }
          const RenderTcParams& pu = L.lv[lv2];
          float mx = unit_abs_max<FT>(pu, reinterpret_cast<const FT*>(pu.feats) + (size_t)n2 * pu.K * pu.C, c02, et, 128);
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
          if (lane == 0) { bars->umax[(unit_it + which) & 7][q] = mx; mbar_arrive(&bars->s_full[(unit_it + which) & 1]); }
        }
      }
      for (int t = 0; t < ntiles; ++t, ++tile_it) {
        // same lane -> pixel map as stages 1+2
        const int pix = (t_lo + t) * kTcTileM + q * 32 + (kFloatMaps ? ((lane & 15) << 1) + (lane >> 4) : lane);
        const bool live = pix < P;
#pragma unroll 1
        for (int h = 0; h < n_parts; ++h) {
          mbar_wait_drain(&bars->d_full[h], tile_it & 1);
          if (q == 0 && tile_it == 0) TC_STAMP(5 + h);   // first D half ready
          tc_fence_after();
          if constexpr (kH2) { if (t == 0 && h == 0) inv = bars->unit_inv[unit_it & 7]; }   // staged before the unit's first MMA
          const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(h * c_half);
          OT* const o = out + (size_t)(h * c_half) * P + pix;   // this pixel in the half's first channel plane
          const int ch_left = p.C - (c0 + h * c_half);          // valid channels in this half (may exceed c_half)
          bool done = false;
          if constexpr (kFloatMaps) {
            if (p.pair_ok && ch_left >= c_half && (c_half & 31) == 0) {
              // float fast path: thread t owns pixels 2*(t%16), +1 of channels 16*(t/16) + j of each 32-channel block
              // (two 16-lane loads); the next block's loads are in flight while this one is stored
              const int pix2 = (t_lo + t) * kTcTileM + q * 32 + ((lane & 15) << 1);
              const bool live2 = pix2 < P && !BS_ABL_NO_EPI_STORE;
              const uint32_t tb = taddr + (16u << 16);
              float* const o2 = reinterpret_cast<float*>(out) + (size_t)(h * c_half + ((lane >> 4) << 4)) * P + pix2;
              uint32_t ra[32], rb[32];
              tmem_ld_16x32bx2_x16(taddr, ra); tmem_ld_16x32bx2_x16(tb, ra + 16);
              tmem_wait_ld();
              for (int cc = 0; cc < c_half; cc += 64) {
                if (cc + 32 < c_half) { tmem_ld_16x32bx2_x16(taddr + cc + 32, rb); tmem_ld_16x32bx2_x16(tb + cc + 32, rb + 16); }
                if (kH2 && inv != 1.0f) store_pixel_pairs<true>(o2 + (size_t)cc * P, (size_t)P, ra, live2, inv);
                else store_pixel_pairs<false>(o2 + (size_t)cc * P, (size_t)P, ra, live2, inv);
                tmem_wait_ld();
                if (cc + 32 < c_half) {
                  if (cc + 64 < c_half) { tmem_ld_16x32bx2_x16(taddr + cc + 64, ra); tmem_ld_16x32bx2_x16(tb + cc + 64, ra + 16); }
                  if (kH2 && inv != 1.0f) store_pixel_pairs<true>(o2 + (size_t)(cc + 32) * P, (size_t)P, rb, live2, inv);
                  else store_pixel_pairs<false>(o2 + (size_t)(cc + 32) * P, (size_t)P, rb, live2, inv);
                  tmem_wait_ld();
                }
              }
              done = true;
            }
          }
          if (done) {
          } else if (!kFloatMaps && ch_left >= c_half && (c_half & 31) == 0) {
            // 16-bit fast path: whole 32-column chunks, next TMEM load in flight while the current chunk is stored
            uint32_t ra[32], rb[32];
            tmem_ld32(taddr, ra);
            tmem_wait_ld();
            for (int cc = 0; cc < c_half; cc += 64) {
              if (cc + 32 < c_half) tmem_ld32(taddr + cc + 32, rb);
              OT* oc = o + (size_t)cc * P;                       // chunk base; the 32 planes are immediates when kP > 0
#pragma unroll
              for (int j = 0; j < 32; j += 2)
                store_pair<OT>(oc + (size_t)j * P, oc + (size_t)(j + 1) * P, __uint_as_float(ra[j]), __uint_as_float(ra[j + 1]),
                               live && !BS_ABL_NO_EPI_STORE);
              tmem_wait_ld();
              if (cc + 32 < c_half) {
                if (cc + 64 < c_half) tmem_ld32(taddr + cc + 64, ra);
                oc += (size_t)32 * P;
#pragma unroll
                for (int j = 0; j < 32; j += 2)
                  store_pair<OT>(oc + (size_t)j * P, oc + (size_t)(j + 1) * P, __uint_as_float(rb[j]), __uint_as_float(rb[j + 1]),
                                 live && !BS_ABL_NO_EPI_STORE);
                tmem_wait_ld();
              }
            }
          } else {
            for (int cc = 0; cc < c_half; cc += 16) {           // c_half is a multiple of 16
              uint32_t r[16];
              tmem_ld16(taddr + cc, r);
              tmem_wait_ld();
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (live && cc + j < ch_left) __stcs(o + (size_t)(cc + j) * P, Cvt<OT>::from(kH2 ? __uint_as_float(r[j]) * inv : __uint_as_float(r[j])));
            }
          }
          tc_fence_before();
          if (q == 0 && tile_it == 0) TC_STAMP(7 + h);   // first D half drained
          mbar_arrive(&bars->d_empty[h]);
        }
      }
    } else if (warp == kTcMmaWarp) {
      // ========================================= MMA issue =========================================
      if (lane == 0) {
        const int buf = unit_it % nb, rnd = unit_it / nb;      // this unit's slot of the B ring
        const uint32_t idesc = make_idesc(kTf32 ? 2u : (std::is_same<BT, __half>::value ? 0u : 1u), (uint32_t)c_half,
                                          (kSplit == 0 && BS_B_NMAJOR) ? 1u : 0u);
        const uint32_t b_base = smem_u32(b_smem + (size_t)buf * b_stride);
        const uint32_t lbo = (uint32_t)p.c_tile * 16u, sbo = 128u;
        mbar_wait_mma(&bars->b_full[buf], rnd & 1);
        for (int t = 0; t < ntiles; ++t, ++tile_it) {
          const int abuf = a_bufs == 2 ? (tile_it & 1) : 0, ause = a_bufs == 2 ? (tile_it >> 1) : tile_it;
          mbar_wait_mma(&bars->a_full[abuf], ause & 1);
          tc_fence_after();
          const uint32_t tmem_a = tmem_a0 + (uint32_t)abuf * a_buf_cols;
          for (int h = 0; h < n_parts; ++h) {
            if (tile_it > 0) mbar_wait_mma(&bars->d_empty[h], (tile_it - 1) & 1);
            tc_fence_after();
            const uint32_t d_addr = tmem + (uint32_t)(h * c_half);
            uint32_t acc = 0;
            for (int ks = 0; ks < p.Kp / kKStep; ++ks) {
              // descriptor of the [c_half x kKStep] slab: two 16-byte k-chunks, LBO apart
              const uint32_t b_addr = b_base + (uint32_t)(2 * ks) * lbo + (uint32_t)(h * c_half) * 16u;
              const uint64_t b_hi = make_b_desc(b_addr, lbo, sbo);
              const uint32_t a_hi = tmem_a + (uint32_t)(ks * 8);
              if (!BS_ABL_NO_MMA) umma_ts<kTf32>(d_addr, a_hi, b_hi, idesc, acc);
              acc = 1;
              if constexpr (kFloatMaps) {      // + Ahi*Blo + Alo*Bhi (3xTF32) / + A1*B2 + A2*B1 (2xFP16)
                const uint64_t b_lo = make_b_desc(b_addr + (uint32_t)b_bytes, lbo, sbo);
                if (!BS_ABL_NO_MMA) umma_ts<kTf32>(d_addr, a_hi, b_lo, idesc, 1u);
                if (!BS_ABL_NO_MMA) umma_ts<kTf32>(d_addr, a_hi + (uint32_t)a_cols, b_hi, idesc, 1u);
              }
            }
            tc_commit(&bars->d_full[h]);
          }
          tc_commit(&bars->a_free[abuf]);
        }
        tc_commit(&bars->b_free[buf]);
      }
      __syncwarp();
    } else if constexpr (kRing) {
      // ====================================== operand staging ======================================
      // 3 warps that run ahead of the units: B of unit u goes into ring slot u % nb as soon as the MMAs of unit u - nb
      // have released it, so staging overlaps the tiles of the units before it
      const int buf = unit_it % nb, rnd = unit_it / nb;
      if (rnd > 0) mbar_wait(&bars->b_free[buf], (rnd - 1) & 1);
      tc_stage_b<FT, OT, kSplit>(p, n, c0, b_smem + (size_t)buf * b_stride, b_bytes, (int)threadIdx.x - (kTcMmaWarp + 1) * 32,
                                 kTcStageWarps * 32, bars, unit_it, 7);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // generic-proxy writes -> tensor-core reads
      mbar_arrive(&bars->b_full[buf]);
    }
  }

  // the last commits arrive asynchronously: see them land before the CTA (and its smem barriers) goes away
  if (warp == kTcMmaWarp && lane == 0 && tile_it > 0) {
    {
      const int last = tile_it - 1;
      const int abuf = a_bufs == 2 ? (last & 1) : 0, ause = a_bufs == 2 ? (last >> 1) : last;
      mbar_wait(&bars->a_free[abuf], ause & 1);
    }
    mbar_wait(&bars->b_free[(unit_it - 1) % nb], ((unit_it - 1) / nb) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TC_STAMP(9);
  if (warp == kTcMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)p0.tmem_cols) : "memory");
  }
}


// ---- host side ---------------------------------------------------------------------------------------
static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

struct TcPlan { int Kp, c_tile, nb, tmem_cols, n_parts, a_bufs; size_t smem, b_slot; bool ok; const char* why; };   // smem = fixed part + nb * b_slot

// float32 maps: which split-precision form (header of this file).  BLOBSPLAT_F32_SPLIT=tf32 selects 3xTF32 (A/B knob).
static inline int f32_split() {
  const char* e = getenv("BLOBSPLAT_F32_SPLIT");
  return (e && e[0] == 't') ? 1 : 2;
}
static inline int split_of(int dtype) { return dtype == BLOBSPLAT_F32 ? f32_split() : 0; }
// ... and per launch.  A launch with fewer tiles than SMs is all head (one tile per CTA), and the 2xFP16 form's head is
// longer (per-unit feature maximum before the scaled two-way split): with BS_SMALL_TF32=1 such launches take 3xTF32
// (cfg2, 1 image x 16 blobs x 320 channels: 12.4 -> 10.3 us as a CUDA graph).  OFF by default: the low bits of an image's
// result would then depend on how many images share its launch (a chunked render would differ from the whole batch in the
// seventh digit).  BLOBSPLAT_F32_SPLIT=tf32|fp16 forces one form for every launch.
#ifndef BS_SMALL_TF32
#define BS_SMALL_TF32 0
#endif
static inline int split_for_launch(int dtype, long long tiles) {
  if (dtype != BLOBSPLAT_F32) return 0;
  const char* e = getenv("BLOBSPLAT_F32_SPLIT");
  if (e && e[0] == 't') return 1;
  if (e && e[0] == 'f') return 2;
  static thread_local int cached_dev = -1, sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
    cached_dev = dev;
  return (BS_SMALL_TF32 && tiles > 0 && tiles < sms) ? 1 : 2;
}

static inline int tmem_cols_for(int needed) {
  int c = 32;
  while (c < needed) c <<= 1;
  return c;
}

// Small launches (N * tiles per image below the SM count, e.g. one 64 x 64 image = 32 tiles): a narrower channel tile makes
// more (image, chunk, tile) units, so every SM drains a share of the grid instead of 32 SMs draining all of it; stages 1+2
// are recomputed per chunk (cheap at these sizes), the composed maps are still written by chunk 0 only.
// `tiles` = N * tiles per image (0 = do not narrow).  BS_SPREAD_SMALL=0 turns it off.
#ifndef BS_SPREAD_SMALL
#define BS_SPREAD_SMALL 1
#endif
static inline int spread_c_tile(int c_tile, int C, long long tiles, int floor_c) {
  if (!BS_SPREAD_SMALL || tiles <= 0) return c_tile;
  static thread_local int cached_dev = -1, sms = 148;
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev && cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
    cached_dev = dev;
  if (tiles * ((C + c_tile - 1) / c_tile) >= sms) return c_tile;
  const long long want = (sms + tiles - 1) / tiles;                      // channel chunks for one unit per SM
  int c = (int)((C + want - 1) / want);
  c = std::max(floor_c, (c + 31) / 32 * 32);
  return std::min(c_tile, c);
}

static inline TcPlan plan_tc(int K, int C, int split, long long tiles = 0) {
  TcPlan pl{};
  pl.ok = false;
  const bool tf32 = split == 1;
  const int kstep = tf32 ? 8 : 16;
  pl.Kp = round_up(K + kTcKOff, kstep);
  if (K - 1 > kTcMaxBlobs) { pl.why = "more than 127 blobs: use the FMA engine"; return pl; }
  if (C < 1) { pl.why = "no channels"; return pl; }
  const int a_cols = tf32 ? 2 * pl.Kp : (split == 2 ? pl.Kp : pl.Kp / 2);
  const size_t per_c = (size_t)pl.Kp * (tf32 ? 8 : (split == 2 ? 4 : 2));   // B bytes per channel (hi+lo fp32 | x1+x2 fp16 | 16-bit)
  const size_t fixed = (size_t)(pl.Kp + 4) * kTcTileM * 4 + 2 * kTcTileM * 4 + (kTcMaxBlobs + 1) * sizeof(BlobCoef) + sizeof(TcBarriers) + 512;
  int c_tile = std::min(kTcMaxCTile, round_up(C, 32));     // any C: the last chunk may be ragged (zero B columns, predicated drain)
  c_tile = std::min(c_tile, (512 - a_cols) / 32 * 32);
  c_tile = std::min<long long>(c_tile, (long long)((kTcSmemBudget - fixed) / per_c) / 32 * 32);
  if (c_tile < 32) { pl.why = "K too large for shared/tensor memory"; return pl; }
  // prefer a tile that divides C (no ragged chunk)
  for (int c = c_tile; c >= std::max(32, c_tile / 2); c -= 32)
    if (C % c == 0) { c_tile = c; break; }
  c_tile = spread_c_tile(c_tile, C, tiles, 64);
  pl.c_tile = c_tile;
  // as many B buffers as fit (ring, staged ahead by the staging warps); one when B fills shared memory
  pl.nb = (int)std::min<size_t>(BS_MAX_B, (kTcSmemBudget - fixed) / (per_c * c_tile));
  pl.b_slot = per_c * c_tile;
  pl.smem = fixed + (size_t)pl.nb * pl.b_slot;
  pl.a_bufs = (BS_A_BUFS >= 2 && c_tile + 2 * a_cols <= 512) ? 2 : 1;
  pl.tmem_cols = tmem_cols_for(c_tile + pl.a_bufs * a_cols);
  pl.n_parts = (BS_D_PARTS && c_tile % 64 == 0 && c_tile / 64 >= 2 && c_tile / 64 <= kTcMaxParts) ? c_tile / 64 : 2;
  pl.ok = true;
  return pl;
}

// Fill the shape / work-unit fields shared by both A sources.
static int fill_tc_units(RenderTcParams& p, const TcPlan& pl, int N, int K, int H, int W, int C, int tile_px = kTcTileM) {
  p.N = N; p.M = K - 1; p.H = H; p.W = W; p.C = C; p.K = K; p.Kp = pl.Kp;
  p.c_tile = pl.c_tile; p.c_chunks = (C + pl.c_tile - 1) / pl.c_tile;
  p.tmem_cols = pl.tmem_cols; p.n_parts = pl.n_parts; p.a_bufs = pl.a_bufs;
  const int P = H * W;
  p.tiles_per_image = (P + tile_px - 1) / tile_px;
  const long long total = (long long)N * p.c_chunks * p.tiles_per_image;
  if (total > 0x7fffffffll) BS_UNSUPPORTED("too many tiles for one launch");
  p.total_tiles = (int)total;
  // pick the cheaper partition under the cost model: a tile = 1, staging a unit's operands = BS_STAGE_COST tiles
  static thread_local int cached_dev = -1, cached_sms = 148;   // host-side launch cost matters for the small launches
  int dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess && dev != cached_dev &&
      cudaDeviceGetAttribute(&cached_sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess)
    cached_dev = dev;
  const int sms = cached_sms;
  const int ctas = std::min(sms, p.total_tiles);
  const int tw = tile_px / kTcTileM;                           // cost of a tile in 128-pixel tiles
  long long ranges = 0;                                        // worst CTA under equal tile ranges
  int range_units = 0;
  for (int i = 0; i < ctas; ++i) {
    const int lo = tc_range_begin(p.total_tiles, i, ctas), hi = tc_range_begin(p.total_tiles, i + 1, ctas);
    if (hi <= lo) continue;
    const int units = (hi - 1) / p.tiles_per_image - lo / p.tiles_per_image + 1;
    range_units = std::max(range_units, units);
    ranges = std::max<long long>(ranges, (long long)(hi - lo) * tw + (long long)BS_STAGE_COST * units);   // weights fitted to profiles/schedule_ab_r2.txt
  }
  const long long runs = total / p.tiles_per_image;
  const long long whole_units = ctas > 0 ? (runs + ctas - 1) / ctas : 0;
  const long long whole = whole_units * ((long long)p.tiles_per_image * tw + BS_STAGE_COST);
  p.whole_runs = whole <= ranges ? 1 : 0;
  if (const char* e = getenv("BLOBSPLAT_TC_SCHEDULE")) {      // A/B knob (read per call): "whole" | "ranges"
    if (e[0] == 'w') p.whole_runs = 1; else if (e[0] == 'r') p.whole_runs = 0;
  }
  // B ring + staging warps only where a CTA has enough units for staging to run ahead of; with one or two units per
  // CTA the 256 compute threads stage faster than the 96 staging threads and there is nothing to overlap with
  const long long units_per_cta = p.whole_runs ? whole_units : range_units;
  p.nb = (pl.nb >= 2 && units_per_cta >= BS_RING_MIN_UNITS) ? pl.nb : 1;
  p.smem_bytes = (int)(pl.smem - (size_t)(pl.nb - p.nb) * pl.b_slot);
  return 0;
}

template <typename FT, typename OT, int kSplit, int kHalves, int kP, bool kFromScores, bool kRing>
static int launch_tc_pr(const RenderTcParams& p, cudaStream_t st) {
  static thread_local int configured_dev = -1;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(render_tc_kernel<FT, OT, kSplit, kHalves, kP, kFromScores, kRing>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    configured_dev = dev;
  }
  static thread_local int sm_count = 0, sm_dev = -1;
  if (sm_dev != dev) {
    BS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    sm_dev = dev;
  }
  const int grid = std::min(sm_count, p.total_tiles);
  RenderTcLevels L{};
  L.lv[0] = p;
  L.lv[0].pair_ok = ((p.H * p.W) & 1) == 0 && (reinterpret_cast<uintptr_t>(p.grid) & 7) == 0;
  L.n_levels = 1;
  L.tile_start[1] = p.total_tiles;
  BS_CUDA(launch_pdl(render_tc_kernel<FT, OT, kSplit, kHalves, kP, kFromScores, kRing>, dim3(grid),
                     dim3((4 * kHalves + 5 + (kRing ? kTcStageWarps : 0)) * 32), (size_t)p.smem_bytes, st, L));
  return 0;
}

template <typename FT, typename OT, int kSplit, int kHalves, int kP, bool kFromScores>
static int launch_tc_p(const RenderTcParams& p, size_t, cudaStream_t st) {
  if constexpr (kHalves == 2) {
    if (p.nb > 1) return launch_tc_pr<FT, OT, kSplit, kHalves, kP, kFromScores, true>(p, st);
  }
  return launch_tc_pr<FT, OT, kSplit, kHalves, kP, kFromScores, false>(p, st);
}

template <typename FT, typename OT, int kSplit, int kHalves, bool kFromScores>
static int launch_tc(const RenderTcParams& p, size_t smem, cudaStream_t st) {
  if constexpr (kHalves == 2) {   // plane-stride specialisations for BlobNet's latent resolutions (64/32/16)
    switch (p.H * p.W) {
      case 64: if constexpr (kFromScores) return launch_tc_p<FT, OT, kSplit, kHalves, 64, kFromScores>(p, smem, st); else break;
      case 4096: return launch_tc_p<FT, OT, kSplit, kHalves, 4096, kFromScores>(p, smem, st);
      case 1024: return launch_tc_p<FT, OT, kSplit, kHalves, 1024, kFromScores>(p, smem, st);
      case 256: return launch_tc_p<FT, OT, kSplit, kHalves, 256, kFromScores>(p, smem, st);
    }
  }
  return launch_tc_p<FT, OT, kSplit, kHalves, 0, kFromScores>(p, smem, st);
}

// 8 compute warps (kHalves = 2) win for every dtype once the GPU settles at its sustained clocks
// (profiles/ab_compute_warps_r1.txt); BLOBSPLAT_TC_HALVES=1|2 overrides the choice (A/B knob, read per call).
template <bool kFromScores>
static int launch_tc_dtype(const RenderTcParams& p, size_t smem, int out_dtype, cudaStream_t st) {
  int halves = 2;
  if (!kFromScores) {
    if (const char* e = getenv("BLOBSPLAT_TC_HALVES")) { if (e[0] == '1') halves = 1; }
  }
  if (out_dtype == BLOBSPLAT_F32) {
    if (p.f32_split == 1) return launch_tc<float, float, 1, 2, kFromScores>(p, smem, st);             // 3xTF32 (small launches, A/B partner)
    if constexpr (!kFromScores) { if (halves == 1) return launch_tc<float, float, 2, 1, kFromScores>(p, smem, st); }
    return launch_tc<float, float, 2, 2, kFromScores>(p, smem, st);
  }
  if (out_dtype == BLOBSPLAT_BF16) {
    if constexpr (!kFromScores) { if (halves == 1) return launch_tc<__nv_bfloat16, __nv_bfloat16, 0, 1, kFromScores>(p, smem, st); }
    return launch_tc<__nv_bfloat16, __nv_bfloat16, 0, 2, kFromScores>(p, smem, st);
  }
  if (out_dtype == BLOBSPLAT_F16) {
    if constexpr (!kFromScores) { if (halves == 1) return launch_tc<__half, __half, 0, 1, kFromScores>(p, smem, st); }
    return launch_tc<__half, __half, 0, 2, kFromScores>(p, smem, st);
  }
  BS_UNSUPPORTED("tensor-core engine: unsupported dtype %d", out_dtype);
}

}  // namespace blobsplat
