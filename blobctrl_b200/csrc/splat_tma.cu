// Stage 3 from 16-bit score maps with no thread on the operand path — TMA in, tcgen05 from shared memory, full-line stores out.
//
// Replaces splat_features_from_scores (blobctrl/utils/utils.py:57-77; duplicate at
// blobctrl/pipelines/pipeline_blobnet.py:706-721) for bf16 / f16 maps, one level or a whole pyramid
// (utils.py:226-233: one splat per pyramid level) in ONE launch.  The contraction is taken transposed,
//
//   D[channel, pixel] = sum_k F[k, channel] * S[k, pixel]         (M = 128 channels, N <= 256 pixels, K = blobs + 1)
//
// because that is how both operands already lie in HBM:
//   A = features [K, C], channel-contiguous  = an MN-major operand,  B = scores [K, P], pixel-contiguous = an MN-major
//   operand.  A TMA box of (64 elements = 128 contiguous bytes) x (Kp rows) with the 128-byte swizzle lands as one
//   MN-major SWIZZLE_128B block of the UMMA operand (rows of 128 B, 8-row atoms of 1 KB; SBO = 1 KB, LBO = the block
//   stride): no thread touches an operand, rows k >= K, channels >= C and pixels >= P arrive as zeros (out-of-bounds
//   fill).  [A first version used 16-byte boxes into the no-swizzle layout: correct, but 8x the TMA requests.]
//   D (fp32, TMEM): lane = channel, column = pixel.  A drain thread therefore holds CONSECUTIVE PIXELS of one channel
//   plane: it packs them to 16 bits and writes 16-byte chunks into a 128B-swizzled staging box of 32 channels x 64 pixels
//   (conflict-free); the warp reads the box back row-wise and stores whole 128-byte lines, four per instruction
//   (st.global.cs.v4).  Handing the same box to a TMA tensor store instead (BLOBSPLAT_ST_STORE=tma, kept as an A/B path)
//   is slower: the TMA unit needs ~6 cycles per 128-byte row (profiles/tma_store_bw_r2.txt), and alternating the two paths
//   (=hybrid) is slower than either (profiles/splat_tma_r2.md).
//
// Warp roles (320 threads, 1 CTA/SM, persistent over a cost-balanced contiguous range of items; an item = (level, image,
// pixel tile, 128-channel group), channel group fastest so a tile's scores stay resident):
//   warp 0      producer: one thread issues the TMA loads (scores ring of up to 3, feature ring of up to 8)
//   warp 1      TMEM allocation + single-thread tcgen05.mma issue (SS form) + commits
//   warps 2-9   drain: two groups of four (TMEM lane quarter = warp & 3), each owning one accumulator slot and every
//               other item; private staging box per warp
#include <cuda.h>

#include <cstring>

#include "render_tc.cuh"

#ifndef BS_ST_STORE_MODE
#define BS_ST_STORE_MODE 0     // drain store path: 0 = line stores by the warps, 1 = TMA tensor stores, 2 = alternate (BLOBSPLAT_ST_STORE)
#endif
#ifndef BS_ST_OVERHEAD
#define BS_ST_OVERHEAD 128     // fixed cost of an item in the schedule's cost model, in pixels (swept on cfg3's lower levels:
                               // 0 -> 51 us, 48 -> 37, 96..256 -> 33, 700 -> 37; profiles/splat_tma_r2.md)
#endif

namespace blobsplat {

#define ST_STAMP(role, i, k) do { if constexpr (kDbg) { if (k_dbg && blockIdx.x == p.dbg_cta && (i) < 64) { unsigned long long t_; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_)); k_dbg[((role) * 64 + (i)) * 4 + (k)] = t_; } } } while (0)

constexpr int kStMaxLevels = 4;
constexpr int kStM = 128;                 // channels per item (MMA M)
constexpr int kStBoxPx = 64, kStBoxCh = 32;
constexpr int kStBlk = 64;                // elements of one 128-byte swizzled operand block
constexpr int kStBoxBytes = kStBoxPx * kStBoxCh * 2;
constexpr int kStDrainWarps = 8, kStFirstDrainWarp = 2, kStThreads = (kStFirstDrainWarp + kStDrainWarps) * 32;
constexpr int kStMaxF = 8, kStMaxS = 4;

struct alignas(64) StLevel {
  CUtensorMap feats, scores, out;
  int npx;          // pixels per tile = MMA N (multiple of 16, <= 256)
  int tiles;        // tiles per image
  int groups;       // 128-channel groups
  int P;            // pixels per image
  int item_start;   // first linear item of this level
  int cost;         // schedule cost of one item
  long long cost_start;
};
struct alignas(64) StParams {
  StLevel lv[kStMaxLevels];
  int n_levels, n_items, N;
  long long total_cost;
  int Kp;           // K rounded up to 16
  int nf, ns, nbuf; // feature ring depth, score ring depth, staging boxes per drain warp
  int tma_store;    // 0: the drain reads its boxes back row-wise and stores 128-byte lines itself; 1: hands them to the TMA unit; 2: alternates
  const void* out_ptr[kStMaxLevels]; int out_C[kStMaxLevels];   // for the direct stores
  int s_bytes, f_bytes;
  int slot_cols;    // accumulator columns per slot (power of two >= the widest tile)
  int is_bf16;
  int dbg_cta;
  unsigned long long* dbg;   // measurement only (BLOBSPLAT_ST_DBG_PTR): clock64 stamps of CTA 0, [role][item][4]
  int abl;          // measurement only (BLOBSPLAT_ST_ABL): 1 = no TMA stores, 2 = no MMAs, 4 = feature loads only for the first ring pass, 8 = no packing / staging writes
};

struct StBarriers {
  uint64_t s_full[kStMaxS], s_free[kStMaxS], f_full[kStMaxF], f_free[kStMaxF], d_full[2], d_empty[2];
  uint32_t tmem_base;
};

// first item of CTA i when the sequence is cut into `ctas` contiguous ranges of equal cost
__device__ __forceinline__ int st_range_begin(const StParams& p, int i, int ctas) {
  const long long target = p.total_cost * i / ctas;
  int l = 0;
  while (l + 1 < p.n_levels && target >= p.lv[l + 1].cost_start) ++l;
  const StLevel& L = p.lv[l];
  const long long it = (target - L.cost_start + L.cost - 1) / L.cost;
  const int next = (l + 1 < p.n_levels) ? p.lv[l + 1].item_start : p.n_items;
  return (int)min((long long)next, (long long)L.item_start + it);
}

struct StItem { int level, n, tile, group, key; };
__device__ __forceinline__ StItem st_decode(const StParams& p, int item, int& level) {
  while (level + 1 < p.n_levels && item >= p.lv[level + 1].item_start) ++level;
  const StLevel& L = p.lv[level];
  const int j = item - L.item_start;
  StItem it;
  it.level = level;
  it.key = j / L.groups;                    // (image, tile) within the level
  it.group = j - it.key * L.groups;
  it.n = it.key / L.tiles;
  it.tile = it.key - it.n * L.tiles;
  it.key = it.key * kStMaxLevels + level;   // unique across levels
  return it;
}

// Walks the item sequence (group fastest, then tile, image, level) without per-item divisions: the producer and MMA
// roles are single threads, where a decode with three integer divisions costs several hundred cycles per item.
struct StIter {
  int level, n, tile, group;
  __device__ __forceinline__ void init(const StParams& p, int item) {
    int lv = 0;
    const StItem it = st_decode(p, item, lv);
    level = it.level; n = it.n; tile = it.tile; group = it.group;
  }
  // next item; true when it belongs to another (level, image, tile) — a new score tile
  __device__ __forceinline__ bool advance(const StParams& p) {
    const StLevel& L = p.lv[level];
    if (++group < L.groups) return false;
    group = 0;
    if (++tile < L.tiles) return true;
    tile = 0;
    if (++n < p.N) return true;
    n = 0; ++level;
    return true;
  }
};

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"(dst), "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar)) : "memory");
}
// MN-major SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start >> 4 | LBO >> 4 @16 (stride between
// 64-element MN blocks) | SBO >> 4 @32 (stride between 8-row k groups = 1 KB) | version 1 @46 | layout SWIZZLE_128B (2) @61
__device__ __forceinline__ uint64_t st_desc(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
         (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* tm, uint32_t src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.tile.bulk_group [%0, {%2, %3, %4}], [%1];"
               ::"l"(tm), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// cute::UMMA::InstrDescriptor: c_format F32 @4, a/b format @7/@10 (0 = f16, 1 = bf16), a_major @15 and b_major @16
// (1 = MN-major), N >> 3 @17, M >> 4 @24
__device__ __forceinline__ uint32_t st_idesc(uint32_t bf16, uint32_t n) {
  return (1u << 4) | (bf16 << 7) | (bf16 << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((uint32_t)(kStM >> 4) << 24);
}

template <bool kBf16>
__device__ __forceinline__ uint32_t st_pack(uint32_t a, uint32_t b) {        // low half = the lower pixel
  if constexpr (kBf16) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(__uint_as_float(a), __uint_as_float(b));
    return *reinterpret_cast<const uint32_t*>(&t);
  } else {
    const __half2 t = __floats2half2_rn(__uint_as_float(a), __uint_as_float(b));
    return *reinterpret_cast<const uint32_t*>(&t);
  }
}

// kDbg = false is the production kernel: line stores, no ablation switches, no timeline stamps — those knobs are compiled
// out, because the kernel sits at its 168-register cap and every live knob costs the drain loop registers.  Any of
// BLOBSPLAT_ST_STORE / BLOBSPLAT_ST_ABL / BLOBSPLAT_ST_DBG_PTR selects the kDbg = true instantiation.
template <bool kBf16, bool kDbg>
__global__ void __launch_bounds__(kStThreads, 1)
splat_tma_kernel(const __grid_constant__ StParams p) {
  const int k_abl = kDbg ? p.abl : 0, k_tma_store = kDbg ? p.tma_store : 0;
  unsigned long long* const k_dbg = kDbg ? p.dbg : nullptr;
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* const stage = smem;                                                     // [8 warps][nbuf][4 KB], 1 KB aligned
  unsigned char* const s_ring = stage + (size_t)kStDrainWarps * p.nbuf * kStBoxBytes;    // [2][s_bytes]
  unsigned char* const f_ring = s_ring + (size_t)p.ns * p.s_bytes;                    // [nf][f_bytes]
  StBarriers* const bars = reinterpret_cast<StBarriers*>(f_ring + (size_t)p.nf * p.f_bytes);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t tmem_cols = (uint32_t)(2 * p.slot_cols);
  if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < kStMaxS; ++i) { mbar_init(&bars->s_full[i], 1); mbar_init(&bars->s_free[i], 1); }
      for (int i = 0; i < kStMaxF; ++i) { mbar_init(&bars->f_full[i], 1); mbar_init(&bars->f_free[i], 1); }
      for (int i = 0; i < 2; ++i) { mbar_init(&bars->d_full[i], 1); mbar_init(&bars->d_empty[i], kStDrainWarps / 2); }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();                                            // the score maps may come from the previous kernel in the stream

  if (k_dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); k_dbg[1024 + 2 * blockIdx.x] = t; }
  const int it_begin = st_range_begin(p, (int)blockIdx.x, (int)gridDim.x);
  const int it_end = st_range_begin(p, (int)blockIdx.x + 1, (int)gridDim.x);
  const int ksteps = p.Kp >> 4;

  if (warp == 0) {
    // ================================ producer: TMA loads ================================
    if (lane == 0 && it_begin < it_end) {
      StIter it;
      it.init(p, it_begin);
      bool new_tile = true;
      int sb = 0, s_pass = 0, fs = 0, f_pass = 0;
      for (int item = it_begin, i = 0; item < it_end; ++item, ++i) {
        const StLevel& L = p.lv[it.level];
        if (new_tile) {
          if (s_pass > 0) mbar_wait_spin(&bars->s_free[sb], (uint32_t)((s_pass - 1) & 1));
          const int nblk = (L.npx + kStBlk - 1) / kStBlk;      // 64-pixel blocks of the tile, each Kp rows x 128 B
          mbar_expect_tx(&bars->s_full[sb], (uint32_t)(nblk * p.Kp * 128));
          const uint32_t dst = smem_u32(s_ring + (size_t)sb * p.s_bytes);
          for (int b = 0; b < nblk; ++b)
            tma_load_3d(dst + (uint32_t)(b * p.Kp * 128), &L.scores, it.tile * L.npx + b * kStBlk, 0, it.n, &bars->s_full[sb]);
          if (++sb == p.ns) { sb = 0; ++s_pass; }
        }
        ST_STAMP(0, i, 0);
        if (f_pass > 0) mbar_wait_spin(&bars->f_free[fs], (uint32_t)((f_pass - 1) & 1));
        ST_STAMP(0, i, 1);
        if ((k_abl & 4) && f_pass > 0) mbar_arrive(&bars->f_full[fs]);
        else {
          mbar_expect_tx(&bars->f_full[fs], (uint32_t)(p.Kp * kStM * 2));
          const uint32_t dst = smem_u32(f_ring + (size_t)fs * p.f_bytes);
          for (int b = 0; b < kStM / kStBlk; ++b)
            tma_load_3d(dst + (uint32_t)(b * p.Kp * 128), &L.feats, it.group * kStM + b * kStBlk, 0, it.n, &bars->f_full[fs]);
        }
        ST_STAMP(0, i, 2);
        if (++fs == p.nf) { fs = 0; ++f_pass; }
        new_tile = it.advance(p);
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issue ================================
    if (lane == 0 && it_begin < it_end) {
      StIter it;
      it.init(p, it_begin);
      bool new_tile = true;
      int sb = -1, s_next = 0, s_pass = 0, fs = 0, f_pass = 0;
      // descriptor bits that never change: LBO (block stride) | SBO = 1 KB | version 1 | SWIZZLE_128B
      const uint64_t desc_hi = ((uint64_t)((uint32_t)(p.Kp * 128) >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
      const uint32_t f_base = smem_u32(f_ring) >> 4, s_base = smem_u32(s_ring) >> 4;
      for (int item = it_begin, item_it = 0; item < it_end; ++item, ++item_it) {
        ST_STAMP(1, item_it, 0);
        const StLevel& L = p.lv[it.level];
        if (new_tile) {
          if (sb >= 0) tc_commit(&bars->s_free[sb]);              // every MMA that read the previous tile's scores is done
          sb = s_next;
          mbar_wait_spin(&bars->s_full[sb], (uint32_t)(s_pass & 1));
          if (++s_next == p.ns) { s_next = 0; ++s_pass; }
        }
        mbar_wait_spin(&bars->f_full[fs], (uint32_t)(f_pass & 1));
        ST_STAMP(1, item_it, 1);
        const int slot = item_it & 1;
        if (item_it >= 2) mbar_wait_spin(&bars->d_empty[slot], (uint32_t)(((item_it >> 1) - 1) & 1));
        ST_STAMP(1, item_it, 2);
        tc_fence_after();
        const uint32_t idesc = st_idesc(kBf16 ? 1u : 0u, (uint32_t)L.npx);
        const uint64_t a_desc = desc_hi | (uint64_t)(f_base + (uint32_t)((fs * p.f_bytes) >> 4));
        const uint64_t b_desc = desc_hi | (uint64_t)(s_base + (uint32_t)((sb * p.s_bytes) >> 4));
        const uint32_t d_addr = tmem + (uint32_t)(slot * p.slot_cols);
        for (int ks = 0; ks < ((k_abl & 2) ? 0 : ksteps); ++ks)    // 16 k rows = two 1 KB atoms (128 x 16 B) further into every block
          umma_ss(d_addr, a_desc + (uint64_t)(ks * 128), b_desc + (uint64_t)(ks * 128), idesc, ks > 0 ? 1u : 0u);
        tc_commit(&bars->f_free[fs]);
        tc_commit(&bars->d_full[slot]);
        ST_STAMP(1, item_it, 3);
        if (++fs == p.nf) { fs = 0; ++f_pass; }
        new_tile = it.advance(p);
      }
      // nothing reads s_free / f_free after the last item; the drain's d_full wait orders the kernel's end
    }
    __syncwarp();
  } else if (warp >= kStFirstDrainWarp) {
    // ================================ drain: TMEM -> 16-bit -> swizzled box -> TMA store ================================
    // Two groups of four warps (one warp per TMEM lane quarter); group g owns accumulator slot g and takes the items
    // with item_it & 1 == g, so one group's TMEM reads and packing overlap the other group's stores: the SM's store
    // stream never pauses (in lock-step all eight warps idled ~25 % of an item waiting for the next accumulator).
    const int dw = warp - kStFirstDrainWarp, q = warp & 3, grp = dw >> 2;
    unsigned char* const my_stage = stage + (size_t)dw * p.nbuf * kStBoxBytes;
    const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
    int st_it = 0, t_it = 0;
    StIter it;
    if (it_begin + grp < it_end) it.init(p, it_begin + grp);
    for (int item = it_begin + grp, item_it = grp; item < it_end; item += 2, item_it += 2) {
      const StLevel& L = p.lv[it.level];
      const int use = item_it >> 1;                              // this group's uses of its slot so far
      if (lane == 0 && q == 0) ST_STAMP(2 + grp, use, 0);
      mbar_wait(&bars->d_full[grp], (uint32_t)(use & 1));
      tc_fence_after();
      if (lane == 0 && q == 0) ST_STAMP(2 + grp, use, 1);
      const int px_tile = it.tile * L.npx;                       // first pixel of the tile
      const int live = min(L.npx, L.P - px_tile);                // valid pixels of the tile
      const int boxes = (live + kStBoxPx - 1) / kStBoxPx;        // 64-pixel boxes, in rounds of two (register budget)
      for (int b_lo = 0; b_lo < boxes; b_lo += 2) {
      const int b_hi = min(boxes, b_lo + 2);
      uint32_t pk[2][kStBoxPx / 2];
      const uint32_t t0 = tmem + lane_addr + (uint32_t)(grp * p.slot_cols + b_lo * kStBoxPx);
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        if (b_lo + bb < b_hi) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            uint32_t r[32];
            tmem_ld32(t0 + bb * kStBoxPx + h * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 16; ++j) pk[bb][h * 16 + j] = st_pack<kBf16>(r[2 * j], r[2 * j + 1]);
          }
        }
      }
      if (b_hi == boxes) {                                       // last round: the accumulator slot goes back to the MMA
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->d_empty[grp]);
        if (lane == 0 && q == 0) ST_STAMP(2 + grp, use, 2);
      }
#pragma unroll
      for (int bb = 0; bb < 2; ++bb) {
        const int b = b_lo + bb;
        if (b >= b_hi) break;
        // store path of this box: 0 = line stores by the warp, 1 = TMA tensor store; mode 2 (hybrid) alternates, so the LSU
        // and the TMA unit each carry half of the bytes (box 0 of the staging ring is the line-store box, 1.. the TMA ring)
        const bool via_tma = k_tma_store == 1 || (k_tma_store == 2 && (st_it & 1));
        const int ring0 = k_tma_store == 2 ? 1 : 0, ring = p.nbuf - ring0;
        const int sbuf = via_tma ? ring0 + t_it % ring : 0;
        if (via_tma && t_it >= ring) {                           // the TMA store that last read this box has drained it
          if (lane == 0) {
            if (ring == 3) asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
            else if (ring == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          __syncwarp();
        }
        unsigned char* const box = my_stage + (size_t)sbuf * kStBoxBytes;
        unsigned char* const row = box + lane * 128;
        if (!(k_abl & 8))
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) =
              make_uint4(pk[bb][4 * c], pk[bb][4 * c + 1], pk[bb][4 * c + 2], pk[bb][4 * c + 3]);
        }
        if (via_tma) {
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          __syncwarp();
          if (lane == 0) {
            if (!(k_abl & 1)) tma_store_3d(&L.out, smem_u32(box), px_tile + b * kStBoxPx, it.group * kStM + q * kStBoxCh, it.n);
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
          ++t_it;
        } else {
          // read the box back row-wise: 8 lanes fetch the 8 chunks of one channel row = one whole 128-byte line of the
          // output plane, a warp instruction stores 4 lines
          __syncwarp();
          const int ch0 = it.group * kStM + q * kStBoxCh, C = p.out_C[it.level];
          const int px = px_tile + b * kStBoxPx + ((lane & 7) << 3);
          unsigned char* const o = reinterpret_cast<unsigned char*>(const_cast<void*>(p.out_ptr[it.level])) +
                                   (((size_t)it.n * C + ch0) * L.P + px) * 2;
          const bool px_ok = px < L.P && !(k_abl & 1);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3);
            const uint4 v = *reinterpret_cast<const uint4*>(box + rr * 128 + (((lane & 7) ^ (rr & 7)) << 4));
            if (px_ok && ch0 + rr < C) __stcs(reinterpret_cast<uint4*>(o + (size_t)rr * L.P * 2), v);
          }
          __syncwarp();                                          // the box is rewritten by the next one
        }
        ++st_it;
      }
      }
      if (lane == 0 && q == 0) ST_STAMP(2 + grp, use, 3);
      it.advance(p);
      if (item + 1 < it_end) it.advance(p);
    }
    if (k_tma_store && lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (k_dbg && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); k_dbg[1025 + 2 * blockIdx.x] = t; }
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(tmem_cols) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// Can these levels run on the TMA kernel?  16-bit maps, pixel-contiguous 16-byte aligned score maps, P and C multiples
// of 8 (16-byte TMA strides).
bool splat_tma_usable(int n_levels, const void* const* scores, const int64_t* sn, const int64_t* sk, const int64_t* sp,
                      const void* const* feats, void* const* outs, int N, int K, const int* C, const int* H, const int* W,
                      int dtype) {
  if (dtype != BLOBSPLAT_BF16 && dtype != BLOBSPLAT_F16) return false;
  if (n_levels < 1 || n_levels > kStMaxLevels || N < 1 || K < 1 || K > 256) return false;   // a TMA box holds <= 256 rows
  if (!encode_tiled_fn()) return false;
  for (int i = 0; i < n_levels; ++i) {
    const long long P = (long long)H[i] * W[i];
    if (P < 8 || (P & 7) != 0 || (C[i] & 7) != 0 || C[i] < 8 || P >= (1ll << 31)) return false;
    if (sp[i] != 1 || (sk[i] & 7) != 0 || sk[i] < P) return false;
    if (N > 1 && ((sn[i] & 7) != 0 || sn[i] < 8)) return false;
    if (((reinterpret_cast<uintptr_t>(scores[i]) | reinterpret_cast<uintptr_t>(feats[i]) | reinterpret_cast<uintptr_t>(outs[i])) & 15) != 0)
      return false;
    if ((long long)N * ((P + 15) / 16) * ((C[i] + kStM - 1) / kStM) > 0x3fffffffll) return false;
  }
  return true;
}

struct StPlan { int npx_max, nf, ns, nbuf; size_t smem; bool ok; };
static StPlan plan_splat_tma(int Kp, int widest, bool tma_store) {
  StPlan pl{};
  for (int npx_max : {256, 128, 64}) {
    if (npx_max > 64 && npx_max / 2 >= widest) continue;          // no level needs a tile that wide
    const size_t s_bytes = (size_t)Kp * round_up(std::min(npx_max, widest), kStBlk) * 2, f_bytes = (size_t)Kp * kStM * 2;
    for (int depth = 5; depth >= 0; --depth) {                    // prefer deep rings, shrink until it fits
      static const int kNf[6] = {2, 2, 3, 4, 6, 8};
      const int nf = kNf[depth], ns = depth >= 3 ? 3 : 2, nbuf = !tma_store ? 1 : (depth >= 1 ? 3 : 2);
      const size_t s = (size_t)kStDrainWarps * nbuf * kStBoxBytes + ns * s_bytes + nf * f_bytes + sizeof(StBarriers) + 1024;
      if (s <= kTcSmemBudget) { pl.npx_max = npx_max; pl.nf = nf; pl.ns = ns; pl.nbuf = nbuf; pl.smem = s; pl.ok = true; return pl; }
    }
  }
  return pl;
}

static int encode(CUtensorMap* tm, CUtensorMapDataType dt, int rank, const void* base, const cuuint64_t* dims,
                  const cuuint64_t* strides, const cuuint32_t* box, CUtensorMapSwizzle sw) {
  const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
  const CUresult r = encode_tiled_fn()(tm, dt, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, ones,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) BS_UNSUPPORTED("cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

// One launch for n_levels stage-3 problems of one batch (same N, K, dtype).  Preconditions: splat_tma_usable.
int splat_tma_dispatch(int n_levels, const void* const* scores, const int64_t* sn, const int64_t* sk, const void* const* feats,
                       void* const* outs, int N, int K, const int* C, const int* H, const int* W, int dtype, cudaStream_t st) {
  static thread_local int configured_dev = -1, sm_count = 0;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(splat_tma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaFuncSetAttribute(splat_tma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaFuncSetAttribute(splat_tma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaFuncSetAttribute(splat_tma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    configured_dev = dev;
  }
  StParams p;
  std::memset(&p, 0, sizeof(p));
  p.Kp = round_up(K, 16);
  int widest = 16;
  for (int i = 0; i < n_levels; ++i) widest = std::max(widest, (int)std::min<long long>(256, round_up(H[i] * W[i], 16)));
  const char* sm = std::getenv("BLOBSPLAT_ST_STORE");
  const int store_mode = !sm ? BS_ST_STORE_MODE : (std::strcmp(sm, "tma") == 0 ? 1 : (std::strcmp(sm, "hybrid") == 0 ? 2 : 0));   // A/B knob
  const bool tma_store = store_mode != 0;
  const StPlan pl = plan_splat_tma(p.Kp, widest, tma_store);
  if (!pl.ok) BS_UNSUPPORTED("TMA feature splat: K = %d does not fit in shared memory", K);
  const CUtensorMapDataType dt = dtype == BLOBSPLAT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  p.n_levels = n_levels; p.nf = pl.nf; p.ns = pl.ns; p.nbuf = pl.nbuf; p.tma_store = store_mode; p.is_bf16 = dtype == BLOBSPLAT_BF16;
  p.f_bytes = p.Kp * kStM * 2;
  int items = 0, max_npx = 16, overhead = BS_ST_OVERHEAD;
  if (const char* e = std::getenv("BLOBSPLAT_ST_OVERHEAD")) overhead = std::max(0, std::atoi(e));
  long long cost = 0;
  for (int i = 0; i < n_levels; ++i) {
    StLevel& L = p.lv[i];
    const int P = H[i] * W[i];
    L.P = P;
    p.out_ptr[i] = outs[i]; p.out_C[i] = C[i];
    L.npx = std::min(pl.npx_max, round_up(P, 16));
    L.tiles = (P + L.npx - 1) / L.npx;
    L.groups = (C[i] + kStM - 1) / kStM;
    L.item_start = items;
    L.cost_start = cost;
    L.cost = std::min(L.npx, P) + overhead;
    items += N * L.tiles * L.groups;
    cost += (long long)N * L.tiles * L.groups * L.cost;
    max_npx = std::max(max_npx, L.npx);
    {
      const cuuint64_t dims[3] = {(cuuint64_t)C[i], (cuuint64_t)K, (cuuint64_t)N};
      const cuuint64_t strides[2] = {(cuuint64_t)C[i] * 2, (cuuint64_t)K * C[i] * 2};
      const cuuint32_t box[3] = {kStBlk, (cuuint32_t)p.Kp, 1};
      if (int rc = encode(&L.feats, dt, 3, feats[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {
      const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)K, (cuuint64_t)N};
      const cuuint64_t strides[2] = {(cuuint64_t)sk[i] * 2, (cuuint64_t)(N > 1 ? sn[i] : sk[i] * K) * 2};
      const cuuint32_t box[3] = {kStBlk, (cuuint32_t)p.Kp, 1};
      if (int rc = encode(&L.scores, dt, 3, scores[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
    {
      const cuuint64_t dims[3] = {(cuuint64_t)P, (cuuint64_t)C[i], (cuuint64_t)N};
      const cuuint64_t strides[2] = {(cuuint64_t)P * 2, (cuuint64_t)C[i] * P * 2};
      const cuuint32_t box[3] = {kStBoxPx, kStBoxCh, 1};
      if (int rc = encode(&L.out, dt, 3, outs[i], dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
    }
  }
  p.n_items = items; p.total_cost = cost; p.N = N;
  if (const char* e = std::getenv("BLOBSPLAT_ST_ABL")) p.abl = std::atoi(e);
  if (const char* e = std::getenv("BLOBSPLAT_ST_DBG_CTA")) p.dbg_cta = std::atoi(e);
  if (const char* e = std::getenv("BLOBSPLAT_ST_DBG_PTR")) p.dbg = reinterpret_cast<unsigned long long*>(std::strtoull(e, nullptr, 0));
  p.s_bytes = p.Kp * round_up(max_npx, kStBlk) * 2;
  int slot = 64;                     // a drain box reads 64 columns: slots are whole boxes
  while (slot < max_npx) slot <<= 1;
  p.slot_cols = slot;
  if (items == 0) return 0;
  const size_t smem = (size_t)kStDrainWarps * p.nbuf * kStBoxBytes + (size_t)p.ns * p.s_bytes + (size_t)p.nf * p.f_bytes +
                      sizeof(StBarriers) + 64;
  const int grid = std::min(sm_count, items);
  const bool dbg = p.tma_store != 0 || p.abl != 0 || p.dbg != nullptr;            // measurement knobs: the instrumented instantiation
  if (dbg) {
    if (dtype == BLOBSPLAT_BF16) BS_CUDA(launch_pdl(splat_tma_kernel<true, true>, dim3(grid), dim3(kStThreads), smem, st, p));
    else BS_CUDA(launch_pdl(splat_tma_kernel<false, true>, dim3(grid), dim3(kStThreads), smem, st, p));
  } else {
    if (dtype == BLOBSPLAT_BF16) BS_CUDA(launch_pdl(splat_tma_kernel<true, false>, dim3(grid), dim3(kStThreads), smem, st, p));
    else BS_CUDA(launch_pdl(splat_tma_kernel<false, false>, dim3(grid), dim3(kStThreads), smem, st, p));
  }
  return 0;
}

}  // namespace blobsplat
