// Fused render for 16-bit maps, two pixels per lane — same pipeline as render_tc.cuh (read that first), different
// tile geometry.  Why: with one pixel per TMEM lane a warp owns 32 pixels = 64 bytes of a bf16/f16 plane, so every
// store of the drain and of the composed maps is a half line and each value costs a convert + an extract + a store.
// Here a tile is 256 consecutive pixels and lane r holds the ADJACENT pixels 2r and 2r+1:
//
//   * two A operands in tensor memory (even pixels / odd pixels), two MMAs per k-step that share B:
//       D_even[lane, c] = sum_k A_even[lane, k] B[k, c]      D_odd likewise, in the next cw columns
//   * the drain packs (D_even, D_odd) of a channel into one 32-bit word: one convert + one store per TWO values, and a
//     warp writes 64 pixels = one whole 128-byte line;
//   * the composed maps leave stages 1+2 the same way (one packed 32-bit store per plane and lane);
//   * stages 1+2 share the row terms between the two pixels: per blob 2 FADD + 3 FFMA for the first pixel and
//     2 FADD for the second (u + p, v + r) instead of 2 x (2 FADD + 3 FFMA);
//   * the channel tile is drained in nsub sub-steps of cw channels (2*cw accumulator columns, two slots) so that the MMAs
//     of sub-step j+1 overlap the drain of sub-step j:  TMEM = 4*cw (D) + Kp (A even + odd) <= 512 columns.
//
// Preconditions (checked on the host, render_tc2_usable): 16-bit maps, W even (the two pixels of a lane share a row),
// output bases 4-byte aligned, score maps pixel-contiguous with even strides.  Everything else stays on render_tc.cuh.
#pragma once
#include "render_tc.cuh"

#ifndef BS_CW_MAX
#define BS_CW_MAX 128          // widest drain sub-step (channels)
#endif
#ifndef BS_HEAD_STAGE
#define BS_HEAD_STAGE 1        // two-pixel kernel without the B ring: the drain warps stage the CTA's first unit
#endif
#ifndef BS_PX2
#define BS_PX2 1               // 16-bit maps: use the two-pixels-per-lane kernel where its preconditions hold
#endif

namespace blobsplat {

constexpr int kTc2TilePx = 2 * kTcTileM;   // 256 pixels per tile

// both pixels of a lane (x, x+1 on one row) against one positive-definite blob
__device__ __forceinline__ void blob_opacity_pd2(const BlobCoef& c, float xf, float yf, float& sa, float& sb) {
  const float dyh = yf - c.cy_hi, dxh = xf - c.cx_hi;
  const float tv = fmaf(c.t, dyh, c.v0);
  const float ua = fmaf(c.p, dxh, c.u0), va = fmaf(c.r, dxh, tv);
  const float ub = ua + c.p, vb = va + c.r;
  sa = opacity_from_q2m1(fmaf(ua, ua, fmaf(va, va, c.c0)));
  sb = opacity_from_q2m1(fmaf(ub, ub, fmaf(vb, vb, c.c0)));
}

template <typename OT>
__device__ __forceinline__ uint32_t pack2(float a, float b) {   // low half = a (even pixel)
  if constexpr (std::is_same<OT, __nv_bfloat16>::value) {
    const __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
  } else {
    const __half2 t = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&t);
  }
}
template <typename OT>
__device__ __forceinline__ void unpack2(uint32_t w, float& a, float& b) {
  if constexpr (std::is_same<OT, __nv_bfloat16>::value) {
    const __nv_bfloat162 t = *reinterpret_cast<const __nv_bfloat162*>(&w);
    a = __low2float(t); b = __high2float(t);
  } else {
    const __half2 t = *reinterpret_cast<const __half2*>(&w);
    a = __low2float(t); b = __high2float(t);
  }
}
// two adjacent pixels of one plane
template <typename OT>
__device__ __forceinline__ void store_px2(OT* p, float a, float b, bool pred) {
  const uint32_t w = pack2<OT>(a, b);
  if (pred) __stcs(reinterpret_cast<unsigned int*>(p), w);
}

// two adjacent pixels of one plane, and (fused pyramid) their horizontal half-sum as the level-1 resize sees them: rounded to
// the map dtype first, 0.5 * a + 0.5 * b like pyramid_kernel (resize.cu)
template <typename OT, bool kPyr>
__device__ __forceinline__ void store_px2_h(OT* p, float a, float b, bool pred, float* hs, bool hs_pred) {
  const uint32_t w = pack2<OT>(a, b);
  if (pred) __stcs(reinterpret_cast<unsigned int*>(p), w);
  if constexpr (kPyr) {
    float ra, rb;
    unpack2<OT>(w, ra, rb);
    if (hs_pred) *hs = 0.5f * ra + 0.5f * rb;
  }
}
template <typename OT>
__device__ __forceinline__ float round_to(float v) { return Cvt<OT>::to(Cvt<OT>::from(v)); }

// 16 channel planes of this lane's two pixels: e[i] / o[i] = fp32 accumulators of the even / odd pixel
template <typename OT>
__device__ __forceinline__ void drain16(OT* oc, size_t P, const uint32_t* e, const uint32_t* o, bool live, int left) {
  if (left >= 16) {
#pragma unroll
    for (int i = 0; i < 16; ++i) store_px2<OT>(oc + (size_t)i * P, __uint_as_float(e[i]), __uint_as_float(o[i]), live);
  } else {
#pragma unroll
    for (int i = 0; i < 16; ++i) store_px2<OT>(oc + (size_t)i * P, __uint_as_float(e[i]), __uint_as_float(o[i]), live && i < left);
  }
}

// kP = -1: several stage-3 problems of one pyramid (RenderTcLevels) in ONE launch over the concatenated tile sequence —
// every work unit reads its own level's shape, pointers and strides; Kp, c_tile, cw and the B ring are common.
// kPyr (fused render of 64 x 64 maps only): the halving pyramid of the composed maps (pyramid_resize, utils.py:280-294) leaves
// the same launch.  Every final pair of composed values also goes, as its rounded horizontal half-sum, into a staging
// array hs[slot][k][tile row] (hand-over by mbarriers hs_full / hs_free); three extra warps (the places of the B ring's
// staging warps; one warp alone needs longer than a tile: shuffle / convert chains at one warp's issue rate) turn the tile's four image
// rows into two level-1 rows and one level-2 row — exact 2x2 means, rounded between levels like the reference's repeated
// interpolate.  Level 3 needs two tiles (eight image rows), which equal tile ranges split between CTAs.  A cross-CTA
// hand-over (arrival counter + __threadfence) was measured at +18 us .. +60 us on cfg3's level 64 — MEMBAR.GPU in an SM
// whose store queue is saturated waits for the whole queue.  Instead the CTA whose range starts on the odd tile of a pair
// composites the even tile once more as a ghost (stages 1+2 only, ~1 us of the compute warps' slack), so every pair's
// level-3 row is made inside one CTA.
template <typename OT, int kP, bool kFromScores, bool kRing, bool kPyr = false>
__global__ void __launch_bounds__((13 + (kRing ? kTcStageWarps : 0) + (kPyr ? 3 : 0)) * 32, 1)
render_tc2_kernel(const __grid_constant__ RenderTcLevels L) {
  static_assert(!kPyr || (kP == 4096 && !kFromScores && !kRing), "the fused pyramid is specialised for 64 x 64 renders without the B ring");
  constexpr int kPyrWarp = 13, kPyrWarps = 3;               // kPyr: warps 13-15 turn staged half-sums into pyramid rows (planes k = w, w + 3, ..)
  const RenderTcParams& p0 = L.lv[0];
  constexpr bool kLevels = kP == -1;
  static_assert(!kLevels || kFromScores, "a multi-level launch is stage 3 from score maps");
  constexpr int kComputeWarps = 8, kComputeThreads = 256, kMmaWarp = 12;
  extern __shared__ __align__(1024) unsigned char smem[];

  const int cw = p0.cw, nsub = p0.c_tile / cw;                            // drain sub-steps of cw channels
  const size_t b_bytes = (size_t)(p0.Kp / 8) * p0.c_tile * 16;           // one B buffer (N-major, see tc_stage_b)
  const int nb = kRing ? p0.nb : 1;
  unsigned char* b_smem = smem;
  const int srow = p0.Kp + 4;
  float* stash = reinterpret_cast<float*>(smem + (size_t)nb * b_bytes);  // [2 parities][128 lanes][Kp + 4]
  float* carry = stash + (size_t)2 * kTcTileM * srow;                    // [2][128] front-range transmittances
  BlobCoef* coef = reinterpret_cast<BlobCoef*>(carry + 2 * kTcTileM);
  TcBarriers* bars = reinterpret_cast<TcBarriers*>(coef + kTcMaxBlobs + 1);
  uint64_t* const hs_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<unsigned char*>(bars) + ((sizeof(TcBarriers) + 15) & ~(size_t)15));   // kPyr: full[2], free[2]
  float* const hs = reinterpret_cast<float*>(hs_bar + 4);                                                                // kPyr: [2][K][128]
  float* const l3h = hs + (size_t)2 * p0.K * kTcTileM;                                                                   // kPyr: [K][8] level-3 half-sums of a pair's even tile

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == kMmaWarp) {
    if (lane == 0) {
      mbar_init(&bars->a_full[0], kComputeThreads); mbar_init(&bars->a_free[0], 1);
      for (int i = 0; i < kTcMaxB; ++i) {
        mbar_init(&bars->b_full[i], kRing ? kTcStageWarps * 32 : kComputeThreads); mbar_init(&bars->b_free[i], 1);
      }
      mbar_init(&bars->d_full[0], 1); mbar_init(&bars->d_full[1], 1);
      mbar_init(&bars->d_empty[0], 128); mbar_init(&bars->d_empty[1], 128);
      if constexpr (kPyr) {
        mbar_init(&hs_bar[0], kComputeWarps); mbar_init(&hs_bar[1], kComputeWarps); mbar_init(&hs_bar[2], kPyrWarps); mbar_init(&hs_bar[3], kPyrWarps);
      }
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  pdl_launch_dependents();
  pdl_wait();                                            // set-up above overlapped the previous kernel's tail
  const uint32_t tmem_a = tmem + (uint32_t)(4 * cw);     // A even; A odd follows at + a_cols
  const int a_cols = p0.Kp / 2;                          // two k per 32-bit column

  int unit_it = 0, tile_it = 0;
  int pyr_it = 0;                                        // kPyr: tiles that produced pyramid rows so far (staging slot = pyr_it & 1)
  int sub_it = 0;                                        // drain sub-steps so far (slot = sub_it & 1)

  const int seq_tiles = kLevels ? L.tile_start[L.n_levels] : p0.total_tiles;
  const bool whole_runs = !kLevels && p0.whole_runs;
  const int g_end = whole_runs ? seq_tiles : tc_range_begin(seq_tiles, (int)blockIdx.x + 1, (int)gridDim.x);
  int level = 0;
  for (int gs = whole_runs ? (int)blockIdx.x * p0.tiles_per_image : tc_range_begin(seq_tiles, (int)blockIdx.x, (int)gridDim.x);
       gs < g_end; ++unit_it) {
    if constexpr (kLevels) {
      while (gs >= L.tile_start[level + 1]) ++level;
    }
    const RenderTcParams& p = L.lv[kLevels ? level : 0];
    const int P = kP > 0 ? kP : p.H * p.W;
    const int g = gs - (kLevels ? L.tile_start[level] : 0);         // tile index within the level
    const int img_chunk = g / p.tiles_per_image;
    const int n = img_chunk / p.c_chunks;
    const int chunk = img_chunk - n * p.c_chunks;
    const int c0 = chunk * p.c_tile;
    const int t_lo = g - img_chunk * p.tiles_per_image;
    const int ntiles = min(p.tiles_per_image - t_lo, g_end - gs);
    gs += ntiles;
    if (whole_runs) gs += ((int)gridDim.x - 1) * p.tiles_per_image;

    if (warp < kComputeWarps) {
      // =============================== stages 1+2 (two pixels per lane) ===============================
      const int ctid = warp * 32 + lane;
      const int half = warp >> 2, q = warp & 3;
      const int row = q * 32 + lane;                       // TMEM lane / stash row; pixels 2*row, 2*row + 1 of the tile
      asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory");
      constexpr int kLd = BS_SCORE_LOADS;
      const int k_split = p.K >> 1;
      const int k_lo = half ? 0 : k_split, k_hi = half ? k_split : p.K;
      const long long sk2 = p.sk >> 1;                     // plane stride in 32-bit words (pixel pairs)
      uint32_t pre[kFromScores ? kLd : 1];
      if constexpr (kFromScores) {
        const int pix0 = t_lo * kTc2TilePx + 2 * row;
        const uint32_t* sc0 = reinterpret_cast<const uint32_t*>(reinterpret_cast<const OT*>(p.scores) + (size_t)n * p.sn +
                                                                (pix0 < P ? pix0 : 0));
#pragma unroll
        for (int j = 0; j < kLd; ++j) pre[j] = (pix0 < P && k_lo + j < k_hi) ? __ldg(sc0 + (size_t)(k_lo + j) * sk2) : 0u;
      }
      uint32_t my_general = 0;
      if constexpr (!kFromScores)
        for (int i = ctid; i < p.M; i += kComputeThreads) {
          const size_t b = (size_t)n * p.M + i;
          const float* c = p.covs + 4 * b;
          const BlobCoef bc = make_blob_coef((double)p.xs[b], (double)p.ys[b], (double)c[0], (double)c[1], (double)c[2],
                                             (double)c[3], p.sizes[b], p.H, p.W);
          my_general |= coef_general(bc) ? 1u : 0u;
          coef[i] = bc;
        }
      if (unit_it == 0)      // stash columns that are never written stay zero for the whole kernel
        for (int i = ctid; i < 2 * kTcTileM * srow; i += kComputeThreads) stash[i] = 0.0f;
      if constexpr (!kRing) {
        // the CTA's first unit is staged by the drain warps, which have nothing to drain yet: the features' load round trip
        // leaves the head of the launch (BS_HEAD_STAGE); later units are staged here, behind the previous unit's drain
        if (!(BS_HEAD_STAGE && unit_it == 0)) {
          if (unit_it > 0) mbar_wait(&bars->b_free[0], (unit_it - 1) & 1);
          tc_stage_b<OT, OT, false>(p, n, c0, b_smem, b_bytes, ctid, kComputeThreads);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          mbar_arrive(&bars->b_full[0]);
        }
      }
      uint32_t any_general;
      asm volatile(
          "{\n\t.reg .pred q, r;\n\tsetp.ne.u32 q, %1, 0;\n\t"
          "barrier.cta.red.or.pred r, 1, %2, q;\n\tselp.u32 %0, 1, 0, r;\n\t}"
          : "=r"(any_general) : "r"(my_general), "n"(kComputeThreads) : "memory");

      const int m_split = ((p.M * BS_SPLIT_NUM) >> 4) & ~7;       // back range (it also rescales): the smaller share
      const int m_lo = half ? 0 : m_split, m_hi = half ? m_split : p.M;
      const int pair_bar = 2 + q;
      OT* comp = (p.composed && chunk == 0) ? reinterpret_cast<OT*>(p.composed) + (size_t)n * p.K * P : nullptr;
      float* const my_e = stash + (size_t)row * srow + kTcKOff;   // my_e[k] = plane k of the even pixel
      float* const my_o = my_e + (size_t)kTcTileM * srow;
      // kPyr with level 3: a level-3 row needs a PAIR of tiles; when this CTA's range starts on the odd tile of a pair, it
      // first composites the even tile as a ghost (t = -1: stages 1+2 into the pyramid staging only — no composed-map
      // stores, no A operand), so the pair's level-3 row is made here and no cross-CTA hand-over exists
      const int t_first = (kPyr && !kFromScores && p.pyr_levels >= 3 && p.composed && chunk == 0 && (t_lo & 1)) ? -1 : 0;
      for (int t = t_first; t < ntiles; ++t) {
        const bool ghost = t < 0;
        const int pix0 = (t_lo + t) * kTc2TilePx + 2 * row;
        const bool live = pix0 < P;                           // P even: both pixels or none
        const int y = live ? pix0 / p.W : 0;
        const float xf = (float)(live ? pix0 - y * p.W : 0), yf = (float)y;
        if constexpr (kFromScores) {
          const uint32_t* sc = reinterpret_cast<const uint32_t*>(reinterpret_cast<const OT*>(p.scores) + (size_t)n * p.sn +
                                                                 (live ? pix0 : 0));
          if (t == 0) {
#pragma unroll
            for (int j = 0; j < kLd; ++j)
              if (k_lo + j < k_hi) unpack2<OT>(pre[j], my_e[k_lo + j], my_o[k_lo + j]);
          }
          for (int k = k_lo + (t == 0 ? kLd : 0); k < k_hi; k += kLd) {
            uint32_t v[kLd];
#pragma unroll
            for (int j = 0; j < kLd; ++j) v[j] = (live && k + j < k_hi) ? __ldg(sc + (size_t)(k + j) * sk2) : 0u;
#pragma unroll
            for (int j = 0; j < kLd; ++j)
              if (k + j < k_hi) unpack2<OT>(v[j], my_e[k + j], my_o[k + j]);
          }
        } else {
          float Ta = 1.0f, Tb = 1.0f;
          const bool wr = comp != nullptr && live && !BS_ABL_NO_COMP_STORE && !ghost;
          const bool wr_now = wr && half == 0;
          OT* const comp_px = comp + pix0;
          const bool do_pyr = kPyr && comp != nullptr && p.pyr_levels > 0;          // CTA-uniform
          float* const hs_row = hs + (size_t)(pyr_it & 1) * p.K * kTcTileM + row;    // + k * 128: plane k of this tile row
          const bool hs_wr = do_pyr && !(p.pyr_dbg & 4);
          const bool hs_now = hs_wr && half == 0;
          if constexpr (kPyr) {
            if (do_pyr && pyr_it >= 2) mbar_wait(&hs_bar[2 + (pyr_it & 1)], (uint32_t)(((pyr_it >> 1) - 1) & 1));   // the pyramid warp has read this slot
          }
          int m = m_hi;
          // serial head until the remaining blobs of the range are whole, float4-aligned groups of 4
          for (; m >= m_lo + 1 && (any_general || (m & 3) != 0 || m < m_lo + 4); --m) {
            float sa, sb;
            if (any_general) { sa = blob_opacity(coef[m - 1], xf, yf); sb = blob_opacity(coef[m - 1], xf + 1.0f, yf); }
            else blob_opacity_pd2(coef[m - 1], xf, yf, sa, sb);
            const float da = sa * Ta, db = sb * Tb;
            Ta = fmaf(-sa, Ta, Ta); Tb = fmaf(-sb, Tb, Tb);
            my_e[m] = da; my_o[m] = db;
            store_px2_h<OT, kPyr>(comp_px + (size_t)m * P, da, db, wr_now, hs_row + m * kTcTileM, hs_now);
          }
          // branch-free: 4 blobs x 2 pixels in flight
          for (; m >= m_lo + 4; m -= 4) {
            float sa[4], sb[4], da[4], db[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) blob_opacity_pd2(coef[m - 1 - j], xf, yf, sa[j], sb[j]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              da[j] = sa[j] * Ta; Ta = fmaf(-sa[j], Ta, Ta);
              db[j] = sb[j] * Tb; Tb = fmaf(-sb[j], Tb, Tb);
            }
            OT* const cp = comp_px + (size_t)m * P;
#pragma unroll
            for (int j = 0; j < 4; ++j) store_px2_h<OT, kPyr>(cp - (ptrdiff_t)j * P, da[j], db[j], wr_now, hs_row + (m - j) * kTcTileM, hs_now);
            *reinterpret_cast<float4*>(my_e + m - 3) = make_float4(da[3], da[2], da[1], da[0]);
            *reinterpret_cast<float4*>(my_o + m - 3) = make_float4(db[3], db[2], db[1], db[0]);
          }
          for (; m >= m_lo + 1; --m) {
            float sa, sb;
            if (any_general) { sa = blob_opacity(coef[m - 1], xf, yf); sb = blob_opacity(coef[m - 1], xf + 1.0f, yf); }
            else blob_opacity_pd2(coef[m - 1], xf, yf, sa, sb);
            const float da = sa * Ta, db = sb * Tb;
            Ta = fmaf(-sa, Ta, Ta); Tb = fmaf(-sb, Tb, Tb);
            my_e[m] = da; my_o[m] = db;
            store_px2_h<OT, kPyr>(comp_px + (size_t)m * P, da, db, wr_now, hs_row + m * kTcTileM, hs_now);
          }
          if (half == 0) { carry[row] = Ta; carry[kTcTileM + row] = Tb; }
          asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");
          if (half == 1) {
            const float ca = carry[row], cb = carry[kTcTileM + row];
            int k = m_hi;
            for (; k >= 1 && ((k & 3) != 0 || k < 4); --k) {
              const float va = my_e[k] * ca, vb = my_o[k] * cb;
              my_e[k] = va; my_o[k] = vb;
              store_px2_h<OT, kPyr>(comp_px + (size_t)k * P, va, vb, wr, hs_row + k * kTcTileM, hs_wr);
            }
            for (; k >= 4; k -= 4) {
              float4 a4 = *reinterpret_cast<const float4*>(my_e + k - 3), b4 = *reinterpret_cast<const float4*>(my_o + k - 3);
              a4.x *= ca; a4.y *= ca; a4.z *= ca; a4.w *= ca;
              b4.x *= cb; b4.y *= cb; b4.z *= cb; b4.w *= cb;
              *reinterpret_cast<float4*>(my_e + k - 3) = a4;
              *reinterpret_cast<float4*>(my_o + k - 3) = b4;
              OT* const cp = comp_px + (size_t)k * P;
              store_px2_h<OT, kPyr>(cp, a4.w, b4.w, wr, hs_row + k * kTcTileM, hs_wr);
              store_px2_h<OT, kPyr>(cp - (ptrdiff_t)P, a4.z, b4.z, wr, hs_row + (k - 1) * kTcTileM, hs_wr);
              store_px2_h<OT, kPyr>(cp - (ptrdiff_t)2 * P, a4.y, b4.y, wr, hs_row + (k - 2) * kTcTileM, hs_wr);
              store_px2_h<OT, kPyr>(cp - (ptrdiff_t)3 * P, a4.x, b4.x, wr, hs_row + (k - 3) * kTcTileM, hs_wr);
            }
            for (; k >= 1; --k) {
              const float va = my_e[k] * ca, vb = my_o[k] * cb;
              my_e[k] = va; my_o[k] = vb;
              store_px2_h<OT, kPyr>(comp_px + (size_t)k * P, va, vb, wr, hs_row + k * kTcTileM, hs_wr);
            }
            const float bga = Ta * ca, bgb = Tb * cb;          // background: alpha 1 * total transmittance
            my_e[0] = bga; my_o[0] = bgb;
            store_px2_h<OT, kPyr>(comp_px, bga, bgb, wr, hs_row, hs_wr);
          }
        }
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // quarter's stash rows complete
        if (!ghost) {
        if (tile_it > 0) mbar_wait(&bars->a_free[0], (tile_it - 1) & 1);   // previous tile's MMAs have read A
        tc_fence_after();
        const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
        for (int g = half; g < p.Kp / 16; g += 2) {          // the two warps of a quarter interleave the k-groups
#pragma unroll
          for (int par = 0; par < 2; ++par) {
            const float* src = (par ? my_o : my_e) - kTcKOff + g * 16;     // operand rows 16g .. 16g+15
            const float4 q0 = *reinterpret_cast<const float4*>(src), q1 = *reinterpret_cast<const float4*>(src + 4);
            const float4 q2 = *reinterpret_cast<const float4*>(src + 8), q3 = *reinterpret_cast<const float4*>(src + 12);
            const float wv[16] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, q3.x, q3.y, q3.z, q3.w};
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) pk[j] = pack2<OT>(wv[2 * j], wv[2 * j + 1]);   // low half = even k
            tmem_st8(tmem_a + lane_addr + (uint32_t)(par * a_cols + g * 8), pk);
          }
        }
        tmem_wait_st();
        tc_fence_before();
        mbar_arrive(&bars->a_full[0]);
        asm volatile("bar.sync %0, 64;" ::"r"(pair_bar) : "memory");   // partner has read the stash
        ++tile_it;
        }
        if constexpr (kPyr) {
          if (comp != nullptr && p.pyr_levels > 0) {           // this warp's half-sums of the tile are staged
            __syncwarp();
            if (lane == 0) mbar_arrive(&hs_bar[pyr_it & 1]);
            ++pyr_it;
          }
        }
      }
    } else if (warp < kComputeWarps + 4) {
      // ============================ drain: (even, odd) pixel pairs, 32-bit stores ============================
      const int q = warp - kComputeWarps;
      if constexpr (!kRing) {
        if (BS_HEAD_STAGE && unit_it == 0) {               // 128 threads for 256 arrivals: two each
          tc_stage_b<OT, OT, false>(p, n, c0, b_smem, b_bytes, (int)threadIdx.x - kComputeThreads, 128);
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0], 2;" ::"r"(smem_u32(&bars->b_full[0])) : "memory");
        }
      }
      OT* const out = reinterpret_cast<OT*>(p.grid) + ((size_t)n * p.C + c0) * P;
      const uint32_t lane_addr = ((uint32_t)(q * 32) << 16);
      for (int t = 0; t < ntiles; ++t, ++tile_it) {
        const int pix0 = (t_lo + t) * kTc2TilePx + 2 * (q * 32 + lane);
        const bool live = pix0 < P && !BS_ABL_NO_EPI_STORE;
#pragma unroll 1
        for (int j = 0; j < nsub; ++j, ++sub_it) {
          const int slot = sub_it & 1;
          mbar_wait_drain(&bars->d_full[slot], (sub_it >> 1) & 1);
          tc_fence_after();
          const uint32_t te = tmem + lane_addr + (uint32_t)(slot * 2 * cw), to = te + (uint32_t)cw;
          OT* const o = out + (size_t)(j * cw) * P + pix0;
          const int chs = p.C - (c0 + j * cw);                 // valid channels from this sub-step on
          uint32_t ea[16], oa[16], eb[16], ob[16];
          tmem_ld16(te, ea); tmem_ld16(to, oa);
          tmem_wait_ld();
          for (int cc = 0; cc < cw; cc += 32) {
            const bool more = cc + 16 < cw;
            if (more) { tmem_ld16(te + cc + 16, eb); tmem_ld16(to + cc + 16, ob); }
            drain16<OT>(o + (size_t)cc * P, (size_t)P, ea, oa, live, chs - cc);
            tmem_wait_ld();
            if (more) {
              if (cc + 32 < cw) { tmem_ld16(te + cc + 32, ea); tmem_ld16(to + cc + 32, oa); }
              drain16<OT>(o + (size_t)(cc + 16) * P, (size_t)P, eb, ob, live, chs - cc - 16);
              tmem_wait_ld();
            }
          }
          tc_fence_before();
          mbar_arrive(&bars->d_empty[slot]);
        }
      }
    } else if (warp == kMmaWarp) {
      // ========================================= MMA issue =========================================
      if (lane == 0) {
        const int buf = unit_it % nb, rnd = unit_it / nb;
        const uint32_t idesc = make_idesc(std::is_same<OT, __half>::value ? 0u : 1u, (uint32_t)cw, 1u);
        const uint32_t b_base = smem_u32(b_smem + (size_t)buf * b_bytes);
        const uint32_t lbo = (uint32_t)p.c_tile * 16u, sbo = 128u;
        mbar_wait_mma(&bars->b_full[buf], rnd & 1);
        for (int t = 0; t < ntiles; ++t, ++tile_it) {
          mbar_wait_mma(&bars->a_full[0], tile_it & 1);
          tc_fence_after();
          for (int j = 0; j < nsub; ++j, ++sub_it) {
            const int slot = sub_it & 1;
            if (sub_it >= 2) mbar_wait_mma(&bars->d_empty[slot], ((sub_it >> 1) - 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int par = 0; par < 2; ++par) {
              const uint32_t d_addr = tmem + (uint32_t)(slot * 2 * cw + par * cw);
              uint32_t acc = 0;
              for (int ks = 0; ks < p.Kp / 16; ++ks) {
                const uint32_t b_addr = b_base + (uint32_t)(2 * ks) * lbo + (uint32_t)(j * cw) * 16u;
                if (!BS_ABL_NO_MMA)
                  umma_ts<false>(d_addr, tmem_a + (uint32_t)(par * a_cols + ks * 8), make_b_desc(b_addr, lbo, sbo), idesc, acc);
                acc = 1;
              }
            }
            tc_commit(&bars->d_full[slot]);
          }
          tc_commit(&bars->a_free[0]);
        }
        tc_commit(&bars->b_free[buf]);
      }
      __syncwarp();
    } else if (kPyr && warp >= kPyrWarp) {
      // ============================ fused pyramid: staged half-sums -> level 1 / 2 / 3 rows ============================
      if constexpr (kPyr) {
        if (p.composed != nullptr && chunk == 0 && p.pyr_levels > 0) {
          const int t_first = (p.pyr_levels >= 3 && (t_lo & 1)) ? -1 : 0;      // the ghost tile (see the compute warps)
          for (int t = t_first; t < ntiles; ++t, ++pyr_it) {
            const bool ghost = t < 0;
            mbar_wait(&hs_bar[pyr_it & 1], (uint32_t)((pyr_it >> 1) & 1));
            const float* const H = hs + (size_t)(pyr_it & 1) * p.K * kTcTileM;
            const int tt = t_lo + t;                           // tile of the image: rows 4 tt .. 4 tt + 3
            OT* const l3 = p.pyr_levels >= 3 ? reinterpret_cast<OT*>(p.pyr[2]) + (size_t)n * p.K * 64 + (tt >> 1) * 8 + (lane >> 2) : nullptr;
            // four planes per step: one warp carries the whole tile, so the load -> mean -> round -> shuffle chains of
            // independent planes have to overlap
            OT* const l1 = reinterpret_cast<OT*>(p.pyr[0]) + (size_t)n * p.K * 1024 + (2 * tt) * 32 + lane;
            OT* const l2 = p.pyr_levels >= 2 ? reinterpret_cast<OT*>(p.pyr[1]) + (size_t)n * p.K * 256 + tt * 16 + (lane >> 1) : nullptr;
            constexpr int kU = 4;
            for (int k0 = (warp - kPyrWarp) * kU; k0 < ((p.pyr_dbg & 1) ? 0 : p.K); k0 += kPyrWarps * kU) {
              float v0[kU], v1[kU];
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const float* hk = H + min(k0 + u, p.K - 1) * kTcTileM;
                v0[u] = round_to<OT>(0.5f * hk[lane] + 0.5f * hk[32 + lane]);        // level 1, row 2 tt
                v1[u] = round_to<OT>(0.5f * hk[64 + lane] + 0.5f * hk[96 + lane]);   // level 1, row 2 tt + 1
              }
#pragma unroll
              for (int u = 0; u < kU; ++u) {
                const int k = k0 + u;
                const bool ok = k < p.K && !(p.pyr_dbg & 2) && !ghost;   // a ghost tile's level-1/2 rows belong to the CTA that owns it
                const float n0 = __shfl_down_sync(0xffffffffu, v0[u], 1), n1 = __shfl_down_sync(0xffffffffu, v1[u], 1);
                store_px2<OT>(l1 + (size_t)k * 1024, v0[u], n0, ok && (lane & 1) == 0);
                store_px2<OT>(l1 + (size_t)k * 1024 + 32, v1[u], n1, ok && (lane & 1) == 0);
                if (l2 != nullptr) {
                  const float g0 = 0.5f * v0[u] + 0.5f * __shfl_xor_sync(0xffffffffu, v0[u], 1);
                  const float g1 = 0.5f * v1[u] + 0.5f * __shfl_xor_sync(0xffffffffu, v1[u], 1);
                  const float w2 = round_to<OT>(0.5f * g0 + 0.5f * g1);                       // level 2, row tt, x = lane >> 1
                  const float w2n = __shfl_down_sync(0xffffffffu, w2, 2);
                  store_px2<OT>(l2 + (size_t)k * 256, w2, w2n, ok && (lane & 3) == 0);
                  if (l3 != nullptr) {
                    // level 3 (x = lane >> 2): the even tile of a pair leaves its horizontal half-sums in l3h (this warp's own
                    // planes: no hand-over); the odd tile adds its own and writes the row
                    const float h = 0.5f * w2 + 0.5f * __shfl_xor_sync(0xffffffffu, w2, 2);
                    float* const slot = l3h + min(k, p.K - 1) * 8 + (lane >> 2);
                    if ((tt & 1) == 0) {
                      if ((lane & 3) == 0 && k < p.K) *slot = h;
                    } else {
                      const float w3 = round_to<OT>(0.5f * *slot + 0.5f * h);
                      const float w3n = __shfl_down_sync(0xffffffffu, w3, 4);
                      store_px2<OT>(l3 + (size_t)k * 64, w3, w3n, k < p.K && (lane & 7) == 0);
                    }
                  }
                }
              }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&hs_bar[2 + (pyr_it & 1)]);   // the staging slot may be rewritten
          }
        }
      }
    } else if constexpr (kRing) {
      const int buf = unit_it % nb, rnd = unit_it / nb;
      if (rnd > 0) mbar_wait(&bars->b_free[buf], (rnd - 1) & 1);
      tc_stage_b<OT, OT, false>(p, n, c0, b_smem + (size_t)buf * b_bytes, b_bytes, (int)threadIdx.x - (kMmaWarp + 1) * 32,
                                kTcStageWarps * 32);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(&bars->b_full[buf]);
    }
  }

  if (warp == kMmaWarp && lane == 0 && tile_it > 0) {
    mbar_wait(&bars->a_free[0], (tile_it - 1) & 1);
    mbar_wait(&bars->b_free[(unit_it - 1) % nb], ((unit_it - 1) / nb) & 1);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
struct Tc2Plan { int Kp, c_tile, cw, nb; size_t smem, b_slot; bool ok; };

static inline Tc2Plan plan_tc2(int K, int C, bool pyr = false, long long tiles = 0) {
  Tc2Plan pl{};
  pl.ok = false;
  if (K - 1 > kTcMaxBlobs || C < 1) return pl;
  pl.Kp = round_up(K + kTcKOff, 16);
  const size_t fixed = (size_t)2 * (pl.Kp + 4) * kTcTileM * 4 + 2 * kTcTileM * 4 + (kTcMaxBlobs + 1) * sizeof(BlobCoef) +
                       sizeof(TcBarriers) + 512 + (pyr ? (size_t)2 * K * kTcTileM * 4 + (size_t)K * 32 + 64 : 0);   // + fused pyramid: barriers, hs[2][K][128], l3h[K][8]
  const size_t per_c = (size_t)pl.Kp * 2;
  if (fixed + per_c * 32 > kTcSmemBudget) return pl;
  int c_tile = std::min(kTcMaxCTile, round_up(C, 32));
  c_tile = std::min<long long>(c_tile, (long long)((kTcSmemBudget - fixed) / per_c) / 32 * 32);
  for (int c = c_tile; c >= std::max(32, c_tile / 2); c -= 32)
    if (C % c == 0) { c_tile = c; break; }
  c_tile = spread_c_tile(c_tile, C, tiles, 64);          // small launches: more, narrower units (render_tc.cuh)
  // drain sub-step: the widest multiple of 16 channels that divides the tile and leaves room for two 2*cw-column
  // slots next to the two A operands
  int cw = 0;
  for (int w = BS_CW_MAX; w >= 16; w -= 16)
    if (c_tile % w == 0 && 4 * w + pl.Kp <= 512) { cw = w; break; }
  if (cw == 0) return pl;
  pl.c_tile = c_tile; pl.cw = cw;
  pl.b_slot = per_c * c_tile;
  pl.nb = (int)std::min<size_t>(BS_MAX_B, (kTcSmemBudget - fixed) / pl.b_slot);
  pl.smem = fixed + (size_t)pl.nb * pl.b_slot;
  pl.ok = pl.nb >= 1;
  return pl;
}

// Can this 16-bit problem run on the two-pixels-per-lane kernel?
static inline bool render_tc2_usable(int dtype, int H, int W, const void* composed, const void* grid, const void* scores,
                                     long long sn, long long sk, long long sp, bool any_size = false) {
  if (!BS_PX2 || !BS_B_NMAJOR || (dtype != BLOBSPLAT_BF16 && dtype != BLOBSPLAT_F16)) return false;
  if ((W & 1) != 0) return false;
  if (!any_size && (long long)H * W < kTc2TilePx) return false;        // small images: the 128-pixel tile wastes less (single launches)
  if (((reinterpret_cast<uintptr_t>(composed) | reinterpret_cast<uintptr_t>(grid)) & 3) != 0) return false;
  if (scores && (sp != 1 || (sk & 1) != 0 || (sn & 1) != 0 || (reinterpret_cast<uintptr_t>(scores) & 3) != 0)) return false;
  return true;
}

template <typename OT, int kP, bool kFromScores, bool kRing, bool kPyr = false>
static int launch_tc2_pr(const RenderTcParams& p, cudaStream_t st) {
  static thread_local int configured_dev = -1, sm_count = 0;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(render_tc2_kernel<OT, kP, kFromScores, kRing, kPyr>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    configured_dev = dev;
  }
  const int grid = std::min(sm_count, p.total_tiles);
  RenderTcLevels L{};
  L.lv[0] = p;
  L.n_levels = 1;
  L.tile_start[1] = p.total_tiles;
  BS_CUDA(launch_pdl(render_tc2_kernel<OT, kP, kFromScores, kRing, kPyr>, dim3(grid),
                     dim3((13 + (kRing ? kTcStageWarps : 0) + (kPyr ? 3 : 0)) * 32), (size_t)p.smem_bytes, st, L));
  return 0;
}

template <typename OT, bool kFromScores>
static int launch_tc2(const RenderTcParams& p, cudaStream_t st) {
  const bool ring = p.nb > 1;
  if constexpr (!kFromScores) {
    if (p.pyr_levels > 0) return launch_tc2_pr<OT, 4096, false, false, true>(p, st);   // 64 x 64 render + pyramid (caller checked nb == 1)
  }
  switch (p.H * p.W) {
    case 4096: return ring ? launch_tc2_pr<OT, 4096, kFromScores, true>(p, st) : launch_tc2_pr<OT, 4096, kFromScores, false>(p, st);
    case 1024: return ring ? launch_tc2_pr<OT, 1024, kFromScores, true>(p, st) : launch_tc2_pr<OT, 1024, kFromScores, false>(p, st);
    case 256: return ring ? launch_tc2_pr<OT, 256, kFromScores, true>(p, st) : launch_tc2_pr<OT, 256, kFromScores, false>(p, st);
  }
  return ring ? launch_tc2_pr<OT, 0, kFromScores, true>(p, st) : launch_tc2_pr<OT, 0, kFromScores, false>(p, st);
}

// Several stage-3 problems of one pyramid in one launch (kP = -1): L.lv[*] filled by fill_tc_units with the 256-pixel tile.
template <typename OT, bool kRing>
static int launch_tc2_levels(RenderTcLevels& L, cudaStream_t st) {
  static thread_local int configured_dev = -1, sm_count = 0;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(render_tc2_kernel<OT, -1, true, kRing>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    configured_dev = dev;
  }
  long long total = 0;
  for (int i = 0; i < L.n_levels; ++i) { L.tile_start[i] = (int)total; total += L.lv[i].total_tiles; }
  if (total > 0x7fffffffll) BS_UNSUPPORTED("too many tiles for one launch");
  L.tile_start[L.n_levels] = (int)total;
  if (total == 0) return 0;
  const int grid = (int)std::min<long long>(sm_count, total);
  BS_CUDA(launch_pdl(render_tc2_kernel<OT, -1, true, kRing>, dim3(grid), dim3((13 + (kRing ? kTcStageWarps : 0)) * 32),
                     (size_t)L.lv[0].smem_bytes, st, L));
  return 0;
}

// Fill the plan-dependent fields and launch.  p: pointers (and score strides) already set.
template <bool kFromScores>
static int run_tc2(RenderTcParams& p, const Tc2Plan& pl, int N, int K, int H, int W, int C, int dtype, cudaStream_t st) {
  TcPlan base{};
  base.Kp = pl.Kp; base.c_tile = pl.c_tile; base.nb = pl.nb; base.smem = pl.smem; base.b_slot = pl.b_slot; base.ok = true;
  if (int rc = fill_tc_units(p, base, N, K, H, W, C, kTc2TilePx)) return rc;
  p.cw = pl.cw;
  if (dtype == BLOBSPLAT_BF16) return launch_tc2<__nv_bfloat16, kFromScores>(p, st);
  return launch_tc2<__half, kFromScores>(p, st);
}

}  // namespace blobsplat
