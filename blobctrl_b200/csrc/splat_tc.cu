// Stand-alone stage 3 on tensor cores — instantiations of render_tc.cuh with the A operand loaded from
// precomputed score maps.  Replaces splat_features_from_scores (blobctrl/utils/utils.py:57-77; duplicate at
// blobctrl/pipelines/pipeline_blobnet.py:706-721) when K and C make it a real dense contraction
// (K >= 12, C >= 64); smaller shapes stay on the FMA engine (feature_splat.cu).
#include "render_tc2.cuh"

namespace blobsplat {

int feature_splat_tc_dispatch(const void* scores, int64_t sn, int64_t sk, int64_t sp, const void* feats, void* out,
                              int N, int K, int C, int H, int W, int dtype, cudaStream_t st) {
  const long long px = (long long)H * W, tiles = N * ((px + kTcTileM - 1) / kTcTileM);
  int split = split_for_launch(dtype, tiles);
  TcPlan pl = plan_tc(K, C, split, tiles);
  if (!pl.ok && split == 1) { split = 2; pl = plan_tc(K, C, split, tiles); }     // the TF32 form needs more shared / tensor memory
  if (!pl.ok) BS_UNSUPPORTED("tensor-core feature splat: %s", pl.why);
  if (dtype == BLOBSPLAT_F64) BS_UNSUPPORTED("tensor-core feature splat: float64 runs on the FMA engine");
  RenderTcParams p{};
  p.f32_split = split;
  p.scores = scores; p.sn = sn; p.sk = sk; p.sp = sp; p.feats = feats; p.grid = out;
  if (render_tc2_usable(dtype, H, W, nullptr, out, scores, sn, sk, sp)) {        // 16-bit maps: two pixels per lane
    const Tc2Plan pl2 = plan_tc2(K, C, false, N * ((px + kTc2TilePx - 1) / kTc2TilePx));
    if (pl2.ok) return run_tc2<true>(p, pl2, N, K, H, W, C, dtype, st);
  }
  if (int rc = fill_tc_units(p, pl, N, K, H, W, C)) return rc;
  return launch_tc_dtype<true>(p, pl.smem, dtype, st);
}

// Several stage-3 problems (pyramid levels) in one launch; L.lv[*] filled by fill_tc_units, equal Kp / c_tile / dtype.
template <typename FT, typename OT, int kSplit, bool kRing>
static int launch_tc_levels_r(RenderTcLevels& L, size_t smem, cudaStream_t st) {
  static thread_local int configured_dev = -1, sm_count = 0;
  int dev = 0;
  BS_CUDA(cudaGetDevice(&dev));
  if (configured_dev != dev) {
    BS_CUDA(cudaFuncSetAttribute(render_tc_kernel<FT, OT, kSplit, 2, -1, true, kRing>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    BS_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    configured_dev = dev;
  }
  long long total = 0;
  for (int i = 0; i < L.n_levels; ++i) {
    L.tile_start[i] = (int)total;
    total += L.lv[i].total_tiles;
    L.lv[i].pair_ok = ((L.lv[i].H * L.lv[i].W) & 1) == 0 && (reinterpret_cast<uintptr_t>(L.lv[i].grid) & 7) == 0;
  }
  if (total > 0x7fffffffll) BS_UNSUPPORTED("too many tiles for one launch");
  L.tile_start[L.n_levels] = (int)total;
  if (total == 0) return 0;
  const int grid = (int)std::min<long long>(sm_count, total);
  BS_CUDA(launch_pdl(render_tc_kernel<FT, OT, kSplit, 2, -1, true, kRing>, dim3(grid), dim3((13 + (kRing ? kTcStageWarps : 0)) * 32),
                     smem, st, L));
  return 0;
}

template <typename FT, typename OT, int kSplit>
static int launch_tc_levels_t(RenderTcLevels& L, size_t smem, cudaStream_t st) {
  return L.lv[0].nb > 1 ? launch_tc_levels_r<FT, OT, kSplit, true>(L, smem, st)
                        : launch_tc_levels_r<FT, OT, kSplit, false>(L, smem, st);
}

static int launch_tc_levels(RenderTcLevels& L, size_t smem, int dtype, cudaStream_t st) {
  if (dtype == BLOBSPLAT_F32)
    return f32_split() == 1 ? launch_tc_levels_t<float, float, 1>(L, smem, st) : launch_tc_levels_t<float, float, 2>(L, smem, st);
  if (dtype == BLOBSPLAT_BF16) return launch_tc_levels_t<__nv_bfloat16, __nv_bfloat16, 0>(L, smem, st);
  if (dtype == BLOBSPLAT_F16) return launch_tc_levels_t<__half, __half, 0>(L, smem, st);
  BS_UNSUPPORTED("tensor-core engine: unsupported dtype %d", dtype);
}

// Up to kTcMaxLevels stage-3 problems of one pyramid (same N, K, dtype; per-level H, W, C) as ONE launch.  Returns
// 1 (nothing launched) when the levels cannot share a launch — different channel tiles or operand depth — so the
// caller runs them one by one.
int feature_splat_levels_tc_dispatch(int n_levels, const void* const* scores, const int64_t* sn, const int64_t* sk,
                                     const int64_t* sp, const void* const* feats, void* const* outs, int N, int K,
                                     const int* C, const int* H, const int* W, int dtype, cudaStream_t st) {
  if (n_levels < 2 || n_levels > kTcMaxLevels || dtype == BLOBSPLAT_F64) return 1;
  // 16-bit maps: the two-pixels-per-lane kernel when every level qualifies (even widths, aligned pixel-contiguous maps) and
  // the levels share Kp / channel tile / drain sub-step
  if (dtype == BLOBSPLAT_BF16 || dtype == BLOBSPLAT_F16) {
    bool ok2 = true;
    Tc2Plan first2{};
    RenderTcLevels L2{};
    for (int i = 0; i < n_levels && ok2; ++i) {
      ok2 = render_tc2_usable(dtype, H[i], W[i], nullptr, outs[i], scores[i], sn[i], sk[i], sp[i], /*any_size=*/true);
      if (!ok2) break;
      const Tc2Plan pl2 = plan_tc2(K, C[i]);
      if (!pl2.ok) { ok2 = false; break; }
      if (i == 0) first2 = pl2;
      else if (pl2.Kp != first2.Kp || pl2.c_tile != first2.c_tile || pl2.cw != first2.cw || pl2.nb != first2.nb) { ok2 = false; break; }
      RenderTcParams& p = L2.lv[i];
      p.scores = scores[i]; p.sn = sn[i]; p.sk = sk[i]; p.sp = sp[i]; p.feats = feats[i]; p.grid = outs[i];
      TcPlan base{};
      base.Kp = pl2.Kp; base.c_tile = pl2.c_tile; base.nb = pl2.nb; base.smem = pl2.smem; base.b_slot = pl2.b_slot; base.ok = true;
      if (int rc = fill_tc_units(p, base, N, K, H[i], W[i], C[i], kTc2TilePx)) return rc;
      p.cw = pl2.cw;
    }
    if (ok2) {
      L2.n_levels = n_levels;
      // a pyramid launch has many short units per CTA: always use the B ring when more than one buffer fits
      const bool ring = first2.nb > 1;
      for (int i = 0; i < n_levels; ++i) {
        L2.lv[i].nb = ring ? first2.nb : 1;
        L2.lv[i].smem_bytes = (int)(first2.smem - (size_t)(first2.nb - L2.lv[i].nb) * first2.b_slot);
      }
      if (dtype == BLOBSPLAT_BF16)
        return ring ? launch_tc2_levels<__nv_bfloat16, true>(L2, st) : launch_tc2_levels<__nv_bfloat16, false>(L2, st);
      return ring ? launch_tc2_levels<__half, true>(L2, st) : launch_tc2_levels<__half, false>(L2, st);
    }
  }
  RenderTcLevels L{};
  TcPlan first{};
  for (int i = 0; i < n_levels; ++i) {
    const TcPlan pl = plan_tc(K, C[i], split_of(dtype));
    if (!pl.ok) return 1;
    if (i == 0) first = pl;
    else if (pl.Kp != first.Kp || pl.c_tile != first.c_tile || pl.smem != first.smem) return 1;
    RenderTcParams& p = L.lv[i];
    p.scores = scores[i]; p.sn = sn[i]; p.sk = sk[i]; p.sp = sp[i]; p.feats = feats[i]; p.grid = outs[i];
    if (int rc = fill_tc_units(p, pl, N, K, H[i], W[i], C[i])) return rc;
  }
  L.n_levels = n_levels;
  // a pyramid launch has many short units per CTA: always use the B ring when more than one buffer fits
  for (int i = 0; i < n_levels; ++i) L.lv[i].nb = first.nb;
  return launch_tc_levels(L, first.smem, dtype, st);
}

}  // namespace blobsplat

#if BS_TIMING
extern "C" __attribute__((visibility("default"))) int blobsplat_debug_timing(unsigned long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, ::g_tc_timing, sizeof(unsigned long long) * 16);
}
#endif
