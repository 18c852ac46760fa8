// Fused render (stages 1+2+3) — instantiations of render_tc.cuh with the A operand rendered from blob
// parameters.  Kernel, pipeline and layout notes: render_tc.cuh.
#include "render_tc2.cuh"

namespace blobsplat {

void render_tc_limits(int* max_k, int* c_multiple, int* max_c) { *max_k = kTcMaxBlobs + 1; *c_multiple = 1; *max_c = 1 << 20; }

int render_tc_supported(int K, int C, int H, int W, int feat_dtype, int out_dtype, const char** why) {
  *why = "";
  if (out_dtype == BLOBSPLAT_F64 || feat_dtype == BLOBSPLAT_F64) { *why = "float64 runs on the FMA engine"; return 0; }
  if (feat_dtype != out_dtype) { *why = "features and maps must share a dtype"; return 0; }
  const TcPlan pl = plan_tc(K, C, split_of(out_dtype));
  if (!pl.ok) { *why = pl.why; return 0; }
  (void)H; (void)W;
  return 1;
}

// The fused two-pixel render with up to three halvings of the composed maps in the same launch: 64 x 64 maps in 16 bits,
// levels [N,K,32,32], [N,K,16,16] and [N,K,8,8].  Returns 1 (nothing launched) when the shape is outside that envelope —
// the caller runs render + pyramid kernel.
int render_tc_pyramid_dispatch(const float* xs, const float* ys, const float* covs, const float* sizes, const void* features,
                               int N, int M, int S, int C, void* composed, void* grid, void* const* pyr, int pyr_levels,
                               int dtype, cudaStream_t st) {
  if (S != 64 || pyr_levels < 1 || pyr_levels > 3 || !composed) return 1;
  if (!render_tc2_usable(dtype, S, S, composed, grid, nullptr, 0, 0, 1)) return 1;
  for (int l = 0; l < pyr_levels; ++l)
    if (!pyr[l] || (reinterpret_cast<uintptr_t>(pyr[l]) & 15) != 0) return 1;      // bulk stores: 16-byte aligned rows
  const Tc2Plan pl2 = plan_tc2(M + 1, C, /*pyr=*/true, (long long)N * 16);
  if (!pl2.ok) return 1;
  RenderTcParams p{};
  p.xs = xs; p.ys = ys; p.covs = covs; p.sizes = sizes; p.feats = features; p.composed = composed; p.grid = grid;
  for (int l = 0; l < pyr_levels; ++l) p.pyr[l] = pyr[l];
  p.pyr_levels = pyr_levels;
  if (const char* e = std::getenv("BLOBSPLAT_PYR_DBG")) p.pyr_dbg = std::atoi(e);   // measurement only
  // the pyramid warp takes the place of the B ring's staging warps: large batches, where the ring pays, keep the
  // separate pyramid kernel (a few per cent of such a render)
  TcPlan base{};
  base.Kp = pl2.Kp; base.c_tile = pl2.c_tile; base.nb = pl2.nb; base.smem = pl2.smem; base.b_slot = pl2.b_slot; base.ok = true;
  if (int rc = fill_tc_units(p, base, N, M + 1, S, S, C, kTc2TilePx)) return rc;
  if (p.nb > 1) return 1;
  p.cw = pl2.cw;
  if (dtype == BLOBSPLAT_BF16) return launch_tc2<__nv_bfloat16, false>(p, st);
  return launch_tc2<__half, false>(p, st);
}

int render_tc_dispatch(const float* xs, const float* ys, const float* covs, const float* sizes, const void* features,
                       int feat_dtype, int N, int M, int H, int W, int C, void* composed, void* grid, int out_dtype,
                       cudaStream_t st) {
  (void)feat_dtype;
  const long long px = (long long)H * W, tiles = N * ((px + kTcTileM - 1) / kTcTileM);
  int split = split_for_launch(out_dtype, tiles);
  TcPlan pl = plan_tc(M + 1, C, split, tiles);
  if (!pl.ok && split == 1) { split = 2; pl = plan_tc(M + 1, C, split, tiles); }     // the TF32 form needs more shared / tensor memory
  if (!pl.ok) BS_UNSUPPORTED("fused render: %s", pl.why);
  RenderTcParams p{};
  p.f32_split = split;
  p.xs = xs; p.ys = ys; p.covs = covs; p.sizes = sizes; p.feats = features; p.composed = composed; p.grid = grid;
  if (render_tc2_usable(out_dtype, H, W, composed, grid, nullptr, 0, 0, 1)) {   // 16-bit maps: two pixels per lane
    const Tc2Plan pl2 = plan_tc2(M + 1, C, false, N * ((px + kTc2TilePx - 1) / kTc2TilePx));
    if (pl2.ok) return run_tc2<false>(p, pl2, N, M + 1, H, W, C, out_dtype, st);
  }
  if (int rc = fill_tc_units(p, pl, N, M + 1, H, W, C)) return rc;
  return launch_tc_dtype<false>(p, pl.smem, out_dtype, st);
}

}  // namespace blobsplat

#if BS_TIMING
// debug builds only: the stamps of the kernels instantiated in THIS translation unit (renders from blob parameters;
// g_tc_timing is per translation unit, splat_tc.cu exports its own as blobsplat_debug_timing)
extern "C" __attribute__((visibility("default"))) int blobsplat_debug_timing_render(unsigned long long* out16) {
  return (int)cudaMemcpyFromSymbol(out16, ::g_tc_timing, sizeof(unsigned long long) * 16);
}
#endif
