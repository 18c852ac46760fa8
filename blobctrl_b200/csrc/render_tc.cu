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

int render_tc_dispatch(const float* xs, const float* ys, const float* covs, const float* sizes, const void* features,
                       int feat_dtype, int N, int M, int H, int W, int C, void* composed, void* grid, int out_dtype,
                       cudaStream_t st) {
  (void)feat_dtype;
  const TcPlan pl = plan_tc(M + 1, C, split_of(out_dtype));
  if (!pl.ok) BS_UNSUPPORTED("fused render: %s", pl.why);
  RenderTcParams p{};
  p.xs = xs; p.ys = ys; p.covs = covs; p.sizes = sizes; p.feats = features; p.composed = composed; p.grid = grid;
  if (render_tc2_usable(out_dtype, H, W, composed, grid, nullptr, 0, 0, 1)) {   // 16-bit maps: two pixels per lane
    const Tc2Plan pl2 = plan_tc2(M + 1, C);
    if (pl2.ok) return run_tc2<false>(p, pl2, N, M + 1, H, W, C, out_dtype, st);
  }
  if (int rc = fill_tc_units(p, pl, N, M + 1, H, W, C)) return rc;
  return launch_tc_dtype<false>(p, pl.smem, out_dtype, st);
}

}  // namespace blobsplat
