// Stand-alone stage 3 on tensor cores — instantiations of render_tc.cuh with the A operand loaded from
// precomputed score maps.  Replaces splat_features_from_scores (blobctrl/utils/utils.py:57-77; duplicate at
// blobctrl/pipelines/pipeline_blobnet.py:706-721) when K and C make it a real dense contraction
// (K >= 12, C >= 64, C % 32 == 0); smaller shapes stay on the FMA engine (feature_splat.cu).
#include "render_tc.cuh"

namespace blobsplat {

int feature_splat_tc_dispatch(const void* scores, int64_t sn, int64_t sk, int64_t sp, const void* feats, void* out,
                              int N, int K, int C, int H, int W, int dtype, cudaStream_t st) {
  const TcPlan pl = plan_tc(K, C, dtype == BLOBSPLAT_F32);
  if (!pl.ok) BS_UNSUPPORTED("tensor-core feature splat: %s", pl.why);
  if (dtype == BLOBSPLAT_F64) BS_UNSUPPORTED("tensor-core feature splat: float64 runs on the FMA engine");
  RenderTcParams p{};
  p.scores = scores; p.sn = sn; p.sk = sk; p.sp = sp; p.feats = feats; p.grid = out;
  if (int rc = fill_tc_units(p, pl, N, K, H, W, C)) return rc;
  return launch_tc_dtype<true>(p, pl.smem, dtype, st);
}

}  // namespace blobsplat
