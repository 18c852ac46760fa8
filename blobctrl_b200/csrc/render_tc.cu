// placeholder until the tcgen05 render lands
#include "common.cuh"
namespace blobsplat {
void render_tc_limits(int* max_k, int* c_multiple, int* max_c) { *max_k = 0; *c_multiple = 0; *max_c = 0; }
int render_tc_supported(int, int, int, int, int, int, const char** why) { *why = "not built"; return 0; }
int render_tc_dispatch(const float*, const float*, const float*, const float*, const void*, int, int, int, int, int,
                       int, void*, void*, int, cudaStream_t) { BS_UNSUPPORTED("not built"); }
}
