// placeholder until the tcgen05 path lands (see gemm_tc.cuh)
#include "flame.cuh"
namespace msmd {
int flame_decode_tc(msmd_flame* fh, int64_t B, float* verts_out, cudaStream_t st) {
  set_error("flame tensor-core path not built");
  return MSMD_ERR_UNSUPPORTED;
}
void flame_tc_destroy(msmd_flame* fh) {}
}
