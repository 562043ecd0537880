// tcgen05 3xTF32 GEMM tile (engine 2 of satk_gemm) — placeholder until the TMA/TMEM kernel lands:
// reports "unsupported" so engine 0 (auto) routes to the fp32 SIMT tile.
#include "common.cuh"
namespace satk {
int gemm_tc_launch(const satk_gemm_desc* d, cudaStream_t st, bool* supported) {
  (void)d; (void)st;
  *supported = false;
  return SATK_OK;
}
}  // namespace satk
