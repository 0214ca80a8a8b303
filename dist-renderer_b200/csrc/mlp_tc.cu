// tcgen05 tensor-core engine (DIST_ENGINE_TC) -- placeholder until the UMMA tile kernel lands.
#include "common.cuh"
namespace dist {
int mlp_tc_launch(const dist_net_t*, const NetDev&, int, const MlpArgs&, cudaStream_t) {
  set_error("tensor-core engine not built in this library");
  return DIST_E_UNSUPPORTED;
}
}  // namespace dist
