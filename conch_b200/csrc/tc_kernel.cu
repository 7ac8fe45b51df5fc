// tcgen05 prefill kernel -- placeholder until the pipeline lands.
#include "common.cuh"

namespace cb200 {
TcTuning g_tc_tuning = {0, 0};
bool tc_supported(const Problem&, int) { return false; }
size_t tc_workspace_bytes(int64_t, int64_t, int64_t) { return 0; }
int launch_tc(const Problem&, int, int, int, void*, size_t, cudaStream_t, int*) {
  set_error("tcgen05 path not built");
  return CB200_ERR_UNSUPPORTED;
}
}  // namespace cb200
