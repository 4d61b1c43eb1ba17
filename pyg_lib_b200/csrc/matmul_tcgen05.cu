// placeholder until the tcgen05 kernel lands (next commit): nothing is routed here yet.
#include "common.cuh"
namespace pygb200 {
bool tcgen05_supported(i64, i64, i64, i64, int, const void*, const void*, const void*) { return false; }
int segment_matmul_tcgen05(const void*, const i64*, const void*, const void*, void*, i64, i64, i64, i64, int, cudaStream_t) {
  set_error("tcgen05 path not built");
  return PYGB200_ERR_UNSUPPORTED;
}
}  // namespace pygb200
