// screen_tc.cu -- K2: tcgen05 bf16 GEMM screen (placeholder until the kernel lands).
#include "internal.cuh"
namespace sdb {
bool screen_tc_available() { return false; }
sdb_status screen_tc_pass(Corpus*, uint32_t, const PassDesc&, cudaStream_t) {
  set_error("tcgen05 screen not built");
  return SDB_EUNSUPPORTED;
}
}  // namespace sdb
