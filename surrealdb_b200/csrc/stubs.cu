// stubs.cu -- entry points whose kernels are not implemented yet.
#include "internal.cuh"
using namespace sdb;
extern "C" {
sdb_status sdb_hnsw_load(sdb_ctx*, uint32_t, sdb_metric, uint64_t, const float*, uint32_t, const uint64_t* const*,
                         const uint32_t* const*, int64_t, sdb_hnsw**) { set_error("hnsw: not implemented"); return SDB_EUNSUPPORTED; }
void sdb_hnsw_destroy(sdb_hnsw*) {}
sdb_status sdb_hnsw_search(sdb_hnsw*, const float*, uint32_t, uint32_t, uint32_t, uint64_t*, double*, uint32_t*, uint64_t*) { set_error("hnsw: not implemented"); return SDB_EUNSUPPORTED; }
sdb_status sdb_graph_load_csr(sdb_ctx*, uint64_t, const uint64_t*, const uint32_t*, sdb_graph**) { set_error("graph: not implemented"); return SDB_EUNSUPPORTED; }
void sdb_graph_destroy(sdb_graph*) {}
sdb_status sdb_graph_expand(sdb_graph* const*, uint32_t, const uint32_t*, uint64_t, uint32_t, uint32_t**, uint64_t*) { set_error("graph: not implemented"); return SDB_EUNSUPPORTED; }
sdb_status sdb_graph_collect(sdb_graph*, const uint32_t*, uint64_t, uint32_t, uint32_t, int, uint32_t**, uint64_t*) { set_error("graph: not implemented"); return SDB_EUNSUPPORTED; }
}
