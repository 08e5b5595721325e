// TEMPORARY: entry points not implemented yet return CTR_ERR_UNSUPPORTED (removed as kernels land).
#include "ctr_common.cuh"
#define STUB(name, ...) extern "C" int name(__VA_ARGS__) { ctr::set_error(#name ": not implemented yet"); return CTR_ERR_UNSUPPORTED; }
STUB(ctr_cin_fwd, const float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, float*, float*, int, void*)
STUB(ctr_cin_bwd, const float*, const float*, const float*, const float*, int64_t, int64_t, int64_t, int64_t, int64_t, float*, float*, float*, void*, int64_t, void*)
extern "C" int64_t ctr_cin_bwd_workspace_bytes(int64_t, int64_t, int64_t, int64_t, int64_t) { return 0; }
