// ndt.cuh — STUB (replaced below in this round)
#pragma once
#include "engine.cuh"
namespace b2r {
struct NdtVoxelMap {};
struct NdtWork { void release() {} };
inline void ndt_free_map(NdtVoxelMap*) {}
inline int ndt_ensure_map(const b2r_config&, Cloud&, NdtWork&, cudaStream_t) { return fail(B2R_EUNSUPPORTED, "ndt stub"); }
inline int ndt_dump(Cloud&, cudaStream_t, size_t, size_t*, int64_t*, int32_t*, double*, double*, int32_t*, int32_t*) { return fail(B2R_EUNSUPPORTED, "ndt stub"); }
inline int ndt_derivatives_at(const b2r_config&, Cloud&, Cloud&, NdtWork&, cudaStream_t, const double*, double*, double*, double*, uint64_t*) { return fail(B2R_EUNSUPPORTED, "ndt stub"); }
inline int ndt_align(const b2r_config&, Cloud&, Cloud&, NdtWork&, cudaStream_t, const float*, float*, bool*, int*) { return fail(B2R_EUNSUPPORTED, "ndt stub"); }
}
