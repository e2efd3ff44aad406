// voxelgrid.cuh — STUB (replaced below in this round)
#pragma once
#include "engine.cuh"
namespace b2r {
struct VoxelWork { void release() {} };
inline int voxelgrid_filter(VoxelWork&, cudaStream_t, const void*, size_t, size_t, float, void*, size_t*, int32_t*, int32_t*) { return fail(B2R_EUNSUPPORTED, "voxelgrid stub"); }
}
