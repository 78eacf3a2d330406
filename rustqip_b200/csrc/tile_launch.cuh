// tile_launch.cuh -- launch interface of the fused tile pass (tile_kernel.cu).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/qipb200.h"
#include "tile.cuh"

namespace qipb200 {

// Opt the kernels in to > 48 KiB of dynamic shared memory (once per process/device).
cudaError_t tile_pass_configure();

// Tensor map of the local state for one pass's tile geometry (5-D, 128-byte swizzle: see tile_kernel.cu).
// Returns false when the geometry does not fit a map (tiny states): no TMA, no generated kernel.
bool make_tile_map(CUtensorMap *map, qip_prec prec, void *psi, uint32_t n_local, const PassHeader &h);

// Run one serialised pass (passed by value as a kernel parameter) over the local state.
// groups_per_thread: 1 (3 CTAs/SM) or 2 (2 CTAs/SM, descriptors decoded once per two groups).
// use_tma: move the tile with TMA tensor copies when the pass geometry allows (m >= 3).
cudaError_t launch_tile_pass(qip_prec prec, void *psi, uint32_t n_local, PassParams &pp, int groups_per_thread,
                             bool use_tma, cudaStream_t s, uint64_t *launches);

}  // namespace qipb200
