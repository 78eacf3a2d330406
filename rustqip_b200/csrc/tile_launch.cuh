// tile_launch.cuh -- launch interface of the fused tile pass (tile_kernel.cu).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/qipb200.h"

namespace qipb200 {

// Opt the kernels in to > 48 KiB of dynamic shared memory (once per process/device).
cudaError_t tile_pass_configure();

// Run one serialised pass (device blob: PassHeader + micro-op records) over the local state.
cudaError_t launch_tile_pass(qip_prec prec, void *psi, uint32_t n_local, uint32_t T, const unsigned char *d_blob,
                             cudaStream_t s, uint64_t *launches);

}  // namespace qipb200
