// state.h -- context and device-resident state objects behind the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "opcompile.h"

struct qipb200_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  uint64_t tile_launches = 0;      // fused tile passes among `launches`
  uint64_t exchange_launches = 0;  // NVLink pair-exchange kernels among `launches`
  uint64_t fused_gates = 0;        // reference ops folded into tile passes
  // staging buffers of the host-buffer drop-ins (qipb200_apply_op*)
  void *d_in = nullptr, *d_out = nullptr;
  size_t d_in_bytes = 0, d_out_bytes = 0;
  double *d_scalar = nullptr;   // one device double for reductions
  void *h_pinned = nullptr;     // pinned bounce buffer for downloads/uploads
  size_t h_pinned_bytes = 0;
};

struct qipb200_state {
  qipb200_ctx *ctx = nullptr;
  qip_prec prec = QIP_F64;
  uint32_t n = 0;        // qubits of the whole state
  uint32_t n_local = 0;  // index bits held by this rank
  int rank = 0, world = 1;
  void *buf = nullptr;      // 2^n_local amplitudes
  void *scratch = nullptr;  // same size, lazily allocated (out-of-place row kernel only)
  size_t bytes = 0;
  // logical index bit b (= n-1-q) -> physical index bit.  Physical bits >= n_local are
  // rank bits.  Identity unless an exchange migrated a qubit (multi-GPU only).
  std::vector<uint32_t> phys_of_logical;
  // CUDA-IPC peer mappings (multi-GPU)
  std::vector<void *> peer_buf;
  uint32_t *flags = nullptr;  // world slots, written by peers
  std::vector<uint32_t *> peer_flags;
  uint32_t epoch = 0;
  uint64_t exchange_bytes = 0;
  bool ipc_ready = false;
};

namespace qipb200 {

inline size_t amp_bytes(qip_prec p) { return p == QIP_F32 ? 8 : 16; }

}  // namespace qipb200
