// state.h -- context and device-resident state objects behind the C ABI.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <utility>
#include <vector>

#include "jit_runtime.h"
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
  uint64_t jit_launches = 0;       // tile passes among `tile_launches` that ran a generated (specialised) kernel
  std::vector<std::pair<const qipb200::JitCubin *, qipb200::JitLoaded>> jit_loaded;  // modules loaded on this device
  std::string jit_note;            // why the generated-kernel path was not taken last time (diagnostics)
  // second stream + events for the exchange that overlaps tile passes (api.cu: exchange_bits_split)
  cudaStream_t stream2 = nullptr;
  cudaEvent_t ev_pass[2] = {nullptr, nullptr};  // recorded on `stream`: the lower / upper half of the shard is final
  cudaEvent_t ev_exch[2] = {nullptr, nullptr};  // recorded on `stream2`: that half has been exchanged
  // multi-device context (qipb200_init_multi): one child context per device, owned by the parent; states created
  // on the parent are sharded over the children inside this one process (peer access instead of CUDA IPC)
  std::vector<qipb200_ctx *> children;
  qipb200_ctx *parent = nullptr;
  // one released state buffer kept for the next state of the same size (qipb200_calculate_state re-allocates
  // 2^n amplitudes per call otherwise)
  void *pool_buf = nullptr;
  size_t pool_bytes = 0;
  bool tile_configured = false;    // the tile kernels' > 48 KiB shared-memory opt-in was done on this device
  // optional per-category device timing (qipb200_profile_enable): CUDA-event pairs recorded on `stream`
  // around every fused tile pass [0] and every NVLink exchange incl. its two flag barriers [1]
  bool profile = false;
  struct ProfEvent {
    cudaEvent_t first, second;
    double weight;  // 1 per launch; 0.5 for each half of a pass launched in two halves around a migration
  };
  std::vector<ProfEvent> prof_events[2];
  std::vector<cudaEvent_t> prof_pool;
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
  void *buf = nullptr;      // 2^n_local amplitudes (sharded states: followed by a staging area of the same size)
  bool has_stage = false;   // the allocation of `buf` is 2 * bytes: [state | staging of the push exchange]
  uint32_t pair_seq = 0;        // migrations fused into a tile pass so far (value written into the per-tile flag words)
  bool halves_pending = false;  // the two halves of the shard become valid at ctx->ev_exch[0/1] (an overlapped migration)
  int send_stage = 0;       // pending exchange: 0 nothing yet, 1 its opening barrier is queued, 2 ... and the last tile
                            // pass has pushed the give-half into the partner's staging area
  uint32_t send_R = 0, send_l = 0;  // ... for this (rank bit, local bit) pair
  void *scratch = nullptr;  // same size, lazily allocated (out-of-place row kernel only)
  size_t bytes = 0;
  // logical index bit b (= n-1-q) -> physical index bit.  Physical bits >= n_local are
  // rank bits.  Identity unless an exchange migrated a qubit (multi-GPU only).
  std::vector<uint32_t> phys_of_logical;
  // CUDA-IPC peer mappings (multi-GPU)
  std::vector<void *> peer_buf;
  uint32_t *flags = nullptr;  // world slots, written by peers
  std::vector<uint32_t *> peer_flags;
  std::vector<const double *> peer_comm;  // every rank's reduction slot (behind its flag page, dist.cuh)
  std::vector<qipb200_state *> shards;    // parent state of a multi-device context: one sharded state per device
  bool ipc_mapped = false;                // peers were mapped with cudaIpcOpenMemHandle (to be closed on free)
  uint32_t epoch = 0;
  uint64_t exchange_bytes = 0;
  bool ipc_ready = false;
};

namespace qipb200 {

// RAII bracket of one profiled region (no-op unless ctx->profile).
struct ProfileScope {
  qipb200_ctx *ctx;
  int cat;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double weight = 1.0;
  ProfileScope(qipb200_ctx *c, int category, double w = 1.0) : ctx(c), cat(category), weight(w) {
    if (!ctx->profile) return;
    auto take = [&]() {
      cudaEvent_t e = nullptr;
      if (!ctx->prof_pool.empty()) {
        e = ctx->prof_pool.back();
        ctx->prof_pool.pop_back();
      } else if (cudaEventCreate(&e) != cudaSuccess) {
        e = nullptr;
      }
      return e;
    };
    e0 = take();
    e1 = take();
    if (e0) cudaEventRecord(e0, ctx->stream);
  }
  ~ProfileScope() {
    if (!e0 || !e1) return;
    cudaEventRecord(e1, ctx->stream);
    qipb200_ctx::ProfEvent pe = {e0, e1, weight};
    ctx->prof_events[cat].push_back(pe);
  }
};

inline size_t amp_bytes(qip_prec p) { return p == QIP_F32 ? 8 : 16; }

}  // namespace qipb200
