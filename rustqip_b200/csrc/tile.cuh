// tile.cuh -- fused shared-memory tile pass: data structures shared by the host
// planner (planner.cpp) and the kernel (tile_kernel.cu).
//
// One PASS sweeps the state once: every CTA loads a TILE of 2^T amplitudes (the L
// lowest index bits, contiguous in HBM, plus m = T-L arbitrary higher "tile bits")
// into shared memory, applies a list of MICRO-OPS there, and writes the tile back.
// Any gate whose non-diagonal target bits are tile bits can run inside the pass;
// control bits and diagonal gates may sit on ANY bit (a non-tile bit is a CTA-uniform
// predicate).  Algorithmic HBM traffic of a pass = 2 * 2^n * sizeof(amplitude), the
// same as ONE gate of the reference's per-entry loop (qip/src/builder.rs:423-514).
#pragma once

#include <cstdint>
#include <vector>

#include "opcompile.h"

namespace qipb200 {

static const uint32_t kTileMaxHigh = 8;    // m <= 8 -> 256 chunk offsets
static const uint32_t kTileStageBytes = 1280;  // per staged micro-op: header + matrix / diag terms
static const uint32_t kMaxDiagTerms = 24;  // per DIAG micro-op (24 * 48 B = 1152 B for f64)

enum MicroKind { MK_DENSE = 0, MK_DIAG = 1, MK_EXCH = 2 };

// Device-visible micro-op header (fixed 128 bytes), followed in the blob by its data:
//   MK_DENSE: 2^k x 2^k complex<R> (re,im interleaved), sub-index bit i <-> lbit[i]
//   MK_DIAG : nterms x DiagTerm<R>
struct alignas(16) MicroOp {
  uint32_t kind;
  uint32_t k;            // dense: number of target bits (1..3)
  uint32_t ins_n;        // number of tile-local positions removed from the group counter
  uint32_t ins_pos[6];   // ascending tile-local positions (targets and local controls)
  uint32_t lor_mask;     // tile-local control bits (forced to 1)
  uint32_t off[8];       // dense: tile-local offset of sub-index u; exch: off[0] <-> off[1]
  uint32_t groups_log2;  // T - ins_n
  uint32_t nterms;       // diag
  uint32_t data_bytes;   // bytes of data following the header in the blob
  uint32_t pad0;
  uint64_t gmask;        // control bits outside the tile: tested against the tile's base index
  uint64_t pad1[3];
};
static_assert(sizeof(MicroOp) == 128, "MicroOp must be 128 bytes");

template <typename R>
struct alignas(16) DiagTerm {  // multiply by (re,im) where (global & gmask)==gval and (local & lmask)==lval
  uint64_t gmask, gval;
  uint32_t lmask, lval;
  R re, im;
};

struct PassHeader {
  uint32_t T, L, m, n_ops;
  uint32_t hi_pos[kTileMaxHigh];           // the m high tile bit positions, ascending
  uint64_t chunk_off[1u << kTileMaxHigh];  // amplitude offset of chunk c (bits of c spread over hi_pos)
  uint32_t blob_bytes;                     // bytes of micro-op records after the header
  uint32_t pad[3];
};

// ---- host side ---------------------------------------------------------------------
struct HostMicroOp {
  MicroOp h;
  std::vector<unsigned char> data;
};

struct HostPass {
  PassHeader hdr;
  std::vector<HostMicroOp> ops;
  uint32_t n_gates = 0;  // reference ops folded into this pass
};

// One step of a planned schedule: either a fused pass or a single op run by the
// per-gate kernels (index into the original op list).
struct PlanStep {
  bool is_pass = false;
  HostPass pass;
  size_t op_index = 0;
};

struct PlanConfig {
  uint32_t T = 12;        // tile bits
  uint32_t L = 5;         // contiguous low bits
  uint32_t max_block_k = 3;  // host-side gate fusion into dense blocks of <= this many bits
  bool fuse_blocks = true;
};

// Plan `ops` (already compiled against the current layout and restricted to local bits;
// ops[i].cls == CLASS_IDENTITY entries are dropped) for a local state of n_local bits.
void plan_passes(const std::vector<FlatOp> &ops, uint32_t n_local, qip_prec prec, const PlanConfig &cfg,
                 std::vector<PlanStep> *steps);

// Serialise a pass for the device (header + 128-byte micro-op records with data).
void serialise_pass(const HostPass &p, std::vector<unsigned char> *blob);

// Tile geometry defaults per precision.
PlanConfig default_plan_config(qip_prec prec, uint32_t n_local);

}  // namespace qipb200
