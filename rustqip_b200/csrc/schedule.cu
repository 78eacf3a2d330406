// schedule.cu -- executes a list of ops on a device-resident state.
//
// Reference: the sequential fold over the pipeline in
// LocalBuilder::calculate_state_with_init (qip/src/builder.rs:423-514): one
// apply_op_overwrite sweep per entry.  Here
//   * QIPB200_SCHED_NO_FUSION: every entry is compiled to one in-place per-gate kernel;
//   * default: runs of entries are planned into fused shared-memory tile passes
//     (planner.cpp / tile_kernel.cu), each pass one HBM sweep for many gates.
// On a sharded state a look-ahead over the schedule picks which local qubit to evict
// when a rank-held qubit has to be migrated; an exchange is a fusion barrier.
#include "schedule.h"

#include <string>
#include <vector>

#include "../../include/qipb200.h"
#include "tile.cuh"
#include "tile_launch.cuh"

namespace qipb200 {

static const uint64_t kNever = ~0ull >> 1;

namespace {

// Plan and run a batch of ops that are all local under the current layout.
int flush_fused(qipb200_state *s, std::vector<FlatOp> *pending) {
  if (pending->empty()) return QIPB200_OK;
  qipb200_ctx *ctx = s->ctx;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = tile_pass_configure();
    if (e != cudaSuccess) return report_cuda_error(s, e, "cudaFuncSetAttribute(tile pass)");
    configured = true;
  }
  const PlanConfig cfg = default_plan_config(s->prec, s->n_local);
  std::vector<PlanStep> steps;
  plan_passes(*pending, s->n_local, s->prec, cfg, &steps);
  int st = QIPB200_OK;
  PassParams *pp = new PassParams();
  for (size_t i = 0; i < steps.size() && st == QIPB200_OK; ++i) {
    if (steps[i].is_pass) {
      if (!serialise_pass(steps[i].pass, pp)) {
        st = report_error(s, QIPB200_ERR_UNSUPPORTED, "internal: fused pass exceeds the kernel parameter space");
        break;
      }
      cudaError_t e = launch_tile_pass(s->prec, s->buf, s->n_local, *pp, cfg.groups_per_thread, ctx->stream, &ctx->launches);
      if (e != cudaSuccess) st = report_cuda_error(s, e, "launch_tile_pass");
      ++ctx->tile_launches;
      ctx->fused_gates += steps[i].pass.n_gates;
    } else {
      st = launch_local_op(s, (*pending)[steps[i].op_index]);
    }
  }
  delete pp;
  pending->clear();
  return st;
}

bool needs_exchange(const qipb200_state *s, const FlatOp &f) {
  if (s->world == 1) return false;
  if (f.cls == CLASS_BITSWAP && f.ctrl_mask == 0) return true;  // handled as a relabelling by compile_and_localize
  const uint32_t nl = s->n_local;
  if (f.cls == CLASS_DENSE || f.cls == CLASS_FLIP) {
    for (size_t j = 0; j < f.tgt_sorted.size(); ++j)
      if (f.tgt_sorted[j] >= nl) return true;
  } else if (f.cls == CLASS_BITSWAP) {
    for (size_t j = 0; j < f.swaps.size(); ++j)
      if (f.swaps[j].first >= nl || f.swaps[j].second >= nl) return true;
  } else if (f.cls == CLASS_GENERAL) {
    for (uint32_t j = f.nc; j < f.k; ++j)
      if (f.idx_bits[j] >= nl) return true;
  }
  return false;
}

}  // namespace

int run_schedule(qipb200_state *s, const qip_op *ops, size_t n_ops, uint32_t flags) {
  std::vector<uint64_t> next_use;  // [i * n + logical_bit]
  if (s->world > 1 && n_ops) {
    // next non-diagonal use of every logical bit after op i (backward scan)
    next_use.assign((n_ops + 1) * s->n, kNever);
    for (size_t i = n_ops; i-- > 0;) {
      for (uint32_t b = 0; b < s->n; ++b) next_use[i * s->n + b] = next_use[(i + 1) * s->n + b];
      FlatOp f;
      std::string err;
      if (compile_op(&ops[i], s->prec, s->n, &f, &err) != QIPB200_OK) continue;  // reported when executed
      if (f.cls == CLASS_DENSE || f.cls == CLASS_FLIP)
        for (size_t j = 0; j < f.tgt_sorted.size(); ++j) next_use[i * s->n + f.tgt_sorted[j]] = i;
      else if (f.cls == CLASS_GENERAL)
        for (uint32_t j = f.nc; j < f.k; ++j) next_use[i * s->n + f.idx_bits[j]] = i;
    }
  }
  const bool fuse = !(flags & QIPB200_SCHED_NO_FUSION) && s->n_local >= 6;
  std::vector<FlatOp> pending;
  for (size_t i = 0; i < n_ops; ++i) {
    FlatOp f;
    const uint64_t *nu = next_use.empty() ? nullptr : &next_use[(i + 1) * s->n];
    if (fuse) {
      // peek: compile under the current layout; an op that needs an exchange is a barrier
      std::string err;
      int st = compile_op(&ops[i], s->prec, s->n, &f, &err, s->phys_of_logical.data());
      if (st != QIPB200_OK) {
        flush_fused(s, &pending);
        return report_error(s, st, err);
      }
      if (needs_exchange(s, f)) {
        if ((st = flush_fused(s, &pending)) != QIPB200_OK) return st;
        if ((st = compile_and_localize(s, &ops[i], &f, nu)) != QIPB200_OK) return st;
      }
      FlatOp local;
      bool skip = false;
      if ((st = restrict_to_rank(s, f, &local, &skip)) != QIPB200_OK) return st;
      if (!skip) pending.push_back(local);
    } else {
      int st = compile_and_localize(s, &ops[i], &f, nu);
      if (st != QIPB200_OK) return st;
      st = apply_flat_local(s, f);
      if (st != QIPB200_OK) return st;
    }
  }
  return flush_fused(s, &pending);
}

}  // namespace qipb200
