// schedule.cu -- executes a list of ops on a device-resident state.
//
// Reference: the sequential fold over the pipeline in
// LocalBuilder::calculate_state_with_init (qip/src/builder.rs:423-514): one
// apply_op_overwrite sweep per entry.  Here every entry is compiled to an in-place
// kernel; on a sharded state a look-ahead over the schedule picks which local
// qubit to evict when a rank-held qubit has to be migrated.
#include "schedule.h"

#include <string>
#include <vector>

#include "../../include/qipb200.h"

namespace qipb200 {

static const uint64_t kNever = ~0ull >> 1;

int run_schedule(qipb200_state *s, const qip_op *ops, size_t n_ops, uint32_t flags) {
  (void)flags;
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
  for (size_t i = 0; i < n_ops; ++i) {
    FlatOp f;
    const uint64_t *nu = next_use.empty() ? nullptr : &next_use[(i + 1) * s->n];
    int st = compile_and_localize(s, &ops[i], &f, nu);
    if (st != QIPB200_OK) return st;
    st = apply_flat_local(s, f);
    if (st != QIPB200_OK) return st;
  }
  return QIPB200_OK;
}

}  // namespace qipb200
