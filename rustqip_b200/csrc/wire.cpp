// wire.cpp -- "QIPS" gate-schedule wire format behind the C ABI.
//
// The reference's only export is OpenQASM 2.0 (qip/src/qasm.rs:112-184), lossy for this path.
// QIPS carries exactly what the C ABI consumes (format: rustqip_b200/wire.py, little-endian):
//
//   u32 magic 0x53504951 | u32 version 1 | u32 prec | u32 n_qubits | u64 n_ops, then per op (recursive):
//     u8 kind | u32 n_indices | u32 n_control | u64 indices[n_indices]
//     Matrix: u64 n_entries | complex<prec>[n_entries]
//     Sparse: u64 n_rows | per row: u64 nnz | nnz x (u64 col, complex<prec> val)   -> CSR in qip_op
//     Swap:   -                     Control: one nested record
//
// Host-only code: parsing needs no GPU.  A parsed schedule owns every array its qip_op records
// point to; the records can be handed to qipb200_state_apply_schedule / qipb200_calculate_state.
#include <cstdio>
#include <cstring>
#include <deque>
#include <new>
#include <string>
#include <vector>

#include "../../include/qipb200.h"

namespace {

const uint32_t kMagic = 0x53504951u;
const uint32_t kMaxNesting = 64;       // Control(Control(...)) depth
const uint32_t kMaxIndices = 64;       // an op cannot name more qubits than an index has bits

struct Reader {
  const unsigned char *p;
  size_t left;
  bool take(void *dst, size_t n) {
    if (n > left) return false;
    memcpy(dst, p, n);
    p += n;
    left -= n;
    return true;
  }
  template <typename T>
  bool get(T *v) {
    return take(v, sizeof(T));
  }
};

}  // namespace

struct qipb200_schedule {
  uint32_t n_qubits = 0;
  int prec = QIP_F64;
  std::vector<qip_op> ops;               // top-level records
  std::deque<qip_op> nested;             // inner ops of Control records (stable addresses)
  std::deque<std::vector<uint64_t>> u64s;  // indices, CSR row pointers / columns
  std::deque<std::vector<unsigned char>> blobs;  // dense matrices, sparse values
};

namespace {

bool fail(std::string *err, const char *msg) {
  if (err) *err = msg;
  return false;
}

bool parse_op(Reader &r, qipb200_schedule *s, size_t amp, qip_op *out, uint32_t depth, std::string *err) {
  if (depth > kMaxNesting) return fail(err, "schedule: Control nesting too deep");
  uint8_t kind;
  uint32_t n_idx, n_ctrl;
  if (!r.get(&kind) || !r.get(&n_idx) || !r.get(&n_ctrl)) return fail(err, "schedule truncated (op header)");
  if (kind > QIP_OP_CONTROL) return fail(err, "schedule: unknown op kind");
  if (n_idx > kMaxIndices) return fail(err, "schedule: too many indices in one op");
  if (kind != QIP_OP_CONTROL && n_ctrl != 0) return fail(err, "schedule: control count on an op that is not a Control");
  memset(out, 0, sizeof(*out));
  out->kind = kind;
  out->n_indices = n_idx;
  out->n_control = n_ctrl;
  s->u64s.emplace_back(n_idx);
  std::vector<uint64_t> &idx = s->u64s.back();
  if (n_idx && !r.take(idx.data(), 8 * (size_t)n_idx)) return fail(err, "schedule truncated (indices)");
  out->indices = idx.data();
  switch (kind) {
    case QIP_OP_MATRIX: {
      uint64_t n_ent;
      if (!r.get(&n_ent)) return fail(err, "schedule truncated (matrix size)");
      if (n_ent > r.left / amp) return fail(err, "schedule truncated (matrix data)");
      s->blobs.emplace_back((size_t)n_ent * amp);
      r.take(s->blobs.back().data(), (size_t)n_ent * amp);
      out->n_entries = n_ent;
      out->dense = s->blobs.back().data();
      return true;
    }
    case QIP_OP_SPARSE: {
      uint64_t n_rows;
      if (!r.get(&n_rows)) return fail(err, "schedule truncated (sparse rows)");
      if (n_rows > r.left / 8) return fail(err, "schedule truncated (sparse rows)");
      if (n_idx > 20 || n_rows != (1ull << n_idx)) return fail(err, "schedule: Sparse record whose row count is not 2^n_indices");
      s->u64s.emplace_back();
      std::vector<uint64_t> &rowptr = s->u64s.back();
      rowptr.reserve((size_t)n_rows + 1);
      rowptr.push_back(0);
      s->u64s.emplace_back();
      std::vector<uint64_t> &cols = s->u64s.back();
      s->blobs.emplace_back();
      std::vector<unsigned char> &vals = s->blobs.back();
      for (uint64_t row = 0; row < n_rows; ++row) {
        uint64_t nnz;
        if (!r.get(&nnz)) return fail(err, "schedule truncated (sparse row)");
        if (nnz > r.left / (8 + amp)) return fail(err, "schedule truncated (sparse entries)");
        for (uint64_t e = 0; e < nnz; ++e) {
          uint64_t c;
          r.get(&c);
          cols.push_back(c);
          const size_t at = vals.size();
          vals.resize(at + amp);
          r.take(vals.data() + at, amp);
        }
        rowptr.push_back((uint64_t)cols.size());
      }
      out->n_entries = n_rows;
      out->sp_rowptr = rowptr.data();
      out->sp_col = cols.data();
      out->sp_val = vals.data();
      return true;
    }
    case QIP_OP_SWAP:
      return true;
    default: {  // QIP_OP_CONTROL
      s->nested.emplace_back();
      qip_op *inner = &s->nested.back();
      if (!parse_op(r, s, amp, inner, depth + 1, err)) return false;
      out->inner = inner;
      return true;
    }
  }
}

struct Writer {
  unsigned char *p;
  size_t cap, used;
  void put(const void *src, size_t n) {
    if (used + n <= cap && p) memcpy(p + used, src, n);
    used += n;
  }
  template <typename T>
  void val(T v) {
    put(&v, sizeof(T));
  }
};

bool write_op(Writer &w, const qip_op *op, size_t amp, uint32_t depth) {
  if (!op || depth > kMaxNesting || op->kind < 0 || op->kind > QIP_OP_CONTROL) return false;
  if (op->n_indices && !op->indices) return false;
  w.val<uint8_t>((uint8_t)op->kind);
  w.val<uint32_t>(op->n_indices);
  w.val<uint32_t>(op->kind == QIP_OP_CONTROL ? op->n_control : 0u);
  w.put(op->indices, 8 * (size_t)op->n_indices);
  switch (op->kind) {
    case QIP_OP_MATRIX:
      if (op->n_entries && !op->dense) return false;
      w.val<uint64_t>(op->n_entries);
      w.put(op->dense, (size_t)op->n_entries * amp);
      return true;
    case QIP_OP_SPARSE: {
      if (op->n_entries && (!op->sp_rowptr || !op->sp_col || !op->sp_val)) return false;
      w.val<uint64_t>(op->n_entries);
      for (uint64_t row = 0; row < op->n_entries; ++row) {
        const uint64_t b = op->sp_rowptr[row], e = op->sp_rowptr[row + 1];
        if (e < b) return false;
        w.val<uint64_t>(e - b);
        for (uint64_t k = b; k < e; ++k) {
          w.val<uint64_t>(op->sp_col[k]);
          w.put(static_cast<const unsigned char *>(op->sp_val) + (size_t)k * amp, amp);
        }
      }
      return true;
    }
    case QIP_OP_SWAP:
      return true;
    default:
      return write_op(w, op->inner, amp, depth + 1);
  }
}

}  // namespace

extern "C" {

int qipb200_schedule_parse(const void *buf, size_t len, qipb200_schedule **out, char *errbuf, size_t errlen) {
  if (errbuf && errlen) errbuf[0] = 0;
  if (!buf || !out) return QIPB200_ERR_INVALID_ARG;
  *out = nullptr;
  Reader r = {static_cast<const unsigned char *>(buf), len};
  std::string err;
  uint32_t magic, version, prec, n;
  uint64_t n_ops;
  qipb200_schedule *s = new (std::nothrow) qipb200_schedule();
  if (!s) return QIPB200_ERR_OOM;
  bool ok = false;
  try {  // the parser allocates through std::vector: an allocation failure is a status, never an abort
    ok = r.get(&magic) && r.get(&version) && r.get(&prec) && r.get(&n) && r.get(&n_ops);
    if (!ok)
      err = "schedule truncated (file header)";
    else if (magic != kMagic || version != 1 || prec > QIP_F64)
      ok = fail(&err, "not a QIPS version-1 schedule");
    else if (n_ops > r.left / 9)  // every record has at least a 9-byte header
      ok = fail(&err, "schedule truncated (fewer records than announced)");
    if (ok) {
      s->n_qubits = n;
      s->prec = (int)prec;
      const size_t amp = prec == QIP_F32 ? 8 : 16;
      s->ops.resize((size_t)n_ops);
      for (uint64_t i = 0; ok && i < n_ops; ++i) ok = parse_op(r, s, amp, &s->ops[(size_t)i], 0, &err);
      if (ok && r.left != 0) ok = fail(&err, "schedule: trailing bytes after the last record");
    }
  } catch (const std::bad_alloc &) {
    if (errbuf && errlen) snprintf(errbuf, errlen, "out of host memory while parsing the schedule");
    delete s;
    return QIPB200_ERR_OOM;
  }
  if (!ok) {
    if (errbuf && errlen) snprintf(errbuf, errlen, "%s", err.c_str());
    delete s;
    return QIPB200_ERR_INVALID_ARG;
  }
  *out = s;
  return QIPB200_OK;
}

const qip_op *qipb200_schedule_ops(const qipb200_schedule *s, size_t *n_ops, uint32_t *n_qubits, qip_prec *prec) {
  if (!s) return nullptr;
  if (n_ops) *n_ops = s->ops.size();
  if (n_qubits) *n_qubits = s->n_qubits;
  if (prec) *prec = (qip_prec)s->prec;
  return s->ops.data();
}

void qipb200_schedule_free(qipb200_schedule *s) { delete s; }

size_t qipb200_schedule_serialise(qip_prec prec, uint32_t n_qubits, const qip_op *ops, size_t n_ops, void *buf,
                                  size_t cap) {
  if ((prec != QIP_F32 && prec != QIP_F64) || (n_ops && !ops)) return 0;
  Writer w = {static_cast<unsigned char *>(buf), buf ? cap : 0, 0};
  w.val<uint32_t>(kMagic);
  w.val<uint32_t>(1u);
  w.val<uint32_t>((uint32_t)prec);
  w.val<uint32_t>(n_qubits);
  w.val<uint64_t>((uint64_t)n_ops);
  const size_t amp = prec == QIP_F32 ? 8 : 16;
  for (size_t i = 0; i < n_ops; ++i)
    if (!write_op(w, &ops[i], amp, 0)) return 0;
  return w.used;
}

}  // extern "C"
