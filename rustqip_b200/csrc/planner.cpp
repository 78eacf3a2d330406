// planner.cpp -- host-side fusion planner: turns a list of compiled local ops into
// fused shared-memory tile passes (tile.cuh) and single-op steps.  Pure C++.
//
// This is the "fusion scheduler" row of SURVEY.md section 8f (N4): what the reference's
// apply_ops (qip-iterators/src/matrix_ops.rs:158-219) aspired to.  The fused result
// is DEFINED as the sequential product of the single-op semantics.
//
// Reordering rule.  Every op acts on each of its index bits either DIAGONALLY (a
// control bit, or a bit of a diagonal gate) or NON-DIAGONALLY (a target of a dense
// block / X / Swap).  Two ops commute if on every shared bit both act diagonally.
// A pass takes ops in program order, skipping an op only when it does not fit; an
// op may overtake a skipped one only if the two commute by the rule above.
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "tile.cuh"

namespace qipb200 {

namespace {

inline int popc(uint64_t x) { return __builtin_popcountll(x); }

struct OpInfo {
  uint64_t nd = 0;        // bits acted on non-diagonally
  uint64_t dg = 0;        // bits acted on diagonally
  uint64_t need_tile = 0; // bits that must be tile bits for the op to run in a pass
  bool tile_ok = false;
  double unfused_cost = 1.0;  // HBM sweeps of the per-gate kernel
};

OpInfo analyse(const FlatOp &f, const PlanConfig &cfg) {
  OpInfo o;
  const int nctrl = popc(f.ctrl_mask);
  switch (f.cls) {
    case CLASS_DENSE:
      for (size_t i = 0; i < f.tgt_sorted.size(); ++i) o.nd |= 1ull << f.tgt_sorted[i];
      o.dg = f.ctrl_mask;
      o.need_tile = o.nd;
      o.tile_ok = f.tgt_sorted.size() <= 3 && f.tgt_sorted.size() + nctrl <= 6;
      o.unfused_cost = 1.0 / (1 << nctrl);
      break;
    case CLASS_FLIP:
      o.nd = 1ull << f.tgt_sorted[0];
      o.dg = f.ctrl_mask;
      o.need_tile = o.nd;
      o.tile_ok = nctrl + 1 <= 6;
      o.unfused_cost = 1.0 / (1 << nctrl);
      break;
    case CLASS_BITSWAP:
      for (size_t i = 0; i < f.swaps.size(); ++i) o.nd |= (1ull << f.swaps[i].first) | (1ull << f.swaps[i].second);
      o.dg = f.ctrl_mask;
      o.need_tile = o.nd;
      o.tile_ok = nctrl + 2 <= 6 && f.swaps.size() <= 3;
      o.unfused_cost = 0.5 * f.swaps.size() / (1 << nctrl);
      break;
    case CLASS_DIAGONAL:
      o.dg = f.ctrl_mask;
      for (size_t i = 0; i < f.diag_bits.size(); ++i) o.dg |= 1ull << f.diag_bits[i];
      o.tile_ok = f.diag_bits.size() <= 3;  // expands to <= 8 masked-phase terms
      o.unfused_cost = 1.0 / (1 << nctrl);
      break;
    default:  // CLASS_GENERAL (CLASS_IDENTITY never reaches here)
      for (uint32_t j = 0; j < f.k; ++j) o.nd |= 1ull << f.idx_bits[j];
      o.tile_ok = false;
      break;
  }
  (void)cfg;
  return o;
}

// ---- small dense linear algebra on the host (block fusion) ----------------------------
struct Block {
  std::vector<uint32_t> bits;  // tile-local bit positions, ascending; sub-index bit i <-> bits[i]
  std::vector<cplx> m;         // 2^k x 2^k row-major
};

// Embed a matrix on `from` bits (ascending) into the sub-space of `to` bits (ascending, superset).
std::vector<cplx> embed(const std::vector<cplx> &m, const std::vector<uint32_t> &from,
                        const std::vector<uint32_t> &to) {
  const size_t kf = from.size(), kt = to.size();
  std::vector<int> pos(kf);
  for (size_t i = 0; i < kf; ++i) pos[i] = (int)(std::find(to.begin(), to.end(), from[i]) - to.begin());
  const size_t St = 1u << kt, Sf = 1u << kf;
  std::vector<cplx> out(St * St, cplx(0, 0));
  uint32_t from_mask = 0;
  for (size_t i = 0; i < kf; ++i) from_mask |= 1u << pos[i];
  for (size_t r = 0; r < St; ++r)
    for (size_t c = 0; c < St; ++c) {
      if ((r & ~from_mask) != (c & ~from_mask)) continue;  // identity on the other bits
      size_t rs = 0, cs = 0;
      for (size_t i = 0; i < kf; ++i) {
        rs |= ((r >> pos[i]) & 1) << i;
        cs |= ((c >> pos[i]) & 1) << i;
      }
      out[r * St + c] = m[rs * Sf + cs];
    }
  return out;
}

std::vector<cplx> matmul(const std::vector<cplx> &a, const std::vector<cplx> &b, size_t S) {
  std::vector<cplx> out(S * S, cplx(0, 0));
  for (size_t r = 0; r < S; ++r)
    for (size_t k = 0; k < S; ++k) {
      const cplx av = a[r * S + k];
      if (av.real() == 0.0 && av.imag() == 0.0) continue;
      for (size_t c = 0; c < S; ++c) out[r * S + c] += av * b[k * S + c];
    }
  return out;
}

// Dense matrix (on the op's own target+local-control bits) of an op whose every bit is tile-local.
// Returns false if the op cannot be expressed as a small host matrix (e.g. it has non-tile controls).
struct LocalOp {
  // a compiled op translated to tile-local coordinates
  const FlatOp *f = nullptr;
  uint64_t gmask = 0;                   // controls on non-tile bits
  uint32_t lctrl = 0;                   // controls on tile-local bits
  std::vector<uint32_t> ltgt;           // dense/flip targets (tile-local, ascending)
  std::vector<std::pair<uint32_t, uint32_t>> lswaps;
};

bool as_block(const LocalOp &lo, size_t max_k, Block *out) {
  const FlatOp &f = *lo.f;
  if (lo.gmask) return false;
  if (f.cls == CLASS_DIAGONAL) return false;  // handled as phase terms (may involve non-tile bits)
  std::vector<uint32_t> bits;
  for (uint32_t b = 0; b < 32; ++b)
    if ((lo.lctrl >> b) & 1) bits.push_back(b);
  std::vector<uint32_t> tg;
  if (f.cls == CLASS_BITSWAP) {
    for (size_t i = 0; i < lo.lswaps.size(); ++i) {
      tg.push_back(lo.lswaps[i].first);
      tg.push_back(lo.lswaps[i].second);
    }
  } else {
    tg = lo.ltgt;
  }
  for (size_t i = 0; i < tg.size(); ++i) bits.push_back(tg[i]);
  std::sort(bits.begin(), bits.end());
  if (bits.size() > max_k) return false;
  const size_t k = bits.size(), S = 1u << k;
  // inner matrix on the target bits
  std::vector<uint32_t> tsorted = tg;
  std::sort(tsorted.begin(), tsorted.end());
  const size_t kt = tsorted.size(), St = 1u << kt;
  std::vector<cplx> inner(St * St, cplx(0, 0));
  if (f.cls == CLASS_DENSE) {
    inner = f.m_sorted;  // already in ascending-bit order
  } else if (f.cls == CLASS_FLIP) {
    inner = {cplx(0, 0), cplx(1, 0), cplx(1, 0), cplx(0, 0)};
  } else {  // BITSWAP: permutation exchanging each pair of bits
    for (size_t c = 0; c < St; ++c) {
      size_t r = c;
      for (size_t i = 0; i < lo.lswaps.size(); ++i) {
        const int p = (int)(std::find(tsorted.begin(), tsorted.end(), lo.lswaps[i].first) - tsorted.begin());
        const int q = (int)(std::find(tsorted.begin(), tsorted.end(), lo.lswaps[i].second) - tsorted.begin());
        const size_t bp = (r >> p) & 1, bq = (r >> q) & 1;
        r = (r & ~((size_t)1 << p) & ~((size_t)1 << q)) | (bq << p) | (bp << q);
      }
      inner[r * St + c] = cplx(1, 0);
    }
  }
  std::vector<cplx> full = embed(inner, tsorted, bits);
  // controls: identity rows/cols where any local control bit is 0
  uint32_t cmask = 0;
  for (size_t i = 0; i < k; ++i)
    if ((lo.lctrl >> bits[i]) & 1) cmask |= 1u << i;
  if (cmask) {
    for (size_t r = 0; r < S; ++r)
      for (size_t c = 0; c < S; ++c)
        if ((r & cmask) != cmask || (c & cmask) != cmask) full[r * S + c] = (r == c) ? cplx(1, 0) : cplx(0, 0);
  }
  out->bits = bits;
  out->m = full;
  return true;
}

template <typename R>
void push_dense(const Block &b, HostPass *pass, uint32_t T) {
  HostMicroOp mo;
  memset(&mo.h, 0, sizeof(mo.h));
  mo.h.kind = MK_DENSE;
  mo.h.k = (uint32_t)b.bits.size();
  mo.h.ins_n = mo.h.k;
  for (uint32_t i = 0; i < mo.h.k; ++i) mo.h.ins_pos[i] = b.bits[i];
  const uint32_t S = 1u << mo.h.k;
  for (uint32_t u = 0; u < S; ++u) {
    uint32_t off = 0;
    for (uint32_t i = 0; i < mo.h.k; ++i)
      if ((u >> i) & 1) off |= 1u << b.bits[i];
    mo.h.off[u] = off;
  }
  mo.h.groups_log2 = T - mo.h.ins_n;
  mo.data.resize((size_t)S * S * 2 * sizeof(R));
  R *d = reinterpret_cast<R *>(mo.data.data());
  for (uint32_t i = 0; i < S * S; ++i) {
    d[2 * i] = (R)b.m[i].real();
    d[2 * i + 1] = (R)b.m[i].imag();
  }
  mo.h.data_bytes = (uint32_t)mo.data.size();
  pass->ops.push_back(mo);
}

// A dense / flip op that keeps its controls as predicates (non-tile controls and/or too many bits).
template <typename R>
void push_controlled_dense(const LocalOp &lo, HostPass *pass, uint32_t T) {
  const FlatOp &f = *lo.f;
  HostMicroOp mo;
  memset(&mo.h, 0, sizeof(mo.h));
  std::vector<uint32_t> ins = lo.ltgt;
  for (uint32_t b = 0; b < 32; ++b)
    if ((lo.lctrl >> b) & 1) ins.push_back(b);
  std::sort(ins.begin(), ins.end());
  mo.h.ins_n = (uint32_t)ins.size();
  for (size_t i = 0; i < ins.size(); ++i) mo.h.ins_pos[i] = ins[i];
  mo.h.lor_mask = lo.lctrl;
  mo.h.gmask = lo.gmask;
  mo.h.groups_log2 = T - mo.h.ins_n;
  if (f.cls == CLASS_FLIP) {
    mo.h.kind = MK_EXCH;
    mo.h.off[0] = 0;
    mo.h.off[1] = 1u << lo.ltgt[0];
  } else {
    mo.h.kind = MK_DENSE;
    mo.h.k = (uint32_t)lo.ltgt.size();
    const uint32_t S = 1u << mo.h.k;
    for (uint32_t u = 0; u < S; ++u) {
      uint32_t off = 0;
      for (uint32_t i = 0; i < mo.h.k; ++i)
        if ((u >> i) & 1) off |= 1u << lo.ltgt[i];
      mo.h.off[u] = off;
    }
    mo.data.resize((size_t)S * S * 2 * sizeof(R));
    R *d = reinterpret_cast<R *>(mo.data.data());
    for (uint32_t i = 0; i < S * S; ++i) {
      d[2 * i] = (R)f.m_sorted[i].real();
      d[2 * i + 1] = (R)f.m_sorted[i].imag();
    }
  }
  mo.h.data_bytes = (uint32_t)mo.data.size();
  pass->ops.push_back(mo);
}

template <typename R>
void push_exch(uint32_t a, uint32_t b, uint32_t lctrl, uint64_t gmask, HostPass *pass, uint32_t T) {
  HostMicroOp mo;
  memset(&mo.h, 0, sizeof(mo.h));
  mo.h.kind = MK_EXCH;
  std::vector<uint32_t> ins = {a, b};
  for (uint32_t c = 0; c < 32; ++c)
    if ((lctrl >> c) & 1) ins.push_back(c);
  std::sort(ins.begin(), ins.end());
  mo.h.ins_n = (uint32_t)ins.size();
  for (size_t i = 0; i < ins.size(); ++i) mo.h.ins_pos[i] = ins[i];
  mo.h.lor_mask = lctrl;
  mo.h.gmask = gmask;
  mo.h.groups_log2 = T - mo.h.ins_n;
  mo.h.off[0] = 1u << a;
  mo.h.off[1] = 1u << b;
  pass->ops.push_back(mo);
}

template <typename R>
struct DiagAccum {  // consecutive diagonal gates merge into one DIAG micro-op
  std::vector<DiagTerm<R>> terms;
  void flush(HostPass *pass) {
    if (terms.empty()) return;
    HostMicroOp mo;
    memset(&mo.h, 0, sizeof(mo.h));
    mo.h.kind = MK_DIAG;
    mo.h.nterms = (uint32_t)terms.size();
    mo.data.resize(terms.size() * sizeof(DiagTerm<R>));
    memcpy(mo.data.data(), terms.data(), mo.data.size());
    mo.h.data_bytes = (uint32_t)mo.data.size();
    pass->ops.push_back(mo);
    terms.clear();
  }
};

// Translate the taken ops of one pass into micro-ops, fusing runs of small all-local ops
// into dense blocks of <= cfg.max_block_k bits.
template <typename R>
void emit_pass(const std::vector<FlatOp> &ops, const std::vector<size_t> &taken, const PassHeader &hdr,
               const PlanConfig &cfg, HostPass *pass) {
  const uint32_t T = hdr.T, L = hdr.L;
  // physical bit -> tile-local bit (or -1)
  int local_of[64];
  for (int b = 0; b < 64; ++b) local_of[b] = -1;
  for (uint32_t b = 0; b < L; ++b) local_of[b] = (int)b;
  for (uint32_t i = 0; i < hdr.m; ++i) local_of[hdr.hi_pos[i]] = (int)(L + i);

  std::vector<Block> open;  // open blocks with pairwise-disjoint bit sets
  DiagAccum<R> diag;

  auto flush_blocks_touching = [&](uint32_t lmask, bool all) {
    for (size_t i = 0; i < open.size();) {
      uint32_t bm = 0;
      for (size_t j = 0; j < open[i].bits.size(); ++j) bm |= 1u << open[i].bits[j];
      if (all || (bm & lmask)) {
        push_dense<R>(open[i], pass, T);
        open.erase(open.begin() + i);
      } else {
        ++i;
      }
    }
  };

  // Merge a new block with the open blocks it touches if the union stays within
  // max_block_k bits; otherwise emit those and open the new one.
  auto merge_block = [&](const Block &nb) {
    uint32_t nbm = 0;
    for (size_t j = 0; j < nb.bits.size(); ++j) nbm |= 1u << nb.bits[j];
    std::vector<size_t> hit;
    uint32_t um = nbm;
    for (size_t i = 0; i < open.size(); ++i) {
      uint32_t bm = 0;
      for (size_t j = 0; j < open[i].bits.size(); ++j) bm |= 1u << open[i].bits[j];
      if (bm & nbm) {
        hit.push_back(i);
        um |= bm;
      }
    }
    if ((uint32_t)popc(um) > cfg.max_block_k) {
      flush_blocks_touching(nbm, false);
      open.push_back(nb);
      return;
    }
    std::vector<uint32_t> ubits;
    for (uint32_t b = 0; b < 32; ++b)
      if ((um >> b) & 1) ubits.push_back(b);
    const size_t S = (size_t)1 << ubits.size();
    std::vector<cplx> acc(S * S, cplx(0, 0));
    for (size_t r = 0; r < S; ++r) acc[r * S + r] = cplx(1, 0);
    for (size_t h = 0; h < hit.size(); ++h)  // disjoint blocks commute: any order
      acc = matmul(embed(open[hit[h]].m, open[hit[h]].bits, ubits), acc, S);
    acc = matmul(embed(nb.m, nb.bits, ubits), acc, S);  // the new op acts last
    for (size_t h = hit.size(); h-- > 0;) open.erase(open.begin() + hit[h]);
    Block merged;
    merged.bits = ubits;
    merged.m = acc;
    open.push_back(merged);
  };

  for (size_t ti = 0; ti < taken.size(); ++ti) {
    const FlatOp &f = ops[taken[ti]];
    LocalOp lo;
    lo.f = &f;
    for (uint32_t b = 0; b < 64; ++b)
      if ((f.ctrl_mask >> b) & 1) {
        if (local_of[b] >= 0)
          lo.lctrl |= 1u << local_of[b];
        else
          lo.gmask |= 1ull << b;
      }
    if (f.cls == CLASS_DIAGONAL) {
      uint32_t touch = lo.lctrl;
      bool all_local = lo.gmask == 0;
      for (size_t i = 0; i < f.diag_bits.size(); ++i) {
        if (local_of[f.diag_bits[i]] >= 0)
          touch |= 1u << local_of[f.diag_bits[i]];
        else
          all_local = false;
      }
      // A small all-local diagonal op that touches an open block is folded into it for free.
      if (cfg.fuse_blocks && all_local && (uint32_t)popc(touch) <= cfg.max_block_k) {
        uint32_t um = touch;
        bool hits = false;
        for (size_t i = 0; i < open.size(); ++i) {
          uint32_t bm = 0;
          for (size_t j = 0; j < open[i].bits.size(); ++j) bm |= 1u << open[i].bits[j];
          if (bm & touch) {
            hits = true;
            um |= bm;
          }
        }
        if (hits && (uint32_t)popc(um) <= cfg.max_block_k && diag.terms.empty()) {
          Block db;
          for (uint32_t b = 0; b < 32; ++b)
            if ((touch >> b) & 1) db.bits.push_back(b);
          const size_t S = (size_t)1 << db.bits.size();
          db.m.assign(S * S, cplx(0, 0));
          for (size_t sidx = 0; sidx < S; ++sidx) {
            bool ctrl_ok = true;
            size_t u = 0;
            for (size_t j = 0; j < db.bits.size(); ++j) {
              const uint32_t lb = db.bits[j];
              const size_t v = (sidx >> j) & 1;
              if (((lo.lctrl >> lb) & 1) && !v) ctrl_ok = false;
              for (size_t i = 0; i < f.diag_bits.size(); ++i)
                if ((uint32_t)local_of[f.diag_bits[i]] == lb) u |= v << i;
            }
            db.m[sidx * S + sidx] = ctrl_ok ? f.diag[u] : cplx(1, 0);
          }
          merge_block(db);
          continue;
        }
      }
      // masked-phase terms; diagonal ops commute with each other and with every block bit they
      // touch only diagonally -- but not with open blocks acting NON-diagonally on shared bits.
      flush_blocks_touching(touch, false);
      const size_t nd = f.diag_bits.size();
      for (size_t u = 0; u < f.diag.size(); ++u) {
        if (f.diag[u].real() == 1.0 && f.diag[u].imag() == 0.0) continue;
        DiagTerm<R> t;
        t.gmask = lo.gmask;
        t.gval = lo.gmask;
        t.lmask = lo.lctrl;
        t.lval = lo.lctrl;
        for (size_t i = 0; i < nd; ++i) {
          const uint32_t b = f.diag_bits[i];
          const uint64_t v = (u >> i) & 1;
          if (local_of[b] >= 0) {
            t.lmask |= 1u << local_of[b];
            t.lval |= (uint32_t)v << local_of[b];
          } else {
            t.gmask |= 1ull << b;
            t.gval |= v << b;
          }
        }
        t.re = (R)f.diag[u].real();
        t.im = (R)f.diag[u].imag();
        if (diag.terms.size() >= kMaxDiagTerms) diag.flush(pass);
        diag.terms.push_back(t);
      }
      continue;
    }
    // non-diagonal op: pending diagonal terms must be emitted first if they share a bit
    // (simple and safe: always flush them)
    diag.flush(pass);
    if (f.cls == CLASS_BITSWAP) {
      for (size_t i = 0; i < f.swaps.size(); ++i)
        lo.lswaps.push_back(std::make_pair((uint32_t)local_of[f.swaps[i].first], (uint32_t)local_of[f.swaps[i].second]));
    } else {
      for (size_t i = 0; i < f.tgt_sorted.size(); ++i) lo.ltgt.push_back((uint32_t)local_of[f.tgt_sorted[i]]);
      // local order equals physical order (low bits identity, high bits ascending), so m_sorted stays valid
    }
    Block nb;
    const bool blockable = cfg.fuse_blocks && as_block(lo, cfg.max_block_k, &nb);
    if (!blockable) {
      uint32_t touch = lo.lctrl;
      for (size_t i = 0; i < lo.ltgt.size(); ++i) touch |= 1u << lo.ltgt[i];
      for (size_t i = 0; i < lo.lswaps.size(); ++i) touch |= (1u << lo.lswaps[i].first) | (1u << lo.lswaps[i].second);
      flush_blocks_touching(touch, false);
      if (f.cls == CLASS_BITSWAP) {
        for (size_t i = 0; i < lo.lswaps.size(); ++i)
          push_exch<R>(std::min(lo.lswaps[i].first, lo.lswaps[i].second),
                       std::max(lo.lswaps[i].first, lo.lswaps[i].second), lo.lctrl, lo.gmask, pass, T);
      } else {
        push_controlled_dense<R>(lo, pass, T);
      }
      continue;
    }
    merge_block(nb);
  }
  diag.flush(pass);
  flush_blocks_touching(0, true);
}

}  // namespace

PlanConfig default_plan_config(qip_prec prec, uint32_t n_local) {
  PlanConfig c;
  c.T = prec == QIP_F32 ? 13 : 12;  // 64 KiB of amplitudes per tile
  c.L = prec == QIP_F32 ? 6 : 5;    // 512-byte contiguous runs in HBM
  // test / tuning knobs (tile geometry only; results are unaffected)
  if (const char *e = getenv("QIPB200_TILE_T")) c.T = std::min<uint32_t>((uint32_t)atoi(e), c.T);
  if (const char *e = getenv("QIPB200_TILE_L")) c.L = (uint32_t)atoi(e);
  if (const char *e = getenv("QIPB200_BLOCK_K")) c.max_block_k = std::max(1, std::min(3, atoi(e)));
  if (const char *e = getenv("QIPB200_NO_BLOCK_FUSION")) c.fuse_blocks = atoi(e) == 0;
  if (n_local < c.T) c.T = n_local;
  if (c.L > c.T) c.L = c.T;
  if (c.T - c.L > kTileMaxHigh) c.L = c.T - kTileMaxHigh;
  if (prec == QIP_F32 && c.L == 0 && c.T >= 1) c.L = 1;  // a 16-byte unit holds two f32 amplitudes
  return c;
}

void plan_passes(const std::vector<FlatOp> &ops, uint32_t n_local, qip_prec prec, const PlanConfig &cfg_in,
                 std::vector<PlanStep> *steps) {
  PlanConfig cfg = cfg_in;
  if (cfg.T > n_local) cfg.T = n_local;
  if (cfg.L > cfg.T) cfg.L = cfg.T;
  const uint32_t m = cfg.T - cfg.L;
  const uint64_t low_mask = (1ull << cfg.L) - 1ull;
  std::vector<OpInfo> info(ops.size());
  std::vector<size_t> remaining;
  for (size_t i = 0; i < ops.size(); ++i) {
    if (ops[i].cls == CLASS_IDENTITY) continue;
    info[i] = analyse(ops[i], cfg);
    remaining.push_back(i);
  }
  while (!remaining.empty()) {
    uint64_t S_high = 0, pend_d = 0, pend_nd = 0;
    std::vector<size_t> taken, left;
    double unfused = 0.0;
    for (size_t r = 0; r < remaining.size(); ++r) {
      const size_t idx = remaining[r];
      const OpInfo &o = info[idx];
      const bool conflict = (o.nd & (pend_d | pend_nd)) || (o.dg & pend_nd);
      if (!conflict && o.tile_ok) {
        const uint64_t need = o.need_tile & ~low_mask & ~S_high;
        if ((uint32_t)popc(S_high | need) <= m) {
          S_high |= need;
          taken.push_back(idx);
          unfused += o.unfused_cost;
          continue;
        }
      }
      pend_d |= o.dg;
      pend_nd |= o.nd;
      left.push_back(idx);
    }
    if (taken.empty() || unfused <= 1.05) {
      // not worth a full sweep: run the first remaining op with its per-gate kernel
      PlanStep st;
      st.is_pass = false;
      st.op_index = remaining[0];
      steps->push_back(st);
      remaining.erase(remaining.begin());
      continue;
    }
    PlanStep st;
    st.is_pass = true;
    PassHeader &h = st.pass.hdr;
    memset(&h, 0, sizeof(h));
    h.T = cfg.T;
    h.L = cfg.L;
    h.m = m;
    // pad the tile-bit set to exactly m bits with the highest unused bits
    for (int b = (int)n_local - 1; b >= (int)cfg.L && (uint32_t)popc(S_high) < m; --b)
      if (!((S_high >> b) & 1)) S_high |= 1ull << b;
    uint32_t c = 0;
    for (uint32_t b = 0; b < 64; ++b)
      if ((S_high >> b) & 1) h.hi_pos[c++] = b;
    for (uint32_t ch = 0; ch < (1u << m); ++ch) {
      uint64_t off = 0;
      for (uint32_t i = 0; i < m; ++i)
        if ((ch >> i) & 1) off |= 1ull << h.hi_pos[i];
      h.chunk_off[ch] = off;
    }
    if (prec == QIP_F32)
      emit_pass<float>(ops, taken, h, cfg, &st.pass);
    else
      emit_pass<double>(ops, taken, h, cfg, &st.pass);
    h.n_ops = (uint32_t)st.pass.ops.size();
    st.pass.n_gates = (uint32_t)taken.size();
    steps->push_back(st);
    remaining.swap(left);
  }
}

void serialise_pass(const HostPass &p, std::vector<unsigned char> *blob) {
  PassHeader h = p.hdr;
  size_t bytes = 0;
  for (size_t i = 0; i < p.ops.size(); ++i) bytes += sizeof(MicroOp) + ((p.ops[i].data.size() + 15) & ~(size_t)15);
  h.blob_bytes = (uint32_t)bytes;
  blob->resize(sizeof(PassHeader) + bytes);
  memcpy(blob->data(), &h, sizeof(h));
  unsigned char *w = blob->data() + sizeof(PassHeader);
  for (size_t i = 0; i < p.ops.size(); ++i) {
    MicroOp mh = p.ops[i].h;
    mh.data_bytes = (uint32_t)((p.ops[i].data.size() + 15) & ~(size_t)15);
    memcpy(w, &mh, sizeof(mh));
    w += sizeof(mh);
    if (!p.ops[i].data.empty()) memcpy(w, p.ops[i].data.data(), p.ops[i].data.size());
    w += mh.data_bytes;
  }
}

}  // namespace qipb200
