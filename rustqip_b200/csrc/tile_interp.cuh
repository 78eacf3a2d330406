// tile_interp.cuh -- the register-resident interpreter of a MK_SUPER micro-op, written as
// inline PTX on NAMED registers.
//
// Why PTX: the natural C++ (arrays re[8], im[8] updated inside `for (elem) switch (opcode)`)
// compiles to a loop whose 32..64 live values are copied register-to-register on every
// iteration (phi copies at the switch merge): measured 33-38 % of all issued instructions.
// Here the amplitudes of a group live in PTX registers declared once per kernel
// (q?r0..7 / q?i0..7, prefix a/b for the two groups a thread owns); every elementary op
// is an in-place update of those names, so there is nothing to copy at any merge point.
//
// Macro parameters:  T = "f64" | "f32";  P = "a" | "b" (group);  I0, I1 = sub-indices.
#pragma once

#define QIP_DECL_GROUP(T, P)                                                                        \
  asm volatile(".reg ." T " q" P "r0, q" P "r1, q" P "r2, q" P "r3, q" P "r4, q" P "r5, q" P "r6, q" P \
               "r7, q" P "i0, q" P "i1, q" P "i2, q" P "i3, q" P "i4, q" P "i5, q" P "i6, q" P "i7;")

#define QIP_LD(T, P, I, ADDR) \
  asm volatile("ld.shared.v2." T " {q" P "r" #I ", q" P "i" #I "}, [%0];" ::"r"(ADDR) : "memory")
#define QIP_ST(T, P, I, ADDR) \
  asm volatile("st.shared.v2." T " [%0], {q" P "r" #I ", q" P "i" #I "};" ::"r"(ADDR) : "memory")

// real 2x2 on the pair (I0, I1):  x' = m00 x + m01 y ; y' = m10 x + m11 y   (8 flops per component pair)
#define QIP_D1R(T, C, P, I0, I1, M00, M01, M10, M11)                                   \
  asm volatile("{\n\t.reg ." T " t1, t2, t3, t4;\n\t"                                  \
               "mul." T " t1, %1, q" P "r" #I1 ";\n\t"                                 \
               "mul." T " t2, %3, q" P "r" #I1 ";\n\t"                                 \
               "mul." T " t3, %1, q" P "i" #I1 ";\n\t"                                 \
               "mul." T " t4, %3, q" P "i" #I1 ";\n\t"                                 \
               "fma.rn." T " q" P "r" #I1 ", %2, q" P "r" #I0 ", t2;\n\t"              \
               "fma.rn." T " q" P "i" #I1 ", %2, q" P "i" #I0 ", t4;\n\t"              \
               "fma.rn." T " q" P "r" #I0 ", %0, q" P "r" #I0 ", t1;\n\t"              \
               "fma.rn." T " q" P "i" #I0 ", %0, q" P "i" #I0 ", t3;\n\t}" ::C(M00),   \
               C(M01), C(M10), C(M11))

// un-normalised Hadamard butterfly on the pair (I0, I1):  x' = x + y ; y' = x - y
#define QIP_HAD(T, P, I0, I1)                                                     \
  asm volatile("{\n\t.reg ." T " t1, t2;\n\t"                                    \
               "add.rn." T " t1, q" P "r" #I0 ", q" P "r" #I1 ";\n\t"             \
               "add.rn." T " t2, q" P "i" #I0 ", q" P "i" #I1 ";\n\t"             \
               "sub.rn." T " q" P "r" #I1 ", q" P "r" #I0 ", q" P "r" #I1 ";\n\t" \
               "sub.rn." T " q" P "i" #I1 ", q" P "i" #I0 ", q" P "i" #I1 ";\n\t" \
               "mov." T " q" P "r" #I0 ", t1;\n\t"                                \
               "mov." T " q" P "i" #I0 ", t2;\n\t}" ::)

// complex 2x2 on the pair (I0, I1); operands %0..%7 = m00r m00i m01r m01i m10r m10i m11r m11i,
// %8..%11 = -m00i -m01i -m10i -m11i (negated once per op in C++)
#define QIP_D1C(T, C, P, I0, I1, A, B, Cc, D, E, F, G, H, NB, ND, NF, NH)                \
  asm volatile("{\n\t.reg ." T " a1, a2, b1, b2, tx;\n\t"                               \
               "mul." T " a1, %2, q" P "r" #I1 ";\n\t"           /* m01r*yr          */ \
               "fma.rn." T " a1, %9, q" P "i" #I1 ", a1;\n\t"    /* - m01i*yi        */ \
               "mul." T " a2, %2, q" P "i" #I1 ";\n\t"           /* m01r*yi          */ \
               "fma.rn." T " a2, %3, q" P "r" #I1 ", a2;\n\t"    /* + m01i*yr        */ \
               "mul." T " b1, %6, q" P "r" #I1 ";\n\t"           /* m11r*yr          */ \
               "fma.rn." T " b1, %11, q" P "i" #I1 ", b1;\n\t"   /* - m11i*yi        */ \
               "mul." T " b2, %6, q" P "i" #I1 ";\n\t"           /* m11r*yi          */ \
               "fma.rn." T " b2, %7, q" P "r" #I1 ", b2;\n\t"    /* + m11i*yr        */ \
               "fma.rn." T " q" P "r" #I1 ", %4, q" P "r" #I0 ", b1;\n\t"              /* yr' = m10r*xr + b1 */ \
               "fma.rn." T " q" P "r" #I1 ", %10, q" P "i" #I0 ", q" P "r" #I1 ";\n\t" /*       - m10i*xi    */ \
               "fma.rn." T " q" P "i" #I1 ", %4, q" P "i" #I0 ", b2;\n\t"              /* yi' = m10r*xi + b2 */ \
               "fma.rn." T " q" P "i" #I1 ", %5, q" P "r" #I0 ", q" P "i" #I1 ";\n\t"  /*       + m10i*xr    */ \
               "mov." T " tx, q" P "r" #I0 ";\n\t"                                      \
               "fma.rn." T " q" P "r" #I0 ", %0, q" P "r" #I0 ", a1;\n\t"              /* xr' = m00r*xr + a1 */ \
               "fma.rn." T " q" P "r" #I0 ", %8, q" P "i" #I0 ", q" P "r" #I0 ";\n\t"  /*       - m00i*xi    */ \
               "fma.rn." T " q" P "i" #I0 ", %0, q" P "i" #I0 ", a2;\n\t"              /* xi' = m00r*xi + a2 */ \
               "fma.rn." T " q" P "i" #I0 ", %1, tx, q" P "i" #I0 ";\n\t}"             /*       + m00i*xr(old) */ \
               ::C(A), C(B), C(Cc), C(D), C(E), C(F), C(G), C(H), C(NB), C(ND), C(NF), C(NH))

// exchange amplitudes I0 <-> I1 (X on the pair): register moves only, bit-exact
#define QIP_XCH(T, P, I0, I1)                                               \
  asm volatile("{\n\t.reg ." T " t1, t2;\n\t"                              \
               "mov." T " t1, q" P "r" #I0 ";\n\t"                          \
               "mov." T " t2, q" P "i" #I0 ";\n\t"                          \
               "mov." T " q" P "r" #I0 ", q" P "r" #I1 ";\n\t"              \
               "mov." T " q" P "i" #I0 ", q" P "i" #I1 ";\n\t"              \
               "mov." T " q" P "r" #I1 ", t1;\n\t"                          \
               "mov." T " q" P "i" #I1 ", t2;\n\t}" ::)

// amplitude I *= (wr + i wi)
#define QIP_PH(T, C, P, I, WR, WI)                                                   \
  asm volatile("{\n\t.reg ." T " t1, t2;\n\t"                                        \
               "mul." T " t1, %1, q" P "i" #I ";\n\t"                                \
               "neg." T " t1, t1;\n\t"                                               \
               "mul." T " t2, %1, q" P "r" #I ";\n\t"                                \
               "fma.rn." T " q" P "r" #I ", %0, q" P "r" #I ", t1;\n\t"              \
               "fma.rn." T " q" P "i" #I ", %0, q" P "i" #I ", t2;\n\t}" ::C(WR), C(WI))

// read / write a named register pair from C++ (dense 8x8 path only)
#define QIP_GET(T, CO, P, I, RE, IM) \
  asm volatile("mov." T " %0, q" P "r" #I ";\n\tmov." T " %1, q" P "i" #I ";" : CO(RE), CO(IM))
#define QIP_SET(T, C, P, I, RE, IM) \
  asm volatile("mov." T " q" P "r" #I ", %0;\n\tmov." T " q" P "i" #I ", %1;" ::C(RE), C(IM))

// pairs (p, i0, i1) of sub-indices differing in sub-bit J, p ascending
#define QIP_PAIRS_0(X) X(0, 0, 1) X(1, 2, 3) X(2, 4, 5) X(3, 6, 7)
#define QIP_PAIRS_1(X) X(0, 0, 2) X(1, 1, 3) X(2, 4, 6) X(3, 5, 7)
#define QIP_PAIRS_2(X) X(0, 0, 4) X(1, 1, 5) X(2, 2, 6) X(3, 3, 7)
#define QIP_AMPS(X) X(0) X(1) X(2) X(3) X(4) X(5) X(6) X(7)

#define QIP_CD(x) "d"(x)
#define QIP_CF(x) "f"(x)
#define QIP_COD(x) "=d"(x)
#define QIP_COF(x) "=f"(x)

// Defines  template <int G> void NAME(uint32_t tile_saddr, const MicroOp *mo, const unsigned char *data, uint64_t base,
//                                     const R *tbl, uint64_t condbits)
// for one precision.  tile_saddr = shared-memory byte address of the tile; condbits = this CTA's
// evaluation of the pass's condition table (bit s <-> CondTerm s).
//
// Per group: the group counter is expanded with three `t += t & (~0 << p)` steps (inserting a
// zero bit at position p: low + 2*high = t + high), the 8 addresses are swz(t) ^ soff[u] with the
// swizzled byte offsets precomputed on the host (the XOR swizzle is GF(2)-linear).  The
// elementary ops are dispatched by ONE jump on the host-computed case id; every frequent shape
// (full 2x2, 2x2 under one/two in-group controls, phase on one/two sub-bits) is straight-line
// code without mask tests.
#define QIP_DEFINE_RUN_SUPER(NAME, R, T, C, CO, SWZ, ESHIFT)                                              \
  template <int G>                                                                                              \
  __device__ __forceinline__ void NAME(uint32_t tile_saddr, const MicroOp *mo, const unsigned char *data,      \
                                       uint64_t base, const R *tbl, uint64_t condbits) {                        \
    typedef R QipReal;                                                                                          \
    const uint32_t groups = 1u << mo->groups_log2;                                                              \
    const uint32_t hm0 = ~0u << mo->ins_pos[0], hm1 = ~0u << mo->ins_pos[1], hm2 = ~0u << mo->ins_pos[2];      \
    for (uint32_t g = threadIdx.x; g < groups; g += G * kTileThreads) {                                         \
      const bool two = (G == 2) && (g + kTileThreads < groups); /* warp-uniform */                              \
      uint32_t ta = g, tb = two ? g + kTileThreads : g;                                                         \
      ta += ta & hm0; ta += ta & hm1; ta += ta & hm2;                                                           \
      tb += tb & hm0; tb += tb & hm1; tb += tb & hm2;                                                           \
      const uint32_t sa = SWZ(ta) << ESHIFT, sb = SWZ(tb) << ESHIFT;                                            \
      uint32_t aa[8], ab[8];                                                                                    \
      _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                                           \
        const uint32_t so = mo->soff[u];                                                                        \
        aa[u] = tile_saddr + (sa ^ so);                                                                         \
        ab[u] = tile_saddr + (sb ^ so);                                                                         \
      }                                                                                                         \
      QIP_LD(T, "a", 0, aa[0]); QIP_LD(T, "a", 1, aa[1]); QIP_LD(T, "a", 2, aa[2]); QIP_LD(T, "a", 3, aa[3]);   \
      QIP_LD(T, "a", 4, aa[4]); QIP_LD(T, "a", 5, aa[5]); QIP_LD(T, "a", 6, aa[6]); QIP_LD(T, "a", 7, aa[7]);   \
      if (G == 2) {                                                                                             \
        QIP_LD(T, "b", 0, ab[0]); QIP_LD(T, "b", 1, ab[1]); QIP_LD(T, "b", 2, ab[2]); QIP_LD(T, "b", 3, ab[3]); \
        QIP_LD(T, "b", 4, ab[4]); QIP_LD(T, "b", 5, ab[5]); QIP_LD(T, "b", 6, ab[6]); QIP_LD(T, "b", 7, ab[7]); \
      }                                                                                                         \
      {                                                                                                         \
      const unsigned char *ep = data;                                                                           \
      uint32_t op_next = reinterpret_cast<const Elem<R> *>(ep)->op;                                             \
      for (;;) {                                                                                                \
        const uint32_t op = op_next;                                                                            \
        const uint32_t id = op & kElemCaseMask;                                                                 \
        if (id == EC_END) break;                                                                                \
        const Elem<R> *e = reinterpret_cast<const Elem<R> *>(ep);                                               \
        ep += ((op >> 20) & 0x7ffu) << 4;                                                                       \
        op_next = reinterpret_cast<const Elem<R> *>(ep)->op; /* next descriptor in flight during this op */     \
        asm volatile("" ::"r"(op_next));                     /* (keeps the load above the arithmetic) */        \
        if ((int32_t)op < 0) { /* control outside the tile: evaluated once per CTA */                           \
          const uint32_t slot = (op >> kElemCondShift) & 63u;                                                   \
          if (slot != kCondOverflow) {                                                                          \
            if (!((condbits >> slot) & 1ull)) continue;                                                         \
          } else if ((base & e->gmask) != e->gval) {                                                            \
            continue;                                                                                           \
          }                                                                                                     \
        }                                                                                                       \
        /* dispatch: hot shapes first, compare chains (a flat switch is lowered to a balanced compare    \
           tree plus small jump tables whose target load sits on the critical path) */                        \
        if (id >= EC_HAD) { /* Hadamard as an add/sub butterfly, scale folded into another gate by the planner */\
          if (id == EC_HAD) { _QIP_HAD_U(T, 0, 1) _QIP_HAD_U(T, 2, 3) _QIP_HAD_U(T, 4, 5) _QIP_HAD_U(T, 6, 7) } \
          else if (id == EC_HAD + 1) { _QIP_HAD_U(T, 0, 2) _QIP_HAD_U(T, 1, 3) _QIP_HAD_U(T, 4, 6) _QIP_HAD_U(T, 5, 7) } \
          else { _QIP_HAD_U(T, 0, 4) _QIP_HAD_U(T, 1, 5) _QIP_HAD_U(T, 2, 6) _QIP_HAD_U(T, 3, 7) }              \
        } else if (id < EC_D1C_FULL) {                                                                          \
          _QIP_LOAD_MR                                                                                          \
          if (id == EC_D1R_FULL) { _QIP_D1R_U(T, C, 0, 1) _QIP_D1R_U(T, C, 2, 3) _QIP_D1R_U(T, C, 4, 5) _QIP_D1R_U(T, C, 6, 7) } \
          else if (id == EC_D1R_FULL + 1) { _QIP_D1R_U(T, C, 0, 2) _QIP_D1R_U(T, C, 1, 3) _QIP_D1R_U(T, C, 4, 6) _QIP_D1R_U(T, C, 5, 7) } \
          else { _QIP_D1R_U(T, C, 0, 4) _QIP_D1R_U(T, C, 1, 5) _QIP_D1R_U(T, C, 2, 6) _QIP_D1R_U(T, C, 3, 7) } \
        } else if (id >= EC_PHASE_J) {                                                                          \
          if (id < EC_D1R_C1) { /* every amplitude with sub-bit j set: T, S, Rz, control outside the tile */   \
            _QIP_LOAD_W                                                                                         \
            if (id == EC_PHASE_J) { _QIP_PH_U(T, C, 1) _QIP_PH_U(T, C, 3) _QIP_PH_U(T, C, 5) _QIP_PH_U(T, C, 7) } \
            else if (id == EC_PHASE_J + 1) { _QIP_PH_U(T, C, 2) _QIP_PH_U(T, C, 3) _QIP_PH_U(T, C, 6) _QIP_PH_U(T, C, 7) } \
            else { _QIP_PH_U(T, C, 4) _QIP_PH_U(T, C, 5) _QIP_PH_U(T, C, 6) _QIP_PH_U(T, C, 7) }                \
          } else if (id < EC_D1R_C2) { /* one control inside the group: pairs 1,3 or 2,3 of sub-bit j */        \
            _QIP_LOAD_MR                                                                                        \
            if (id < EC_D1R_C1 + 2) {                                                                           \
              if (id == EC_D1R_C1) { _QIP_D1R_U(T, C, 2, 3) _QIP_D1R_U(T, C, 6, 7) }                            \
              else { _QIP_D1R_U(T, C, 4, 5) _QIP_D1R_U(T, C, 6, 7) }                                            \
            } else if (id < EC_D1R_C1 + 4) {                                                                    \
              if (id == EC_D1R_C1 + 2) { _QIP_D1R_U(T, C, 1, 3) _QIP_D1R_U(T, C, 5, 7) }                        \
              else { _QIP_D1R_U(T, C, 4, 6) _QIP_D1R_U(T, C, 5, 7) }                                            \
            } else {                                                                                            \
              if (id == EC_D1R_C1 + 4) { _QIP_D1R_U(T, C, 1, 5) _QIP_D1R_U(T, C, 3, 7) }                        \
              else { _QIP_D1R_U(T, C, 2, 6) _QIP_D1R_U(T, C, 3, 7) }                                            \
            }                                                                                                   \
          } else if (id < EC_PHASE_2) { /* both other sub-bits are controls: pair 3 only */                     \
            _QIP_LOAD_MR                                                                                        \
            if (id == EC_D1R_C2) { _QIP_D1R_U(T, C, 6, 7) }                                                     \
            else if (id == EC_D1R_C2 + 1) { _QIP_D1R_U(T, C, 5, 7) }                                            \
            else { _QIP_D1R_U(T, C, 3, 7) }                                                                     \
          } else { /* two sub-bits set: CZ, controlled phase inside the group */                                \
            _QIP_LOAD_W                                                                                         \
            if (id == EC_PHASE_2) { _QIP_PH_U(T, C, 3) _QIP_PH_U(T, C, 7) }                                     \
            else if (id == EC_PHASE_2 + 1) { _QIP_PH_U(T, C, 5) _QIP_PH_U(T, C, 7) }                            \
            else { _QIP_PH_U(T, C, 6) _QIP_PH_U(T, C, 7) }                                                      \
          }                                                                                                     \
        } else if (id < EC_D1R_MASK) {                                                                          \
          _QIP_LOAD_MC                                                                                          \
          if (id == EC_D1C_FULL) { _QIP_D1C_U(T, C, 0, 1) _QIP_D1C_U(T, C, 2, 3) _QIP_D1C_U(T, C, 4, 5) _QIP_D1C_U(T, C, 6, 7) } \
          else if (id == EC_D1C_FULL + 1) { _QIP_D1C_U(T, C, 0, 2) _QIP_D1C_U(T, C, 1, 3) _QIP_D1C_U(T, C, 4, 6) _QIP_D1C_U(T, C, 5, 7) } \
          else { _QIP_D1C_U(T, C, 0, 4) _QIP_D1C_U(T, C, 1, 5) _QIP_D1C_U(T, C, 2, 6) _QIP_D1C_U(T, C, 3, 7) } \
        } else {                                                                                                \
          switch (id) {                                                                                         \
            _QIP_CASES_D1R_MASK(T, C)                                                                           \
            _QIP_CASES_D1C_MASK(T, C)                                                                           \
            _QIP_CASES_PHASE_GEN(T, C)                                                                          \
            _QIP_CASES_X(T)                                                                                     \
            case EC_DENSE3: {                                                                                   \
              const R *m8 = reinterpret_cast<const R *>(e + 1);                                                 \
              _QIP_D3_BODY(T, C, CO, R, "a")                                                                    \
              if (G == 2) { _QIP_D3_BODY(T, C, CO, R, "b") }                                                    \
            } break;                                                                                            \
            default: break;                                                                                     \
          }                                                                                                     \
        }                                                                                                       \
      }                                                                                                         \
      }                                                                                                         \
      QIP_ST(T, "a", 0, aa[0]); QIP_ST(T, "a", 1, aa[1]); QIP_ST(T, "a", 2, aa[2]); QIP_ST(T, "a", 3, aa[3]);   \
      QIP_ST(T, "a", 4, aa[4]); QIP_ST(T, "a", 5, aa[5]); QIP_ST(T, "a", 6, aa[6]); QIP_ST(T, "a", 7, aa[7]);   \
      if (two) {                                                                                                \
        QIP_ST(T, "b", 0, ab[0]); QIP_ST(T, "b", 1, ab[1]); QIP_ST(T, "b", 2, ab[2]); QIP_ST(T, "b", 3, ab[3]); \
        QIP_ST(T, "b", 4, ab[4]); QIP_ST(T, "b", 5, ab[5]); QIP_ST(T, "b", 6, ab[6]); QIP_ST(T, "b", 7, ab[7]); \
      }                                                                                                         \
    }                                                                                                           \
  }

// --- bodies (use the local names of QIP_DEFINE_RUN_SUPER) ---
// Pair p of sub-bit j = the p-th (ascending) sub-index with bit j clear, and its partner:
//   j=0: (0,1) (2,3) (4,5) (6,7)   j=1: (0,2) (1,3) (4,6) (5,7)   j=2: (0,4) (1,5) (2,6) (3,7)
// bit 0 of p <-> the lower of the two other sub-bits, bit 1 of p <-> the higher one.
#define _QIP_HAD_U(T, i0, i1)     \
  QIP_HAD(T, "a", i0, i1);        \
  if (G == 2) QIP_HAD(T, "b", i0, i1);
#define _QIP_D1R_U(T, C, i0, i1)                                        \
  QIP_D1R(T, C, "a", i0, i1, m00, m01, m10, m11);                       \
  if (G == 2) QIP_D1R(T, C, "b", i0, i1, m00, m01, m10, m11);
#define _QIP_D1R_STEP(T, C, p, i0, i1) \
  if ((pm >> p) & 1u) { _QIP_D1R_U(T, C, i0, i1) }
#define _QIP_LOAD_MR const QipReal m00 = e->m[0], m01 = e->m[1], m10 = e->m[2], m11 = e->m[3];
#define _QIP_LOAD_PM const uint32_t pm = (op >> 12) & 0xffu;
#define _QIP_CASES_D1R_MASK(T, C)                                                                                     \
  case EC_D1R_MASK + 0: { _QIP_LOAD_MR _QIP_LOAD_PM                                                                   \
    _QIP_D1R_STEP(T, C, 0, 0, 1) _QIP_D1R_STEP(T, C, 1, 2, 3) _QIP_D1R_STEP(T, C, 2, 4, 5) _QIP_D1R_STEP(T, C, 3, 6, 7) } break; \
  case EC_D1R_MASK + 1: { _QIP_LOAD_MR _QIP_LOAD_PM                                                                   \
    _QIP_D1R_STEP(T, C, 0, 0, 2) _QIP_D1R_STEP(T, C, 1, 1, 3) _QIP_D1R_STEP(T, C, 2, 4, 6) _QIP_D1R_STEP(T, C, 3, 5, 7) } break; \
  case EC_D1R_MASK + 2: { _QIP_LOAD_MR _QIP_LOAD_PM                                                                   \
    _QIP_D1R_STEP(T, C, 0, 0, 4) _QIP_D1R_STEP(T, C, 1, 1, 5) _QIP_D1R_STEP(T, C, 2, 2, 6) _QIP_D1R_STEP(T, C, 3, 3, 7) } break;

#define _QIP_D1C_U(T, C, i0, i1)                                                                \
  QIP_D1C(T, C, "a", i0, i1, m0, m1, m2, m3, m4, m5, m6, m7, n1, n3, n5, n7);                   \
  if (G == 2) QIP_D1C(T, C, "b", i0, i1, m0, m1, m2, m3, m4, m5, m6, m7, n1, n3, n5, n7);
#define _QIP_D1C_STEP(T, C, p, i0, i1) \
  if ((pm >> p) & 1u) { _QIP_D1C_U(T, C, i0, i1) }
#define _QIP_LOAD_MC                                                \
  const QipReal m0 = e->m[0], m1 = e->m[1], m2 = e->m[2], m3 = e->m[3];   \
  const QipReal m4 = e->m[4], m5 = e->m[5], m6 = e->m[6], m7 = e->m[7];   \
  const QipReal n1 = -m1, n3 = -m3, n5 = -m5, n7 = -m7;
#define _QIP_CASES_D1C_MASK(T, C)                                                                                     \
  case EC_D1C_MASK + 0: { _QIP_LOAD_MC _QIP_LOAD_PM                                                                   \
    _QIP_D1C_STEP(T, C, 0, 0, 1) _QIP_D1C_STEP(T, C, 1, 2, 3) _QIP_D1C_STEP(T, C, 2, 4, 5) _QIP_D1C_STEP(T, C, 3, 6, 7) } break; \
  case EC_D1C_MASK + 1: { _QIP_LOAD_MC _QIP_LOAD_PM                                                                   \
    _QIP_D1C_STEP(T, C, 0, 0, 2) _QIP_D1C_STEP(T, C, 1, 1, 3) _QIP_D1C_STEP(T, C, 2, 4, 6) _QIP_D1C_STEP(T, C, 3, 5, 7) } break; \
  case EC_D1C_MASK + 2: { _QIP_LOAD_MC _QIP_LOAD_PM                                                                   \
    _QIP_D1C_STEP(T, C, 0, 0, 4) _QIP_D1C_STEP(T, C, 1, 1, 5) _QIP_D1C_STEP(T, C, 2, 2, 6) _QIP_D1C_STEP(T, C, 3, 3, 7) } break;

#define _QIP_X_STEP(T, p, i0, i1)                 \
  if ((xm >> p) & 1u) {                           \
    QIP_XCH(T, "a", i0, i1);                      \
    if (G == 2) QIP_XCH(T, "b", i0, i1);          \
  }
#define _QIP_X_BODY(T)                                                  \
  if (j == 0) {                                                         \
    _QIP_X_STEP(T, 0, 0, 1) _QIP_X_STEP(T, 1, 2, 3) _QIP_X_STEP(T, 2, 4, 5) _QIP_X_STEP(T, 3, 6, 7) \
  } else if (j == 1) {                                                  \
    _QIP_X_STEP(T, 0, 0, 2) _QIP_X_STEP(T, 1, 1, 3) _QIP_X_STEP(T, 2, 4, 6) _QIP_X_STEP(T, 3, 5, 7) \
  } else {                                                              \
    _QIP_X_STEP(T, 0, 0, 4) _QIP_X_STEP(T, 1, 1, 5) _QIP_X_STEP(T, 2, 2, 6) _QIP_X_STEP(T, 3, 3, 7) \
  }
#define _QIP_CASES_X(T)                                                                            \
  case EC_X_FULL + 0: case EC_X_FULL + 1: case EC_X_FULL + 2:                                      \
  case EC_X_MASK + 0: case EC_X_MASK + 1: case EC_X_MASK + 2: {                                    \
    const uint32_t xm = id >= EC_X_MASK ? (op >> 12) & 0xffu : 0xfu;                               \
    const uint32_t j = id >= EC_X_MASK ? id - EC_X_MASK : id - EC_X_FULL;                          \
    _QIP_X_BODY(T)                                                                                 \
  } break;

#define _QIP_PH_U(T, C, c)                          \
  QIP_PH(T, C, "a", c, wr, wi);                     \
  if (G == 2) QIP_PH(T, C, "b", c, wr, wi);
#define _QIP_PH_STEP(T, C, c) \
  if ((pm >> c) & 1u) { _QIP_PH_U(T, C, c) }
#define _QIP_PH_BODY(T, C)                                                                              \
  _QIP_PH_STEP(T, C, 0) _QIP_PH_STEP(T, C, 1) _QIP_PH_STEP(T, C, 2) _QIP_PH_STEP(T, C, 3)               \
  _QIP_PH_STEP(T, C, 4) _QIP_PH_STEP(T, C, 5) _QIP_PH_STEP(T, C, 6) _QIP_PH_STEP(T, C, 7)
#define _QIP_LOAD_W const QipReal wr = e->m[0], wi = e->m[1];
#define _QIP_CASES_PHASE_GEN(T, C)                                                                                    \
  case EC_PHASE: { _QIP_LOAD_W _QIP_LOAD_PM _QIP_PH_BODY(T, C) } break;                                               \
  case EC_PHASEN: { /* run of controlled phases, controls outside the tile: the product was formed */                 \
    const QipReal wr = tbl[2 * e->pad], wi = tbl[2 * e->pad + 1]; /* once per CTA (factor table) */                         \
    _QIP_LOAD_PM _QIP_PH_BODY(T, C) } break;

// dense 8x8 (composed blocks / user 2-3 qubit matrices): through C++ temporaries
#define _QIP_D3_BODY(T, C, CO, R, P)                                                       \
  {                                                                                        \
    R xr[8], xi[8];                                                                        \
    QIP_GET(T, CO, P, 0, xr[0], xi[0]); QIP_GET(T, CO, P, 1, xr[1], xi[1]);                \
    QIP_GET(T, CO, P, 2, xr[2], xi[2]); QIP_GET(T, CO, P, 3, xr[3], xi[3]);                \
    QIP_GET(T, CO, P, 4, xr[4], xi[4]); QIP_GET(T, CO, P, 5, xr[5], xi[5]);                \
    QIP_GET(T, CO, P, 6, xr[6], xi[6]); QIP_GET(T, CO, P, 7, xr[7], xi[7]);                \
    R yr[8], yi[8];                                                                        \
    _Pragma("unroll") for (int u = 0; u < 8; ++u) {                                        \
      R re = (R)0, im = (R)0;                                                              \
      _Pragma("unroll") for (int v = 0; v < 8; ++v) {                                      \
        const R mr = m8[2 * (u * 8 + v)], mi = m8[2 * (u * 8 + v) + 1];                    \
        re = fma(mr, xr[v], re);                                                           \
        re = fma(-mi, xi[v], re);                                                          \
        im = fma(mr, xi[v], im);                                                           \
        im = fma(mi, xr[v], im);                                                           \
      }                                                                                    \
      yr[u] = re;                                                                          \
      yi[u] = im;                                                                          \
    }                                                                                      \
    QIP_SET(T, C, P, 0, yr[0], yi[0]); QIP_SET(T, C, P, 1, yr[1], yi[1]);                  \
    QIP_SET(T, C, P, 2, yr[2], yi[2]); QIP_SET(T, C, P, 3, yr[3], yi[3]);                  \
    QIP_SET(T, C, P, 4, yr[4], yi[4]); QIP_SET(T, C, P, 5, yr[5], yi[5]);                  \
    QIP_SET(T, C, P, 6, yr[6], yi[6]); QIP_SET(T, C, P, 7, yr[7], yi[7]);                  \
  }
