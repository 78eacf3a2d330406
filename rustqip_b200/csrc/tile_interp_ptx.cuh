// tile_interp_ptx.cuh -- EXPERIMENT (QIPB200_TILE_VARIANT bit 3, not measured yet): the elementary-op loop of a
// super-op as ONE block of PTX.
//
// Why: in the C++ interpreter (tile_interp.cuh) the compiler decides where the descriptor word is loaded and how
// the case id is tested; the r1o capture shows a third of the elementary-op time in that dispatch (the word's LDC
// waited for at its use, a chain of ISETP -> BRA pairs).  Here the order is written down: the NEXT op word is
// fetched right after the current one is decoded, the case is reached by one `brx.idx`, the amplitudes stay in the
// kernel-scope named registers (qar0..7 / qai0..7) exactly as in the C++ version, and the per-pair arithmetic is
// the same instruction sequence (so results are bit-identical to the default interpreter).
//
// Covered case ids (the planner marks a super-op MicroOp::pad0 = 1 when every record is one of them; other
// super-ops run through the C++ interpreter): real / complex 2x2 on every pair, real 2x2 under one / two in-group
// controls, phases (one sub-bit, two sub-bits, generic mask, PHASEN), un-normalised Hadamard.
//
// Record layout read here (tile.cuh, Elem<R>): op u32 @0, pad u32 @4, gmask u64 @8, gval u64 @16, m[] @32.
#pragma once

// ---- per-pair / per-amplitude arithmetic on named registers (same sequences as QIP_D1R / QIP_D1C / QIP_PH / QIP_HAD)
#define ZQ_D1R(T, I0, I1)                                   \
  "mul." T " zt1, zm1, qar" #I1 ";\n\t"                     \
  "mul." T " zt2, zm3, qar" #I1 ";\n\t"                     \
  "mul." T " zt3, zm1, qai" #I1 ";\n\t"                     \
  "mul." T " zt4, zm3, qai" #I1 ";\n\t"                     \
  "fma.rn." T " qar" #I1 ", zm2, qar" #I0 ", zt2;\n\t"      \
  "fma.rn." T " qai" #I1 ", zm2, qai" #I0 ", zt4;\n\t"      \
  "fma.rn." T " qar" #I0 ", zm0, qar" #I0 ", zt1;\n\t"      \
  "fma.rn." T " qai" #I0 ", zm0, qai" #I0 ", zt3;\n\t"

#define ZQ_D1C(T, I0, I1)                                          \
  "mul." T " za1, zm2, qar" #I1 ";\n\t"                            \
  "fma.rn." T " za1, zn3, qai" #I1 ", za1;\n\t"                    \
  "mul." T " za2, zm2, qai" #I1 ";\n\t"                            \
  "fma.rn." T " za2, zm3, qar" #I1 ", za2;\n\t"                    \
  "mul." T " zb1, zm6, qar" #I1 ";\n\t"                            \
  "fma.rn." T " zb1, zn7, qai" #I1 ", zb1;\n\t"                    \
  "mul." T " zb2, zm6, qai" #I1 ";\n\t"                            \
  "fma.rn." T " zb2, zm7, qar" #I1 ", zb2;\n\t"                    \
  "fma.rn." T " qar" #I1 ", zm4, qar" #I0 ", zb1;\n\t"             \
  "fma.rn." T " qar" #I1 ", zn5, qai" #I0 ", qar" #I1 ";\n\t"      \
  "fma.rn." T " qai" #I1 ", zm4, qai" #I0 ", zb2;\n\t"             \
  "fma.rn." T " qai" #I1 ", zm5, qar" #I0 ", qai" #I1 ";\n\t"      \
  "mov." T " ztx, qar" #I0 ";\n\t"                                 \
  "fma.rn." T " qar" #I0 ", zm0, qar" #I0 ", za1;\n\t"             \
  "fma.rn." T " qar" #I0 ", zn1, qai" #I0 ", qar" #I0 ";\n\t"      \
  "fma.rn." T " qai" #I0 ", zm0, qai" #I0 ", za2;\n\t"             \
  "fma.rn." T " qai" #I0 ", zm1, ztx, qai" #I0 ";\n\t"

#define ZQ_HAD(T, I0, I1)                                          \
  "add.rn." T " zt1, qar" #I0 ", qar" #I1 ";\n\t"                  \
  "add.rn." T " zt2, qai" #I0 ", qai" #I1 ";\n\t"                  \
  "sub.rn." T " qar" #I1 ", qar" #I0 ", qar" #I1 ";\n\t"           \
  "sub.rn." T " qai" #I1 ", qai" #I0 ", qai" #I1 ";\n\t"           \
  "mov." T " qar" #I0 ", zt1;\n\t"                                 \
  "mov." T " qai" #I0 ", zt2;\n\t"

// amplitude I *= (zwr + i zwi); PRED = "" or "@zqI " (generic masks)
#define ZQ_PH(T, PRED, I)                                          \
  PRED "mul." T " zt1, zwi, qai" #I ";\n\t"                        \
  PRED "neg." T " zt1, zt1;\n\t"                                   \
  PRED "mul." T " zt2, zwi, qar" #I ";\n\t"                        \
  PRED "fma.rn." T " qar" #I ", zwr, qar" #I ", zt1;\n\t"          \
  PRED "fma.rn." T " qai" #I ", zwr, qai" #I ", zt2;\n\t"
#define ZQ_PHM(T, I, BIT)                    \
  "and.b32 zt, zpm, " #BIT ";\n\t"           \
  "setp.ne.u32 zq" #I ", zt, 0;\n\t"         \
  ZQ_PH(T, "@zq" #I " ", I)

// matrix / phase operands of the current record (zcur = its param-space address)
#define ZQ_LOAD_MR(T, O0, O1, O2, O3)            \
  "ld.param." T " zm0, [zcur+" O0 "];\n\t"       \
  "ld.param." T " zm1, [zcur+" O1 "];\n\t"       \
  "ld.param." T " zm2, [zcur+" O2 "];\n\t"       \
  "ld.param." T " zm3, [zcur+" O3 "];\n\t"
#define ZQ_LOAD_MC(T, O0, O1, O2, O3, O4, O5, O6, O7) \
  ZQ_LOAD_MR(T, O0, O1, O2, O3)                       \
  "ld.param." T " zm4, [zcur+" O4 "];\n\t"            \
  "ld.param." T " zm5, [zcur+" O5 "];\n\t"            \
  "ld.param." T " zm6, [zcur+" O6 "];\n\t"            \
  "ld.param." T " zm7, [zcur+" O7 "];\n\t"            \
  "neg." T " zn1, zm1;\n\t"                           \
  "neg." T " zn3, zm3;\n\t"                           \
  "neg." T " zn5, zm5;\n\t"                           \
  "neg." T " zn7, zm7;\n\t"
#define ZQ_LOAD_W(T, O0, O1)                     \
  "ld.param." T " zwr, [zcur+" O0 "];\n\t"       \
  "ld.param." T " zwi, [zcur+" O1 "];\n\t"

#define ZQ_NEXT "bra ZL_NEXT;\n\t"

// Defines  void NAME(uint32_t pe, uint64_t condbits, uint64_t base, uint32_t tbl_saddr):
//   pe = param-space address of the super-op's first record, tbl_saddr = shared address of the PHASEN factor table.
// O0..O7 = byte offsets (strings) of Elem::m[0..7]; WSTRIDE = bytes of one factor-table slot (2 * sizeof(R)).
#define QIP_DEFINE_RUN_ELEMS_PTX(NAME, T, O0, O1, O2, O3, O4, O5, O6, O7, WSTRIDE)                                   \
  __device__ __forceinline__ void NAME(uint32_t pe, uint64_t condbits, uint64_t base, uint32_t tbl_saddr) {          \
    asm volatile(                                                                                                   \
        "{\n\t"                                                                                                     \
        ".reg .b32 zop, zopn, zid, zsz, zcur, zpe, zslot, zt, zpm, zpad, zta;\n\t"                                  \
        ".reg .b64 zcb, zgm, zgv;\n\t"                                                                              \
        ".reg .pred zp, zq0, zq1, zq2, zq3, zq4, zq5, zq6, zq7;\n\t"                                                \
        ".reg ." T " zm0, zm1, zm2, zm3, zm4, zm5, zm6, zm7, zn1, zn3, zn5, zn7, zwr, zwi;\n\t"                     \
        ".reg ." T " zt1, zt2, zt3, zt4, za1, za2, zb1, zb2, ztx;\n\t"                                              \
        /* case id -> label (tile.cuh: enum ElemCase); ids the planner never sends here fall through to NEXT */     \
        "ZTBL: .branchtargets ZL_NEXT, ZL_R0, ZL_R1, ZL_R2, ZL_C0, ZL_C1, ZL_C2, ZL_NEXT, ZL_NEXT, ZL_NEXT, "        \
        "ZL_NEXT, ZL_NEXT, ZL_NEXT, ZL_PHG, ZL_NEXT, ZL_NEXT, ZL_NEXT, ZL_NEXT, ZL_NEXT, ZL_NEXT, "                  \
        "ZL_NEXT, ZL_PHN, ZL_PJ0, ZL_PJ1, ZL_PJ2, ZL_K0, ZL_K1, ZL_K2, ZL_K3, ZL_K4, "                               \
        "ZL_K5, ZL_T0, ZL_T1, ZL_T2, ZL_P20, ZL_P21, ZL_P22, ZL_H0, ZL_H1, ZL_H2;\n\t"                               \
        "mov.u32 zpe, %0;\n\t"                                                                                      \
        "ld.param.u32 zop, [zpe];\n\t"                                                                              \
        "ZL_LOOP:\n\t"                                                                                              \
        "and.b32 zid, zop, 63;\n\t"                                                                                 \
        "setp.eq.u32 zp, zid, 0;\n\t"                                                                               \
        "@zp bra ZL_DONE;\n\t"                                                                                      \
        "shr.u32 zsz, zop, 16;\n\t"                                                                                 \
        "and.b32 zsz, zsz, 32752;\n\t" /* ((op >> 20) & 0x7ff) << 4 */                                              \
        "mov.u32 zcur, zpe;\n\t"                                                                                    \
        "add.u32 zpe, zpe, zsz;\n\t"                                                                                \
        "ld.param.u32 zopn, [zpe];\n\t" /* next op word in flight during this op */                                 \
        "setp.ge.s32 zp, zop, 0;\n\t"   /* bit 31 clear: unconditional */                                           \
        "@zp bra ZL_GO;\n\t"                                                                                        \
        "shr.u32 zslot, zop, 6;\n\t"                                                                                \
        "and.b32 zslot, zslot, 63;\n\t"                                                                             \
        "setp.eq.u32 zp, zslot, 63;\n\t"                                                                            \
        "@zp bra ZL_SLOW;\n\t"                                                                                      \
        "shr.u64 zcb, %1, zslot;\n\t"                                                                               \
        "and.b64 zcb, zcb, 1;\n\t"                                                                                  \
        "setp.eq.u64 zp, zcb, 0;\n\t"                                                                               \
        "@zp bra ZL_NEXT;\n\t"                                                                                      \
        "bra ZL_GO;\n\t"                                                                                            \
        "ZL_SLOW:\n\t" /* condition beyond the table: test the record's own gmask / gval */                         \
        "ld.param.u64 zgm, [zcur+8];\n\t"                                                                           \
        "ld.param.u64 zgv, [zcur+16];\n\t"                                                                          \
        "and.b64 zgm, zgm, %2;\n\t"                                                                                 \
        "setp.ne.u64 zp, zgm, zgv;\n\t"                                                                             \
        "@zp bra ZL_NEXT;\n\t"                                                                                      \
        "ZL_GO:\n\t"                                                                                                \
        "brx.idx zid, ZTBL;\n\t"                                                                                    \
        /* real 2x2, every pair */                                                                                  \
        "ZL_R0:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 0, 1) ZQ_D1R(T, 2, 3) ZQ_D1R(T, 4, 5) ZQ_D1R(T, 6, 7) ZQ_NEXT \
        "ZL_R1:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 0, 2) ZQ_D1R(T, 1, 3) ZQ_D1R(T, 4, 6) ZQ_D1R(T, 5, 7) ZQ_NEXT \
        "ZL_R2:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 0, 4) ZQ_D1R(T, 1, 5) ZQ_D1R(T, 2, 6) ZQ_D1R(T, 3, 7) ZQ_NEXT \
        /* complex 2x2, every pair */                                                                               \
        "ZL_C0:\n\t" ZQ_LOAD_MC(T, O0, O1, O2, O3, O4, O5, O6, O7) ZQ_D1C(T, 0, 1) ZQ_D1C(T, 2, 3) ZQ_D1C(T, 4, 5) ZQ_D1C(T, 6, 7) ZQ_NEXT \
        "ZL_C1:\n\t" ZQ_LOAD_MC(T, O0, O1, O2, O3, O4, O5, O6, O7) ZQ_D1C(T, 0, 2) ZQ_D1C(T, 1, 3) ZQ_D1C(T, 4, 6) ZQ_D1C(T, 5, 7) ZQ_NEXT \
        "ZL_C2:\n\t" ZQ_LOAD_MC(T, O0, O1, O2, O3, O4, O5, O6, O7) ZQ_D1C(T, 0, 4) ZQ_D1C(T, 1, 5) ZQ_D1C(T, 2, 6) ZQ_D1C(T, 3, 7) ZQ_NEXT \
        /* real 2x2 under one in-group control: EC_D1R_C1 + 2*j + w (pairs as in tile_interp.cuh) */                \
        "ZL_K0:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 2, 3) ZQ_D1R(T, 6, 7) ZQ_NEXT                           \
        "ZL_K1:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 4, 5) ZQ_D1R(T, 6, 7) ZQ_NEXT                           \
        "ZL_K2:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 1, 3) ZQ_D1R(T, 5, 7) ZQ_NEXT                           \
        "ZL_K3:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 4, 6) ZQ_D1R(T, 5, 7) ZQ_NEXT                           \
        "ZL_K4:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 1, 5) ZQ_D1R(T, 3, 7) ZQ_NEXT                           \
        "ZL_K5:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 2, 6) ZQ_D1R(T, 3, 7) ZQ_NEXT                           \
        /* both other sub-bits are controls: EC_D1R_C2 + j */                                                       \
        "ZL_T0:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 6, 7) ZQ_NEXT                                           \
        "ZL_T1:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 5, 7) ZQ_NEXT                                           \
        "ZL_T2:\n\t" ZQ_LOAD_MR(T, O0, O1, O2, O3) ZQ_D1R(T, 3, 7) ZQ_NEXT                                           \
        /* phase on every amplitude with sub-bit j set: EC_PHASE_J + j */                                           \
        "ZL_PJ0:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 1) ZQ_PH(T, "", 3) ZQ_PH(T, "", 5) ZQ_PH(T, "", 7) ZQ_NEXT   \
        "ZL_PJ1:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 2) ZQ_PH(T, "", 3) ZQ_PH(T, "", 6) ZQ_PH(T, "", 7) ZQ_NEXT   \
        "ZL_PJ2:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 4) ZQ_PH(T, "", 5) ZQ_PH(T, "", 6) ZQ_PH(T, "", 7) ZQ_NEXT   \
        /* phase on the amplitudes with two sub-bits set: EC_PHASE_2 + q */                                         \
        "ZL_P20:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 3) ZQ_PH(T, "", 7) ZQ_NEXT                                   \
        "ZL_P21:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 5) ZQ_PH(T, "", 7) ZQ_NEXT                                   \
        "ZL_P22:\n\t" ZQ_LOAD_W(T, O0, O1) ZQ_PH(T, "", 6) ZQ_PH(T, "", 7) ZQ_NEXT                                   \
        /* un-normalised Hadamard: EC_HAD + j */                                                                    \
        "ZL_H0:\n\t" ZQ_HAD(T, 0, 1) ZQ_HAD(T, 2, 3) ZQ_HAD(T, 4, 5) ZQ_HAD(T, 6, 7) ZQ_NEXT                         \
        "ZL_H1:\n\t" ZQ_HAD(T, 0, 2) ZQ_HAD(T, 1, 3) ZQ_HAD(T, 4, 6) ZQ_HAD(T, 5, 7) ZQ_NEXT                         \
        "ZL_H2:\n\t" ZQ_HAD(T, 0, 4) ZQ_HAD(T, 1, 5) ZQ_HAD(T, 2, 6) ZQ_HAD(T, 3, 7) ZQ_NEXT                         \
        /* PHASEN: the factor was formed once per CTA (table slot Elem::pad), then as a generic masked phase */      \
        "ZL_PHN:\n\t"                                                                                               \
        "ld.param.u32 zpad, [zcur+4];\n\t"                                                                          \
        "mad.lo.u32 zta, zpad, " WSTRIDE ", %3;\n\t"                                                                \
        "ld.shared.v2." T " {zwr, zwi}, [zta];\n\t"                                                                 \
        "bra ZL_PHM;\n\t"                                                                                           \
        "ZL_PHG:\n\t" ZQ_LOAD_W(T, O0, O1)                                                                          \
        "ZL_PHM:\n\t"                                                                                               \
        "shr.u32 zpm, zop, 12;\n\t"                                                                                 \
        ZQ_PHM(T, 0, 1) ZQ_PHM(T, 1, 2) ZQ_PHM(T, 2, 4) ZQ_PHM(T, 3, 8)                                             \
        ZQ_PHM(T, 4, 16) ZQ_PHM(T, 5, 32) ZQ_PHM(T, 6, 64) ZQ_PHM(T, 7, 128)                                        \
        "ZL_NEXT:\n\t"                                                                                              \
        "mov.u32 zop, zopn;\n\t"                                                                                    \
        "bra ZL_LOOP;\n\t"                                                                                          \
        "ZL_DONE:\n\t"                                                                                              \
        "}" ::"r"(pe),                                                                                              \
        "l"(condbits), "l"(base), "r"(tbl_saddr)                                                                    \
        : "memory");                                                                                                \
  }
