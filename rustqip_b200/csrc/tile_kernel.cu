// tile_kernel.cu -- the fused shared-memory tile pass (see tile.cuh for the model).
//
// One CTA == one tile of 2^T amplitudes (64 KiB).  Phases:
//   1. LOAD   2^m chunks of 2^L contiguous amplitudes, HBM -> shared memory with
//             16-byte cp.async (LDGSTS); the shared-memory image uses the 128-byte XOR
//             swizzle (16-byte unit u lives at u ^ ((u >> 3) & 7)), so that groups of
//             amplitudes at ANY power-of-two stride are at most 2-way bank conflicted.
//   2. APPLY  the pass's micro-ops, one __syncthreads() apiece.  The work-horse is the
//             SUPER-OP: a thread pulls the 8 amplitudes of a 3-bit group into registers
//             and runs a whole list of elementary gates on them before writing back --
//             one shared-memory round trip for many gates.  Gate descriptors arrive as
//             a __grid_constant__ parameter: they are read through the constant bank and
//             never compete with the amplitudes for shared-memory bandwidth.
//   3. STORE  shared memory -> HBM, 16 bytes per lane, same addresses as the load; the
//             product of the CTA-uniform phase terms is folded in here.
// Two CTAs are resident per SM (2 x 64 KiB of shared memory, 128 registers per thread for
// the two register-resident groups), so one CTA's loads and stores overlap the other's
// arithmetic.  HBM traffic per pass: every
// amplitude read once and written once, no matter how many gates the pass folds in.
#include <cuda.h>
#include <cuda_runtime.h>

#include "tile.cuh"
#include "tile_launch.cuh"
#include "tile_interp.cuh"

#include <cstring>

namespace qipb200 {

static const int kTileThreads = 256;

__device__ __forceinline__ uint32_t expand_local(uint32_t g, const MicroOp *mo);

template <typename R>
struct C2;
template <>
struct C2<float> {
  typedef float2 type;
};
template <>
struct C2<double> {
  typedef double2 type;
};

// 128-byte XOR swizzle on element index t (element = one complex<R>)
template <typename R>
__device__ __forceinline__ uint32_t swz(uint32_t t);
template <>
__device__ __forceinline__ uint32_t swz<double>(uint32_t t) {
  return t ^ ((t >> 3) & 7u);
}
template <>
__device__ __forceinline__ uint32_t swz<float>(uint32_t t) {
  return t ^ (((t >> 4) & 7u) << 1);
}

__device__ __forceinline__ uint32_t expand_local(uint32_t g, const MicroOp *mo) {
  uint32_t t = g;
  for (uint32_t i = 0; i < mo->ins_n; ++i) {
    const uint32_t p = mo->ins_pos[i];
    t = ((t >> p) << (p + 1)) | (t & ((1u << p) - 1u));
  }
  return t | mo->lor_mask;
}

// ---- MK_SUPER interpreter: inline PTX on named registers (tile_interp.cuh) ------------------
__device__ __forceinline__ uint32_t swz_d(uint32_t t) { return t ^ ((t >> 3) & 7u); }
__device__ __forceinline__ uint32_t swz_f(uint32_t t) { return t ^ (((t >> 4) & 7u) << 1); }

QIP_DEFINE_RUN_SUPER(run_super_f64, double, "f64", QIP_CD, QIP_COD, swz_d, 4)
QIP_DEFINE_RUN_SUPER(run_super_f32, float, "f32", QIP_CF, QIP_COF, swz_f, 3)

// ---- wide micro-ops (more than 3 involved bits): rare -----------------------------------
template <typename R, int K>
__device__ __forceinline__ void apply_dense(typename C2<R>::type *tile, const MicroOp *mo, const R *mat) {
  typedef typename C2<R>::type V;
  constexpr int S = 1 << K;
  const uint32_t groups = 1u << mo->groups_log2;
  for (uint32_t g = threadIdx.x; g < groups; g += kTileThreads) {
    const uint32_t t0 = expand_local(g, mo);
    V in[S];
    uint32_t addr[S];
#pragma unroll
    for (int u = 0; u < S; ++u) {
      addr[u] = swz<R>(t0 + mo->off[u]);
      in[u] = tile[addr[u]];
    }
#pragma unroll
    for (int u = 0; u < S; ++u) {
      R re = (R)0, im = (R)0;
#pragma unroll
      for (int v = 0; v < S; ++v) {
        const R mr = mat[2 * (u * S + v)], mi = mat[2 * (u * S + v) + 1];
        re = fma(mr, in[v].x, re);
        re = fma(-mi, in[v].y, re);
        im = fma(mr, in[v].y, im);
        im = fma(mi, in[v].x, im);
      }
      V o;
      o.x = re;
      o.y = im;
      tile[addr[u]] = o;
    }
  }
}

template <typename R>
__device__ __forceinline__ void apply_exch(typename C2<R>::type *tile, const MicroOp *mo) {
  typedef typename C2<R>::type V;
  const uint32_t groups = 1u << mo->groups_log2;
  for (uint32_t g = threadIdx.x; g < groups; g += kTileThreads) {
    const uint32_t t0 = expand_local(g, mo);
    const uint32_t a = swz<R>(t0 + mo->off[0]), b = swz<R>(t0 + mo->off[1]);
    const V x = tile[a], y = tile[b];
    tile[a] = y;
    tile[b] = x;
  }
}

template <typename R>
__device__ __forceinline__ void apply_diag(typename C2<R>::type *tile, const MicroOp *mo, const DiagTerm<R> *terms,
                                           uint64_t base, uint32_t T) {
  typedef typename C2<R>::type V;
  const uint32_t n = 1u << T;
  for (uint32_t t = threadIdx.x; t < n; t += kTileThreads) {
    R pr = (R)1, pi = (R)0;
    bool any = false;
    for (uint32_t k = 0; k < mo->nterms; ++k) {
      const DiagTerm<R> &d = terms[k];
      if ((base & d.gmask) != d.gval) continue;  // CTA-uniform
      if ((t & d.lmask) != d.lval) continue;
      const R nr = pr * d.re - pi * d.im;
      pi = pr * d.im + pi * d.re;
      pr = nr;
      any = true;
    }
    if (any) {
      const uint32_t a = swz<R>(t);
      const V v = tile[a];
      V o;
      o.x = fma(pr, v.x, -pi * v.y);
      o.y = fma(pr, v.y, pi * v.x);
      tile[a] = o;
    }
  }
}

// ---- TMA (cp.async.bulk.tensor) + mbarrier helpers -----------------------------------------
// The state is viewed as a 5-D tensor [bits >= h3][h2..h3)[h1..h2)[low3..h1)[128 bytes] (a plain
// reshape of the contiguous buffer, h1 < h2 < h3 = the three lowest high tile bits); one box =
// {128 B, 2^(L-low3), 2, 2, 2} = 2^(L+3) amplitudes lands in shared memory in exactly the
// tile-local index order, and CU_TENSOR_MAP_SWIZZLE_128B is the same XOR swizzle the
// interpreter addresses with.  The remaining m-3 tile bits select 2^(m-3) boxes.
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "DONE:\n\t}" ::"r"(mbar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap *map, uint32_t mbar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(mbar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap *map, uint32_t src, int c0, int c1, int c2, int c3, int c4) {
  asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
               : "memory");
}

// (Round 1 carried four compile-time experiments of this interpreter -- opaque shared-window base, header touch
// before the barrier, descriptor word software-pipelined through a named PTX register, the record loop as one PTX
// block with brx.idx dispatch.  Measured in round 2 (profiles/r2c_interpreter_variants_ab.txt, N=30 f64 circuit):
// 288.8 ms default vs 288.8 / 324.6 / 286.1 / 285.2 / 293.6 ms: nothing beyond 1.2 %.  They were deleted; the pass
// is bound by its shared-memory round trips, and the generated kernels (jit_codegen.cpp) are the product path.)
template <typename R, int G>
__global__ void __launch_bounds__(kTileThreads, (G == 1 ? 3 : 2))
    k_tile_pass(R *__restrict__ psi, const __grid_constant__ PassParams pp, const __grid_constant__ CUtensorMap tmap) {
  typedef typename C2<R>::type V;
  // the named PTX registers that hold the register-resident groups (tile_interp.cuh)
  if (sizeof(R) == 8) {
    QIP_DECL_GROUP("f64", "a");
    if (G == 2) QIP_DECL_GROUP("f64", "b");
  } else {
    QIP_DECL_GROUP("f32", "a");
    if (G == 2) QIP_DECL_GROUP("f32", "b");
  }
  extern __shared__ __align__(1024) unsigned char smem[];
  const PassHeader *h = &pp.h;
  const uint32_t T = h->T, L = h->L, m = h->m, n_ops = h->n_ops;
  const uint32_t tile_bytes = (uint32_t)(2 * sizeof(R)) << T;
  V *tile = reinterpret_cast<V *>(smem);

  uint64_t base = (uint64_t)blockIdx.x << L;
  for (uint32_t i = 0; i < m; ++i) {
    const uint32_t p = h->hi_pos[i];
    base = ((base >> p) << (p + 1)) | (base & ((1ull << p) - 1ull));
  }

  // ---- 1. load ----
  constexpr uint32_t kAmpsPerUnit = 16 / (2 * sizeof(R));
  constexpr uint32_t kLow3 = sizeof(R) == 8 ? 3 : 4;  // index bits covered by one 128-byte row
  const uint32_t units = tile_bytes >> 4;
  const uint32_t lmask = (1u << L) - 1u;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  const bool use_tma = h->use_tma != 0;
  const uint32_t mbar = smem_base + tile_bytes + kMaxPhasen * 16;
  const uint32_t n_boxes = 1u << (m - 3);           // only meaningful with use_tma (m >= 3)
  const uint32_t box_bytes = (uint32_t)(2 * sizeof(R)) << (L + 3);
  if (use_tma) {
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
      mbar_expect_tx(mbar, tile_bytes);
      const uint32_t h1 = h->hi_pos[0], h2 = h->hi_pos[1], h3 = h->hi_pos[2];
      for (uint32_t b = 0; b < n_boxes; ++b) {
        const uint64_t idx = base + h->chunk_off[b << 3];  // chunk index bits 3.. <-> tile bits hi_pos[3..]
        tma_load_5d(smem_base + b * box_bytes, &tmap, mbar, 0, (int)((idx >> kLow3) & ((1ull << (h1 - kLow3)) - 1ull)),
                    (int)((idx >> h1) & ((1ull << (h2 - h1)) - 1ull)), (int)((idx >> h2) & ((1ull << (h3 - h2)) - 1ull)),
                    (int)(idx >> h3));
      }
    }
  } else {
    for (uint32_t u = threadIdx.x; u < units; u += kTileThreads) {
      const uint32_t t = u * kAmpsPerUnit;
      const R *g = psi + 2 * (base + h->chunk_off[t >> L] + (t & lmask));
      const uint32_t sa = smem_base + 16u * (u ^ ((u >> 3) & 7u));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  }

  // While the tile streams in: the per-CTA factor table of the EC_PHASEN ops (product of the
  // conditional factors whose condition this tile's base index satisfies).  Thread t walks
  // micro-op t; the table sits behind the tile in shared memory.
  R *tbl = reinterpret_cast<R *>(smem + tile_bytes);
  if (h->n_phasen && threadIdx.x < n_ops) {
    const unsigned char *rec = pp.recs;
    for (uint32_t i = 0; i < threadIdx.x; ++i) rec += sizeof(MicroOp) + reinterpret_cast<const MicroOp *>(rec)->data_bytes;
    const MicroOp *mo = reinterpret_cast<const MicroOp *>(rec);
    if (mo->kind == MK_SUPER) {
      const unsigned char *ep = rec + sizeof(MicroOp);
      for (;;) {
        const Elem<R> *e = reinterpret_cast<const Elem<R> *>(ep);
        const uint32_t op = e->op;
        if ((op & kElemCaseMask) == EC_END) break;
        const uint32_t size = ((op >> 20) & 0x7ffu) << 4;
        if ((op & kElemCaseMask) == EC_PHASEN) {
          R wr = e->m[0], wi = e->m[1];
          const PhaseTerm<R> *pt = reinterpret_cast<const PhaseTerm<R> *>(e + 1);
          const uint32_t nt = (size - (uint32_t)sizeof(Elem<R>)) / (uint32_t)sizeof(PhaseTerm<R>);
          for (uint32_t k = 0; k < nt; ++k) {
            if ((base & pt[k].gmask) != pt[k].gval) continue;
            const R nr = wr * pt[k].re - wi * pt[k].im;
            wi = wr * pt[k].im + wi * pt[k].re;
            wr = nr;
          }
          tbl[2 * e->pad] = wr;
          tbl[2 * e->pad + 1] = wi;
        }
        ep += size;
      }
    }
  }
  // ... and this CTA's evaluation of the pass's condition table (controls that live outside the
  // tile): thread s tests condition s, two ballots give the 64-bit word every thread keeps.
  uint32_t *condw = reinterpret_cast<uint32_t *>(smem + tile_bytes + kMaxPhasen * 16 + 16);
  if (threadIdx.x < 64) {
    bool on = false;
    if (threadIdx.x < h->n_conds) {
      const CondTerm *ct = reinterpret_cast<const CondTerm *>(pp.recs + h->cond_off) + threadIdx.x;
      on = (base & ct->gmask) == ct->gval;
    }
    const uint32_t bits = __ballot_sync(0xffffffffu, on);
    if ((threadIdx.x & 31u) == 0) condw[threadIdx.x >> 5] = bits;
  }
  if (use_tma)
    mbar_wait(mbar, 0);
  else
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();
  const uint64_t condbits = (uint64_t)condw[0] | ((uint64_t)condw[1] << 32);

  // ---- 2. apply ----
  const unsigned char *rec = pp.recs;
  for (uint32_t i = 0; i < n_ops; ++i) {
    const MicroOp *mo = reinterpret_cast<const MicroOp *>(rec);
    const unsigned char *data = rec + sizeof(MicroOp);
    rec = data + mo->data_bytes;
    if ((base & mo->gmask) == mo->gmask) {
      if (mo->kind == MK_SUPER) {
        if constexpr (sizeof(R) == 8)
          run_super_f64<G>(smem_base, mo, data, base, reinterpret_cast<const double *>(tbl), condbits);
        else
          run_super_f32<G>(smem_base, mo, data, base, reinterpret_cast<const float *>(tbl), condbits);
      } else if (mo->kind == MK_DENSE) {
        const R *mat = reinterpret_cast<const R *>(data);
        if (mo->k == 1)
          apply_dense<R, 1>(tile, mo, mat);
        else if (mo->k == 2)
          apply_dense<R, 2>(tile, mo, mat);
        else
          apply_dense<R, 3>(tile, mo, mat);
      } else if (mo->kind == MK_DIAG) {
        apply_diag<R>(tile, mo, reinterpret_cast<const DiagTerm<R> *>(data), base, T);
      } else {
        apply_exch<R>(tile, mo);
      }
    }
    __syncthreads();
  }

  // ---- 3. store (with the CTA-uniform phase product folded in) ----
  R gr = (R)1, gi = (R)0;
  bool has_g = false;
  {
    const GlobalTerm<R> *gt = reinterpret_cast<const GlobalTerm<R> *>(pp.recs + h->gterm_off);
    for (uint32_t k = 0; k < h->n_gterms; ++k) {
      if ((base & gt[k].gmask) != gt[k].gval) continue;
      const R nr = gr * gt[k].re - gi * gt[k].im;
      gi = gr * gt[k].im + gi * gt[k].re;
      gr = nr;
      has_g = true;
    }
  }
  if (use_tma && !has_g) {
    // shared memory -> HBM with TMA tensor stores (same boxes, same swizzle)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
      const uint32_t h1 = h->hi_pos[0], h2 = h->hi_pos[1], h3 = h->hi_pos[2];
      for (uint32_t b = 0; b < n_boxes; ++b) {
        const uint64_t idx = base + h->chunk_off[b << 3];
        tma_store_5d(&tmap, smem_base + b * box_bytes, 0, (int)((idx >> kLow3) & ((1ull << (h1 - kLow3)) - 1ull)),
                     (int)((idx >> h1) & ((1ull << (h2 - h1)) - 1ull)), (int)((idx >> h2) & ((1ull << (h3 - h2)) - 1ull)),
                     (int)(idx >> h3));
      }
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    return;
  }
  for (uint32_t u = threadIdx.x; u < units; u += kTileThreads) {
    const uint32_t t = u * kAmpsPerUnit;
    R *g = psi + 2 * (base + h->chunk_off[t >> L] + (t & lmask));
    const unsigned char *sp = smem + 16u * (u ^ ((u >> 3) & 7u));
    if (has_g) {
      if (sizeof(R) == 8) {
        double2 v = *reinterpret_cast<const double2 *>(sp);
        double2 o;
        o.x = fma((double)gr, v.x, -(double)gi * v.y);
        o.y = fma((double)gr, v.y, (double)gi * v.x);
        *reinterpret_cast<double2 *>(g) = o;
      } else {
        float4 v = *reinterpret_cast<const float4 *>(sp);
        float4 o;
        o.x = fmaf((float)gr, v.x, -(float)gi * v.y);
        o.y = fmaf((float)gr, v.y, (float)gi * v.x);
        o.z = fmaf((float)gr, v.z, -(float)gi * v.w);
        o.w = fmaf((float)gr, v.w, (float)gi * v.z);
        *reinterpret_cast<float4 *>(g) = o;
      }
    } else {
      *reinterpret_cast<uint4 *>(g) = *reinterpret_cast<const uint4 *>(sp);
    }
  }
}

cudaError_t tile_pass_configure() {
  cudaError_t e;
  const void *fns[] = {(const void *)k_tile_pass<double, 1>, (const void *)k_tile_pass<double, 2>,
                       (const void *)k_tile_pass<float, 1>, (const void *)k_tile_pass<float, 2>};
  for (const void *f : fns)
    if ((e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  return cudaSuccess;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// Tensor map of the local state for one pass (see the comment above mbar_init).  Returns false
// when the geometry does not fit (then the pass runs with the cp.async path).
bool make_tile_map(CUtensorMap *map, qip_prec prec, void *psi, uint32_t n_local, const PassHeader &h) {
  EncodeTiledFn enc = encode_tiled_fn();
  const uint32_t low3 = prec == QIP_F64 ? 3 : 4;
  if (!enc || h.m < 3 || h.L < low3 || h.T > n_local) return false;
  const uint32_t h1 = h.hi_pos[0], h2 = h.hi_pos[1], h3 = h.hi_pos[2];
  const uint64_t amp = prec == QIP_F64 ? 16 : 8;
  const uint64_t esz = prec == QIP_F64 ? 8 : 4;
  cuuint64_t dims[5] = {128 / esz, 1ull << (h1 - low3), 1ull << (h2 - h1), 1ull << (h3 - h2), 1ull << (n_local - h3)};
  cuuint64_t strides[4] = {128, amp << h1, amp << h2, amp << h3};
  cuuint32_t box[5] = {(cuuint32_t)(128 / esz), 1u << (h.L - low3), 2, 2, 2};
  cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  if (box[1] > 256 || dims[4] > (1ull << 32)) return false;
  const CUresult r = enc(map, prec == QIP_F64 ? CU_TENSOR_MAP_DATA_TYPE_UINT64 : CU_TENSOR_MAP_DATA_TYPE_UINT32, 5, psi, dims,
                         strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

cudaError_t launch_tile_pass(qip_prec prec, void *psi, uint32_t n_local, PassParams &pp, int groups_per_thread,
                             bool use_tma, cudaStream_t s, uint64_t *launches) {
  const uint32_t T = pp.h.T;
  alignas(64) CUtensorMap tmap;
  memset(&tmap, 0, sizeof(tmap));
  pp.h.use_tma = (use_tma && make_tile_map(&tmap, prec, psi, n_local, pp.h)) ? 1u : 0u;
  // tile | EC_PHASEN factor table | mbarrier | condition word
  const size_t smem = ((size_t)(prec == QIP_F32 ? 8u : 16u) << T) + kMaxPhasen * 16 + 32;
  const unsigned grid = 1u << (n_local - T);
  if (prec == QIP_F32) {
    if (groups_per_thread == 2)
      k_tile_pass<float, 2><<<grid, kTileThreads, smem, s>>>((float *)psi, pp, tmap);
    else
      k_tile_pass<float, 1><<<grid, kTileThreads, smem, s>>>((float *)psi, pp, tmap);
  } else {
    if (groups_per_thread == 2)
      k_tile_pass<double, 2><<<grid, kTileThreads, smem, s>>>((double *)psi, pp, tmap);
    else
      k_tile_pass<double, 1><<<grid, kTileThreads, smem, s>>>((double *)psi, pp, tmap);
  }
  ++*launches;
  return cudaGetLastError();
}

}  // namespace qipb200
