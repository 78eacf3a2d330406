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
#include <cuda_runtime.h>

#include "tile.cuh"
#include "tile_launch.cuh"

namespace qipb200 {

static const int kTileThreads = 256;

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

// ---- elementary ops on 8 register-resident amplitudes -----------------------------------
template <typename R>
struct Amp8 {
  R re[8], im[8];
};

// The p-th (ascending) sub-index with bit J clear.
template <int J>
__device__ __forceinline__ constexpr int pair_base(int p) {
  return ((p >> J) << (J + 1)) | (p & ((1 << J) - 1));
}

// Every elementary op is an IN-PLACE update of register-resident amplitudes -- no op moves
// a value from one register to another (X and SWAP are issued by the planner as exact
// 0/1 real 2x2 gates), so the register assignment of the G groups is identical on every
// path through the interpreter and no copies are needed at the loop back-edge.
// Each op is applied to G groups at once: the descriptor is decoded and its matrix
// fetched once per G groups.
template <typename R, int G, int J, bool FULL>
__device__ __forceinline__ void e_dense1c(Amp8<R> (&a)[G], const Elem<R> *e, uint32_t pm) {
  const R m00r = e->m[0], m00i = e->m[1], m01r = e->m[2], m01i = e->m[3];
  const R m10r = e->m[4], m10i = e->m[5], m11r = e->m[6], m11i = e->m[7];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (!FULL && !((pm >> p) & 1u)) continue;  // CTA-uniform (a control inside the group)
    const int i0 = pair_base<J>(p), i1 = i0 | (1 << J);
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const R xr = a[q].re[i0], xi = a[q].im[i0], yr = a[q].re[i1], yi = a[q].im[i1];
      a[q].re[i0] = fma(m00r, xr, fma(-m00i, xi, fma(m01r, yr, -m01i * yi)));
      a[q].im[i0] = fma(m00r, xi, fma(m00i, xr, fma(m01r, yi, m01i * yr)));
      a[q].re[i1] = fma(m10r, xr, fma(-m10i, xi, fma(m11r, yr, -m11i * yi)));
      a[q].im[i1] = fma(m10r, xi, fma(m10i, xr, fma(m11r, yi, m11i * yr)));
    }
  }
}

template <typename R, int G, int J, bool FULL>
__device__ __forceinline__ void e_dense1r(Amp8<R> (&a)[G], const Elem<R> *e, uint32_t pm) {
  const R m00 = e->m[0], m01 = e->m[1], m10 = e->m[2], m11 = e->m[3];
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    if (!FULL && !((pm >> p) & 1u)) continue;
    const int i0 = pair_base<J>(p), i1 = i0 | (1 << J);
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const R xr = a[q].re[i0], xi = a[q].im[i0], yr = a[q].re[i1], yi = a[q].im[i1];
      a[q].re[i0] = fma(m00, xr, m01 * yr);
      a[q].im[i0] = fma(m00, xi, m01 * yi);
      a[q].re[i1] = fma(m10, xr, m11 * yr);
      a[q].im[i1] = fma(m10, xi, m11 * yi);
    }
  }
}

template <typename R, int G>
__device__ __forceinline__ void e_phase(Amp8<R> (&a)[G], const Elem<R> *e, uint32_t am) {
  const R wr = e->m[0], wi = e->m[1];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    if (!((am >> c) & 1u)) continue;
#pragma unroll
    for (int q = 0; q < G; ++q) {
      const R xr = a[q].re[c], xi = a[q].im[c];
      a[q].re[c] = fma(wr, xr, -wi * xi);
      a[q].im[c] = fma(wr, xi, wi * xr);
    }
  }
}

template <typename R>
__device__ __forceinline__ void e_dense3(Amp8<R> &a, const R *m) {
  Amp8<R> o;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    R re = (R)0, im = (R)0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const R mr = m[2 * (u * 8 + v)], mi = m[2 * (u * 8 + v) + 1];
      re = fma(mr, a.re[v], re);
      re = fma(-mi, a.im[v], re);
      im = fma(mr, a.im[v], im);
      im = fma(mi, a.re[v], im);
    }
    o.re[u] = re;
    o.im[u] = im;
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    a.re[u] = o.re[u];
    a.im[u] = o.im[u];
  }
}

// Interpreter opcode = kind * 4 + j, + 32 when every pair is active (the common case: no
// mask tests, no selects).
template <typename R, int G>
__device__ __forceinline__ void run_super(typename C2<R>::type *tile, const MicroOp *mo, const unsigned char *data,
                                          uint64_t base) {
  typedef typename C2<R>::type V;
  const uint32_t groups = 1u << mo->groups_log2;
  for (uint32_t g = threadIdx.x; g < groups; g += G * kTileThreads) {
    uint32_t addr[G][8];
    Amp8<R> a[G];
    bool valid[G];
#pragma unroll
    for (int q = 0; q < G; ++q) {
      valid[q] = g + q * kTileThreads < groups;  // warp-uniform (groups is a power of two)
      const uint32_t t0 = expand_local(valid[q] ? g + q * kTileThreads : g, mo);
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        addr[q][u] = swz<R>(t0 + mo->off[u]);
        const V v = tile[addr[q][u]];
        a[q].re[u] = v.x;
        a[q].im[u] = v.y;
      }
    }
    const unsigned char *ep = data;
    for (uint32_t ei = 0; ei < mo->nterms; ++ei) {
      const Elem<R> *e = reinterpret_cast<const Elem<R> *>(ep);
      const uint32_t op = e->op;
      const uint32_t code = op & 0xffu, mask = (op >> 12) & 0xffu;
      const bool is_d3 = code == E_DENSE3 * 4;
      ep += sizeof(Elem<R>) + (is_d3 ? 128 * sizeof(R) : 0);
      const bool on = !(op & kElemHasCond) || (base & e->gmask) == e->gval;  // a control outside the tile
      if (on) {
        switch (code) {
          case 32 + E_DENSE1R * 4 + 0: e_dense1r<R, G, 0, true>(a, e, 0xfu); break;
          case 32 + E_DENSE1R * 4 + 1: e_dense1r<R, G, 1, true>(a, e, 0xfu); break;
          case 32 + E_DENSE1R * 4 + 2: e_dense1r<R, G, 2, true>(a, e, 0xfu); break;
          case 32 + E_DENSE1 * 4 + 0: e_dense1c<R, G, 0, true>(a, e, 0xfu); break;
          case 32 + E_DENSE1 * 4 + 1: e_dense1c<R, G, 1, true>(a, e, 0xfu); break;
          case 32 + E_DENSE1 * 4 + 2: e_dense1c<R, G, 2, true>(a, e, 0xfu); break;
          case E_DENSE1R * 4 + 0: e_dense1r<R, G, 0, false>(a, e, mask); break;
          case E_DENSE1R * 4 + 1: e_dense1r<R, G, 1, false>(a, e, mask); break;
          case E_DENSE1R * 4 + 2: e_dense1r<R, G, 2, false>(a, e, mask); break;
          case E_DENSE1 * 4 + 0: e_dense1c<R, G, 0, false>(a, e, mask); break;
          case E_DENSE1 * 4 + 1: e_dense1c<R, G, 1, false>(a, e, mask); break;
          case E_DENSE1 * 4 + 2: e_dense1c<R, G, 2, false>(a, e, mask); break;
          case E_PHASE * 4: e_phase<R, G>(a, e, mask); break;
          default: {
            const R *m8 = reinterpret_cast<const R *>(e + 1);
#pragma unroll
            for (int q = 0; q < G; ++q) e_dense3<R>(a[q], m8);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < G; ++q) {
      if (!valid[q]) continue;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        V v;
        v.x = a[q].re[u];
        v.y = a[q].im[u];
        tile[addr[q][u]] = v;
      }
    }
  }
}

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

template <typename R, int G>
__global__ void __launch_bounds__(kTileThreads, (G == 1 ? 3 : 2))
    k_tile_pass(R *__restrict__ psi, const __grid_constant__ PassParams pp) {
  typedef typename C2<R>::type V;
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
  const uint32_t units = tile_bytes >> 4;
  const uint32_t lmask = (1u << L) - 1u;
  const uint32_t smem_base = (uint32_t)__cvta_generic_to_shared(smem);
  for (uint32_t u = threadIdx.x; u < units; u += kTileThreads) {
    const uint32_t t = u * kAmpsPerUnit;
    const R *g = psi + 2 * (base + h->chunk_off[t >> L] + (t & lmask));
    const uint32_t sa = smem_base + 16u * (u ^ ((u >> 3) & 7u));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- 2. apply ----
  const unsigned char *rec = pp.recs;
  for (uint32_t i = 0; i < n_ops; ++i) {
    const MicroOp *mo = reinterpret_cast<const MicroOp *>(rec);
    const unsigned char *data = rec + sizeof(MicroOp);
    rec = data + mo->data_bytes;
    if ((base & mo->gmask) == mo->gmask) {
      if (mo->kind == MK_SUPER) {
        run_super<R, G>(tile, mo, data, base);
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
  if ((e = cudaFuncSetAttribute(k_tile_pass<double, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_tile_pass<double, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(k_tile_pass<float, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_tile_pass<float, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

cudaError_t launch_tile_pass(qip_prec prec, void *psi, uint32_t n_local, const PassParams &pp, int groups_per_thread,
                             cudaStream_t s, uint64_t *launches) {
  const uint32_t T = pp.h.T;
  const size_t smem = (prec == QIP_F32 ? 8u : 16u) << T;
  const unsigned grid = 1u << (n_local - T);
  if (prec == QIP_F32) {
    if (groups_per_thread == 1)
      k_tile_pass<float, 1><<<grid, kTileThreads, smem, s>>>((float *)psi, pp);
    else
      k_tile_pass<float, 2><<<grid, kTileThreads, smem, s>>>((float *)psi, pp);
  } else {
    if (groups_per_thread == 1)
      k_tile_pass<double, 1><<<grid, kTileThreads, smem, s>>>((double *)psi, pp);
    else
      k_tile_pass<double, 2><<<grid, kTileThreads, smem, s>>>((double *)psi, pp);
  }
  ++*launches;
  return cudaGetLastError();
}

}  // namespace qipb200
