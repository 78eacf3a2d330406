// tile_kernel.cu -- the fused shared-memory tile pass (see tile.cuh for the model).
//
// One CTA == one tile of 2^T amplitudes (64 KiB).  Phases:
//   1. LOAD   2^m chunks of 2^L contiguous amplitudes, HBM -> shared memory with
//             16-byte cp.async (LDGSTS); the shared-memory image uses the 128-byte XOR
//             swizzle (16-byte unit u lives at u ^ ((u >> 3) & 7)), so that groups of
//             amplitudes at ANY power-of-two stride are at most 2-way bank conflicted.
//   2. APPLY  the pass's micro-ops in shared memory, one __syncthreads() apiece; the
//             next micro-op's matrix / phase terms are staged into a double buffer
//             while the current one runs.
//   3. STORE  shared memory -> HBM, 16 bytes per lane, same addresses as the load.
// Three CTAs are resident per SM (3 x 66.5 KiB of shared memory), so one CTA's loads
// and stores overlap the other CTAs' arithmetic.  HBM traffic per pass: every
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

template <typename R, int K>
__device__ __forceinline__ void apply_dense(typename C2<R>::type *tile, const MicroOp *mo, const R *mat) {
  typedef typename C2<R>::type V;
  constexpr int S = 1 << K;
  const uint32_t groups = 1u << mo->groups_log2;
  const V *m2 = reinterpret_cast<const V *>(mat);
  if (K == 1 || (K == 2 && sizeof(R) == 4)) {
    // small block: keep the matrix in registers for all groups of this thread
    V mr[S * S];
#pragma unroll
    for (int i = 0; i < S * S; ++i) mr[i] = m2[i];
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
          const V mm = mr[u * S + v];
          re = fma(mm.x, in[v].x, re);
          re = fma(-mm.y, in[v].y, re);
          im = fma(mm.x, in[v].y, im);
          im = fma(mm.y, in[v].x, im);
        }
        V o;
        o.x = re;
        o.y = im;
        tile[addr[u]] = o;
      }
    }
  } else {
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
          const V mm = m2[u * S + v];  // broadcast read from the staged matrix
          re = fma(mm.x, in[v].x, re);
          re = fma(-mm.y, in[v].y, re);
          im = fma(mm.x, in[v].y, im);
          im = fma(mm.y, in[v].x, im);
        }
        V o;
        o.x = re;
        o.y = im;
        tile[addr[u]] = o;
      }
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

__device__ __forceinline__ void stage_record(unsigned char *dst, const unsigned char *src, uint32_t bytes) {
  const uint32_t units = bytes >> 4;
  for (uint32_t i = threadIdx.x; i < units; i += kTileThreads)
    reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
}

template <typename R>
__global__ void __launch_bounds__(kTileThreads, 3)
    k_tile_pass(R *__restrict__ psi, const unsigned char *__restrict__ blob) {
  typedef typename C2<R>::type V;
  extern __shared__ __align__(1024) unsigned char smem[];
  const PassHeader *h = reinterpret_cast<const PassHeader *>(blob);
  const uint32_t T = h->T, L = h->L, m = h->m, n_ops = h->n_ops;
  const uint32_t tile_bytes = (uint32_t)(2 * sizeof(R)) << T;
  V *tile = reinterpret_cast<V *>(smem);
  unsigned char *stage = smem + tile_bytes;

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

  const unsigned char *rec = blob + sizeof(PassHeader);
  uint32_t rec_bytes = 0;
  if (n_ops) {
    rec_bytes = (uint32_t)sizeof(MicroOp) + reinterpret_cast<const MicroOp *>(rec)->data_bytes;
    stage_record(stage, rec, rec_bytes);
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- 2. apply ----
  for (uint32_t i = 0; i < n_ops; ++i) {
    unsigned char *cur = stage + (i & 1u) * kTileStageBytes;
    const MicroOp *mo = reinterpret_cast<const MicroOp *>(cur);
    if (i + 1 < n_ops) {  // stage the next record while this one runs
      rec += rec_bytes;
      rec_bytes = (uint32_t)sizeof(MicroOp) + reinterpret_cast<const MicroOp *>(rec)->data_bytes;
      stage_record(stage + ((i + 1) & 1u) * kTileStageBytes, rec, rec_bytes);
    }
    if ((base & mo->gmask) == mo->gmask) {
      const unsigned char *data = cur + sizeof(MicroOp);
      if (mo->kind == MK_DENSE) {
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

  // ---- 3. store ----
  for (uint32_t u = threadIdx.x; u < units; u += kTileThreads) {
    const uint32_t t = u * kAmpsPerUnit;
    R *g = psi + 2 * (base + h->chunk_off[t >> L] + (t & lmask));
    const uint4 v = *reinterpret_cast<const uint4 *>(smem + 16u * (u ^ ((u >> 3) & 7u)));
    *reinterpret_cast<uint4 *>(g) = v;
  }
}

cudaError_t tile_pass_configure() {
  cudaError_t e = cudaFuncSetAttribute(k_tile_pass<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_tile_pass<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
}

cudaError_t launch_tile_pass(qip_prec prec, void *psi, uint32_t n_local, uint32_t T, const unsigned char *d_blob,
                             cudaStream_t s, uint64_t *launches) {
  const size_t tile_bytes = (prec == QIP_F32 ? 8u : 16u) << T;
  const size_t smem = tile_bytes + 2 * kTileStageBytes;
  const unsigned grid = 1u << (n_local - T);
  if (prec == QIP_F32)
    k_tile_pass<float><<<grid, kTileThreads, smem, s>>>((float *)psi, d_blob);
  else
    k_tile_pass<double><<<grid, kTileThreads, smem, s>>>((double *)psi, d_blob);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace qipb200
