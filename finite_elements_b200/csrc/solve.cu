// Boundary conditions in place, CSR SpMV and the Jacobi-preconditioned CG.
//
// Replaces, on the eliminated SPD system, scipy.sparse.linalg.spsolve at
// analysis.py:820-822 (SuperLU with NATURAL ordering: O(n * bandwidth^2), infeasible at
// 16 M triangles) and the Lagrange rows of analysis.py:241-279 / :509-543.
//
// PCG iteration = 3 kernels, all scalars stay on the device:
//   k_spmv<DOT>   q = A p, per-CTA partials of p.q, last CTA sums them in fixed order
//   k_pcg_update  alpha = rz/pq; x += alpha p; r -= alpha q; partials of r.D^-1 r and r.r
//   k_pcg_pupdate beta = rz'/rz; p = D^-1 r + beta p; convergence flag for later kernels
// HBM bytes per iteration: 12 nnz + 108 n.  After convergence the remaining
// launches of a check interval exit at their first instruction.
// Multi-GPU (dist.cu) inserts a halo exchange before k_spmv and an all-reduce after the
// two reducing kernels; the kernels themselves are the same.
#include <math.h>
#include <stdlib.h>

#include "common.cuh"
#include "dist.h"
#include "ptx.cuh"

namespace fe {

// ---------------------------------------------------------------------------------------
// device-resident solver state
// ---------------------------------------------------------------------------------------
struct PcgState {
  double sums[8];   // [0] pq  [1] rz_new  [2] rr  [3] bnorm2   (all-reduced when nranks > 1)
  double rz[2];     // r.z, double-buffered by iteration parity
  double tol2;      // rtol^2
  int iters;        // completed iterations
  int converged;
  int breakdown;
  unsigned ticket;  // last-CTA election
  P2PDev *pp;       // peer-memory all-reduce (NULL: single GPU, or NCCL between the kernels)
  double prof[6];   // persistent kernel, CTA 0: ns spent in SpMV | barrier+reduce | vector update | barrier | init | -
};

// ---- peer-memory all-reduce, fused into the reducing kernels ------------------------------
// Producer side (threads 0 .. nranks-1 of the last CTA, `tot` in shared memory): thread r stores
// this rank's NQ totals into rank r's cells for all-reduce number red_seq + 1.  No fence, no flag:
// the cells carry the sequence number (LL protocol, common.cuh).
template <int NQ>
__device__ __forceinline__ void p2p_push(P2PDev *pp, const double *tot /* shared, NQ */) {
  const unsigned seq = pp->red_seq + 1;
  const int par = (int)(seq & 1), R = pp->nranks, me = pp->rank;
  if ((int)threadIdx.x < R) {
    uint4 *dst = pp->red[threadIdx.x] + ((size_t)par * R + me) * 4;
#pragma unroll
    for (int q = 0; q < NQ; ++q) ll_store(dst + q, tot[q], seq);
  }
  __syncthreads();
  if (threadIdx.x == 0) pp->red_seq = seq;
}
// Block-wide: the all-reduced values of sums[base .. base+NQ), whichever transport is active.
// Peer memory: thread r waits for rank r's cells of the most recent all-reduce; thread 0 adds them
// in rank order -- every CTA of every rank gets bit-identical sums.
template <int NQ>
__device__ __forceinline__ void reduced_sums(PcgState *st, int base, double *out /* NQ */) {
  __shared__ double part[kMaxRanks][NQ];
  __shared__ double bc[NQ];
  if (st->pp) {
    const P2PDev *pp = st->pp;
    const unsigned seq = pp->red_seq;
    const int par = (int)(seq & 1), R = pp->nranks, me = pp->rank;
    if ((int)threadIdx.x < R) {
      const uint4 *src = pp->red[me] + ((size_t)par * R + threadIdx.x) * 4;
#pragma unroll
      for (int q = 0; q < NQ; ++q) part[threadIdx.x][q] = ll_wait(src + q, seq);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) {
        double t = 0.0;
        for (int r = 0; r < R; ++r) t += part[r][q];
        bc[q] = t;
        if (blockIdx.x == 0) st->sums[base + q] = t;  // for the host read-back
      }
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = bc[q];
    __syncthreads();
  } else {
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = st->sums[base + q];
  }
}

constexpr int kRedBlock = 256;
constexpr int kMaxPartials = 148 * 16;  // grids of the reducing kernels are capped to this

// Sum `np` per-CTA partials (stride = number of quantities) in a fixed order.
template <int NQ, int BLOCK = kRedBlock>
__device__ __forceinline__ void final_reduce(const double *__restrict__ partials, int np, double *smem,
                                             double *__restrict__ out, P2PDev *pp) {
  double acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.0;
  for (int i = threadIdx.x; i < np; i += BLOCK) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) acc[q] += __ldcg(partials + (size_t)i * NQ + q);
  }
  __shared__ double tot_s[NQ];
  double tot[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) tot[q] = block_sum<BLOCK>(acc[q], smem);
  if (pp) {  // all-reduce over peer memory: consumers gather (reduced_sums)
    if (threadIdx.x == 0) {
#pragma unroll
      for (int q = 0; q < NQ; ++q) tot_s[q] = tot[q];
    }
    __syncthreads();
    p2p_push<NQ>(pp, tot_s);
  } else if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) out[q] = tot[q];
  }
}

// Writes this CTA's partials, elects the last CTA, which reduces all of them.
template <int NQ, int BLOCK = kRedBlock>
__device__ __forceinline__ void publish_and_reduce(const double (&local)[NQ], double *__restrict__ partials,
                                                   PcgState *__restrict__ st, int out_base, double *smem) {
  __shared__ bool is_last;
  double tot[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) tot[q] = block_sum<BLOCK>(local[q], smem);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) partials[(size_t)blockIdx.x * NQ + q] = tot[q];
    __threadfence();
    const unsigned t = atomicAdd(&st->ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last) {
    __threadfence();
    final_reduce<NQ, BLOCK>(partials, gridDim.x, smem, st->sums + out_base, st->pp);
    if (threadIdx.x == 0) st->ticket = 0;
  }
}

// ---------------------------------------------------------------------------------------
// SpMV: LPR lanes per row (sub-warp), rows strided over a capped grid
// ---------------------------------------------------------------------------------------
template <int LPR, bool DOT>
__global__ void __launch_bounds__(kRedBlock) k_spmv(int32_t n_rows, const int32_t *__restrict__ rowptr,
                                                   const int32_t *__restrict__ colidx,
                                                   const double *__restrict__ vals, const double *__restrict__ x,
                                                   double *__restrict__ y, double *__restrict__ partials,
                                                   PcgState *__restrict__ st) {
  __shared__ double red[kRedBlock / 32];
  if (DOT && (st->converged | st->breakdown)) return;
  const int lane = threadIdx.x % LPR;
  const int64_t rows_per_pass = (int64_t)gridDim.x * (kRedBlock / LPR);
  double dot = 0.0;
  for (int64_t row = (int64_t)blockIdx.x * (kRedBlock / LPR) + threadIdx.x / LPR;
       row < (int64_t)((n_rows + kRedBlock / LPR - 1) / (kRedBlock / LPR)) * (kRedBlock / LPR); row += rows_per_pass) {
    double acc = 0.0;
    if (row < n_rows) {
      const int32_t s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
      for (int32_t j = s + lane; j < e; j += LPR) acc += vals[j] * __ldg(x + __ldg(colidx + j));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (row < n_rows && lane == 0) {
      y[row] = acc;
      if (DOT) dot += acc * x[row];
    }
  }
  if (DOT) {
    const double loc[1] = {dot};
    publish_and_reduce<1>(loc, partials, st, 0, red);
  }
}

// ---------------------------------------------------------------------------------------
// SpMV for node-blocked matrices (block_dim = 2): rows 2i and 2i+1 share one column list made
// of (2m, 2m+1) pairs -- exactly what fe_plan emits for 2 DOF per node, and what
// fe_dirichlet_apply preserves.  8 lanes per node; a lane owns one 2x2 block: ONE column
// index, one 128-bit load of x, two 128-bit loads of vals, four FMAs.  Only the even entries
// of the first row's colidx are read (half of the index bytes never leave HBM).
// ---------------------------------------------------------------------------------------
template <bool DOT>
__global__ void __launch_bounds__(kRedBlock) k_spmv_b2(int32_t n_nodes, const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx,
                                                      const double *__restrict__ vals, const double *__restrict__ x,
                                                      double *__restrict__ y, double *__restrict__ partials,
                                                      PcgState *__restrict__ st) {
  __shared__ double red[kRedBlock / 32];
  if (DOT && (st->converged | st->breakdown)) return;
  constexpr int LPN = 8, NPB = kRedBlock / LPN, U = 2;  // U nodes in flight per lane group
  const int lane = threadIdx.x % LPN;
  const int64_t stride = (int64_t)gridDim.x * NPB;
  const int64_t n_pad = (int64_t)((n_nodes + NPB - 1) / NPB) * NPB;
  double dot = 0.0;
  for (int64_t node0 = (int64_t)blockIdx.x * NPB + threadIdx.x / LPN; node0 < n_pad; node0 += U * stride) {
    int32_t s0[U], len[U];
    double a0[U], a1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int64_t node = node0 + u * stride;
      a0[u] = a1[u] = 0.0;
      s0[u] = 0;
      len[u] = 0;
      if (node < n_nodes) {
        s0[u] = __ldg(rowptr + 2 * node);
        len[u] = (__ldg(rowptr + 2 * node + 2) - s0[u]) >> 1;  // entries per row = 2 * valence
      }
    }
    // first (usually only) block of every node in flight: all loads are issued before any FMA
    int32_t c[U];
    double2 xv[U], v0[U], v1[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool on = 2 * lane < len[u];
      c[u] = on ? __ldg(colidx + s0[u] + 2 * lane) : 0;
      v0[u] = on ? __ldg(reinterpret_cast<const double2 *>(vals + s0[u] + 2 * lane)) : make_double2(0.0, 0.0);
      v1[u] = on ? __ldg(reinterpret_cast<const double2 *>(vals + s0[u] + len[u] + 2 * lane)) : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int u = 0; u < U; ++u) xv[u] = __ldg(reinterpret_cast<const double2 *>(x + c[u]));
#pragma unroll
    for (int u = 0; u < U; ++u) {
      a0[u] = v0[u].x * xv[u].x + v0[u].y * xv[u].y;
      a1[u] = v1[u].x * xv[u].x + v1[u].y * xv[u].y;
      for (int32_t k = 2 * (lane + LPN); k < len[u]; k += 2 * LPN) {  // valence > 8
        const int32_t cc = __ldg(colidx + s0[u] + k);
        const double2 xx = __ldg(reinterpret_cast<const double2 *>(x + cc));
        const double2 w0 = __ldg(reinterpret_cast<const double2 *>(vals + s0[u] + k));
        const double2 w1 = __ldg(reinterpret_cast<const double2 *>(vals + s0[u] + len[u] + k));
        a0[u] += w0.x * xx.x + w0.y * xx.y;
        a1[u] += w1.x * xx.x + w1.y * xx.y;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
      for (int o = LPN / 2; o > 0; o >>= 1) {
        a0[u] += __shfl_xor_sync(0xffffffffu, a0[u], o);
        a1[u] += __shfl_xor_sync(0xffffffffu, a1[u], o);
      }
      const int64_t node = node0 + u * stride;
      if (node < n_nodes && lane == 0) {
        *reinterpret_cast<double2 *>(y + 2 * node) = make_double2(a0[u], a1[u]);
        if (DOT) {
          const double2 xs = *reinterpret_cast<const double2 *>(x + 2 * node);
          dot += a0[u] * xs.x + a1[u] * xs.y;
        }
      }
    }
  }
  if (DOT) {
    const double loc[1] = {dot};
    publish_and_reduce<1>(loc, partials, st, 0, red);
  }
}

#include "spmv_stream.cuh"
#include "pcg_persist.cuh"

constexpr int kBlock2 = -2;  // pseudo "lanes per row" selecting k_spmv_b2

static int spmv_lpr(int32_t n_rows, int64_t nnz, int block_dim) {
  if (block_dim == 2 && n_rows % 2 == 0) return kBlock2;
  const double avg = n_rows > 0 ? (double)nnz / n_rows : 1.0;
  if (avg <= 6) return 4;
  if (avg <= 20) return 8;
  if (avg <= 48) return 16;
  return 32;
}

static int reducing_grid(fe_ctx *ctx, int64_t work_items, int items_per_cta) {
  int64_t g = (work_items + items_per_cta - 1) / items_per_cta;
  const int64_t cap = (int64_t)ctx->num_sms * 16 < kMaxPartials ? (int64_t)ctx->num_sms * 16 : kMaxPartials;
  if (g > cap) g = cap;
  return (int)(g < 1 ? 1 : g);
}

template <bool DOT>
static int launch_spmv(fe_ctx *ctx, cudaStream_t s, int lpr, int32_t n_rows, const int32_t *rowptr,
                       const int32_t *colidx, const double *vals, const double *x, double *y, double *partials,
                       PcgState *st, const StreamPlan *sp = nullptr, const HaloPlan *fused_halo = nullptr) {
  if (sp && sp->on && sp->bs3) {  // persistent TMA-streamed kernel, 3x3 blocks (3 DOF per node)
    if (fused_halo)
      k_spmv_stream3<DOT, true><<<sp->grid, kStreamThreads, sp->smem, s>>>(
          n_rows / 3, sp->cap, sp->bptr, sp->bidx, vals, x, y, partials, st, ctx->p2p_dev, (HaloDev *)ctx->p2p_halo.ptr,
          fused_halo->send_idx);
    else
      k_spmv_stream3<DOT, false><<<sp->grid, kStreamThreads, sp->smem, s>>>(n_rows / 3, sp->cap, sp->bptr, sp->bidx, vals, x,
                                                                           y, partials, st, nullptr, nullptr, nullptr);
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }
  if (sp && sp->on && sp->scalar) {  // persistent TMA-streamed kernel, scalar CSR (1 DOF per node)
#define FE_STREAM1(P)                                                                                                   \
  do {                                                                                                                  \
    if (fused_halo)                                                                                                     \
      k_spmv_stream1<DOT, true, P><<<sp->grid, kStreamThreads, sp->smem, s>>>(                                          \
          n_rows, sp->cap, rowptr, colidx, vals, x, y, partials, st, ctx->p2p_dev, (HaloDev *)ctx->p2p_halo.ptr,        \
          fused_halo->send_idx);                                                                                        \
    else                                                                                                                \
      k_spmv_stream1<DOT, false, P><<<sp->grid, kStreamThreads, sp->smem, s>>>(n_rows, sp->cap, rowptr, colidx, vals, x, \
                                                                               y, partials, st, nullptr, nullptr, nullptr); \
  } while (0)
    if (sp->passes == 4)
      FE_STREAM1(4);
    else if (sp->passes == 2)
      FE_STREAM1(2);
    else
      FE_STREAM1(1);
#undef FE_STREAM1
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }
  if (sp && sp->on) {  // persistent TMA-streamed kernel (PCG)
    if (fused_halo)    // peer-memory transport: the exchange of x's interface values rides in the kernel
      k_spmv_stream<DOT, true><<<sp->grid, kStreamThreads, sp->smem, s>>>(
          n_rows / 2, sp->T, sp->cap, sp->bptr, sp->bidx, vals, x, y, partials, st, ctx->p2p_dev,
          (HaloDev *)ctx->p2p_halo.ptr, fused_halo->send_idx);
    else
      k_spmv_stream<DOT, false><<<sp->grid, kStreamThreads, sp->smem, s>>>(
          n_rows / 2, sp->T, sp->cap, sp->bptr, sp->bidx, vals, x, y, partials, st, nullptr, nullptr, nullptr);
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }
  if (lpr == kBlock2) {  // node-blocked fast path
    const int grid = reducing_grid(ctx, n_rows / 2, kRedBlock / 8);
    k_spmv_b2<DOT><<<grid, kRedBlock, 0, s>>>(n_rows / 2, rowptr, colidx, vals, x, y, partials, st);
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }
  const int grid = reducing_grid(ctx, n_rows, kRedBlock / lpr);
  switch (lpr) {
    case 4: k_spmv<4, DOT><<<grid, kRedBlock, 0, s>>>(n_rows, rowptr, colidx, vals, x, y, partials, st); break;
    case 8: k_spmv<8, DOT><<<grid, kRedBlock, 0, s>>>(n_rows, rowptr, colidx, vals, x, y, partials, st); break;
    case 16: k_spmv<16, DOT><<<grid, kRedBlock, 0, s>>>(n_rows, rowptr, colidx, vals, x, y, partials, st); break;
    default: k_spmv<32, DOT><<<grid, kRedBlock, 0, s>>>(n_rows, rowptr, colidx, vals, x, y, partials, st); break;
  }
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

// ---------------------------------------------------------------------------------------
// PCG vector kernels
// ---------------------------------------------------------------------------------------
// 8 lanes per row look for the diagonal entry (coalesced column reads)
__global__ void __launch_bounds__(256) k_extract_dinv(int32_t n_rows, const int32_t *__restrict__ rowptr,
                                                     const int32_t *__restrict__ colidx,
                                                     const double *__restrict__ vals, double *__restrict__ dinv,
                                                     PcgState *__restrict__ st) {
  const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;
  const int lane = threadIdx.x & 7;
  if (row >= n_rows) return;
  double d = 0.0;
  for (int32_t j = rowptr[row] + lane; j < rowptr[row + 1]; j += 8)
    if (colidx[j] == row) d = vals[j];
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) d += __shfl_xor_sync(0xffu << ((threadIdx.x & 31) & ~7), d, o);  // one lane holds it
  if (lane == 0) {
    if (!(d > 0.0)) st->breakdown = 1;  // not SPD (e.g. no Dirichlet condition at all)
    dinv[row] = 1.0 / d;
  }
}

// r = b - q (q = A x0); p = D^-1 r; sums: rz_new, rr, bnorm2
__global__ void __launch_bounds__(kRedBlock) k_pcg_init(int32_t n, const double *__restrict__ b,
                                                       const double *__restrict__ q, const double *__restrict__ dinv,
                                                       double *__restrict__ r, double *__restrict__ p,
                                                       double *__restrict__ partials, PcgState *__restrict__ st) {
  __shared__ double red[kRedBlock / 32];
  double loc[3] = {0.0, 0.0, 0.0};
  for (int64_t i = (int64_t)blockIdx.x * kRedBlock + threadIdx.x; i < n; i += (int64_t)gridDim.x * kRedBlock) {
    const double bi = b[i];
    const double ri = bi - q[i];
    const double zi = dinv[i] * ri;
    r[i] = ri;
    p[i] = zi;
    loc[0] += ri * zi;
    loc[1] += ri * ri;
    loc[2] += bi * bi;
  }
  publish_and_reduce<3>(loc, partials, st, 1, red);
}

// after the (all-reduced) init sums: rz[0] = rz_new, convergence of the initial guess
__global__ void k_pcg_init_finish(PcgState *st) {
  double t[3];
  reduced_sums<3>(st, 1, t);  // (r.z, r.r, b.b)
  if (threadIdx.x == 0) {
    st->sums[1] = t[0];
    st->sums[2] = t[1];
    st->sums[3] = t[2];
    st->rz[0] = t[0];
    if (!(t[2] > 0.0) || t[1] <= st->tol2 * t[2]) st->converged = 1;
  }
}

__global__ void __launch_bounds__(kRedBlock) k_pcg_update(int32_t n, int parity, const double *__restrict__ p,
                                                         const double *__restrict__ q,
                                                         const double *__restrict__ dinv, double *__restrict__ x,
                                                         double *__restrict__ r, double *__restrict__ partials,
                                                         PcgState *__restrict__ st) {
  __shared__ double red[kRedBlock / 32];
  if (st->converged | st->breakdown) return;
  double pq_[1];
  reduced_sums<1>(st, 0, pq_);
  const double pq = pq_[0];
  if (!(pq > 0.0) || !isfinite(pq)) {  // uniform across the grid
    if (blockIdx.x == 0 && threadIdx.x == 0) st->breakdown = 1;
    return;
  }
  const double alpha = st->rz[parity] / pq;
  double loc[2] = {0.0, 0.0};
  // A CTA walks chunks of U x 256 consecutive elements; a thread's U elements sit 256 apart, so all
  // 5 U loads use one base address per array plus immediate offsets and are issued before the first
  // dependent FMA.  Accumulation order is fixed by (chunk, u): deterministic.
  constexpr int U = 4;
  const int64_t n_chunks = (n + U * kRedBlock - 1) / (U * kRedBlock);
#pragma unroll 1
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t base = chunk * (U * kRedBlock) + threadIdx.x;
    if (base + (U - 1) * kRedBlock < n) {
      double rv[U], qv[U], xv[U], pv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rv[u] = r[base + u * kRedBlock];
        qv[u] = __ldg(q + base + u * kRedBlock);
        xv[u] = x[base + u * kRedBlock];
        pv[u] = __ldg(p + base + u * kRedBlock);
        dv[u] = __ldg(dinv + base + u * kRedBlock);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const double ri = rv[u] - alpha * qv[u];
        x[base + u * kRedBlock] = xv[u] + alpha * pv[u];
        r[base + u * kRedBlock] = ri;
        loc[0] += ri * (dv[u] * ri);
        loc[1] += ri * ri;
      }
    } else {
      for (int64_t i = base; i < n; i += kRedBlock) {
        const double ri = r[i] - alpha * q[i];
        x[i] += alpha * p[i];
        r[i] = ri;
        loc[0] += ri * (dinv[i] * ri);
        loc[1] += ri * ri;
      }
    }
  }
  publish_and_reduce<2>(loc, partials, st, 1, red);
}

__global__ void __launch_bounds__(256) k_pcg_pupdate(int32_t n, int parity, const double *__restrict__ r,
                                                    const double *__restrict__ dinv, double *__restrict__ p,
                                                    PcgState *__restrict__ st) {
  if (st->converged | st->breakdown) return;
  double t2[2];
  reduced_sums<2>(st, 1, t2);  // (r.z, r.r)
  const double rz_new = t2[0];
  const double beta = rz_new / st->rz[parity];
  constexpr int U = 4;  // chunks of U x 256 elements as in k_pcg_update: 3 U loads in flight per thread
  const int64_t n_chunks = (n + U * 256 - 1) / (U * 256);
#pragma unroll 1
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t base = chunk * (U * 256) + threadIdx.x;
    if (base + (U - 1) * 256 < n) {
      double rv[U], pv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        rv[u] = __ldg(r + base + u * 256);
        pv[u] = p[base + u * 256];
        dv[u] = __ldg(dinv + base + u * 256);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) p[base + u * 256] = dv[u] * rv[u] + beta * pv[u];
    } else {
      for (int64_t i = base; i < n; i += 256) p[i] = dinv[i] * r[i] + beta * p[i];
    }
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    st->rz[parity ^ 1] = rz_new;
    st->iters += 1;
    if (t2[1] <= st->tol2 * st->sums[3]) st->converged = 1;
  }
}

// NOTE on the flag race in k_pcg_pupdate: CTA 0 sets `converged` while other CTAs of the
// same launch may still be at their entry test.  A CTA that sees the flag early skips its
// part of the p update, which is harmless: once converged no later kernel reads p.

__global__ void k_pcg_state_init(PcgState *st, double tol2, P2PDev *pp) {
  st->pp = pp;
  for (int i = 0; i < 8; ++i) st->sums[i] = 0.0;
  st->rz[0] = st->rz[1] = 0.0;
  st->tol2 = tol2;
  st->iters = 0;
  st->converged = 0;
  st->breakdown = 0;
  st->ticket = 0;
}

__global__ void k_copy(int64_t n, const double *__restrict__ a, double *__restrict__ b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    b[i] = a[i];
}

// ---------------------------------------------------------------------------------------
// Dirichlet elimination in place
// ---------------------------------------------------------------------------------------
// Only rows next to a condition are touched: O(n_bc * valence) work instead of a sweep over the whole
// CSR (1.0 ms at 16 M triangles for 4 098 conditions, 2.2 x the assembly it follows).  Uses the
// structural symmetry of a finite-element pattern on the owned block: the rows that hold column c are
// the columns of row c.  `map` (int32 per column, all zero between calls) gives the condition index + 1.
__global__ void k_bc_scatter(int32_t n_bc, int32_t n_cols, const int32_t *__restrict__ dof, int32_t *__restrict__ map,
                             int *__restrict__ bad) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bc) return;
  const int32_t d = dof[i];
  if (d < 0 || d >= n_cols) {
    *bad = 1;
    return;
  }
  map[d] = i + 1;
}
__global__ void k_bc_unscatter(int32_t n_bc, int32_t n_cols, const int32_t *__restrict__ dof, int32_t *__restrict__ map) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_bc) return;
  const int32_t d = dof[i];
  if (d >= 0 && d < n_cols) map[d] = 0;
}

// One 8-lane group per condition on an OWNED dof c.  (a) row c becomes the unit row, rhs[c] = g_c;
// (b) every other row r of c's column list is eliminated by the group of its SMALLEST owned condition
// column (one writer per row, fixed summation order): rhs[r] -= sum_j K[r, c_j] g_j over all its
// condition columns (owned or ghost), those entries zeroed.
__global__ void __launch_bounds__(256) k_bc_rows(int32_t n_bc, int32_t n_rows, const int32_t *__restrict__ dof,
                                                const double *__restrict__ val, const int32_t *__restrict__ rowptr,
                                                const int32_t *__restrict__ colidx, double *__restrict__ vals,
                                                double *__restrict__ rhs, const int32_t *__restrict__ map) {
  constexpr int LPR = 8;
  const int32_t i = (blockIdx.x * 256 + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  const unsigned gmask = 0xffu << ((threadIdx.x & 31) / LPR * LPR);  // the group's lanes within the warp
  if (i >= n_bc) return;
  const int32_t c = dof[i];
  if (c >= n_rows) return;  // a ghost column: k_bc_ghost_rows
  const int32_t s = rowptr[c], e = rowptr[c + 1];
  for (int32_t j = s + lane; j < e; j += LPR) vals[j] = (colidx[j] == c) ? 1.0 : 0.0;
  if (lane == 0) rhs[c] = val[i];
  for (int32_t jr = s; jr < e; ++jr) {
    const int32_t r = colidx[jr];
    if (r == c || r >= n_rows || map[r] != 0) continue;  // uniform over the group
    const int32_t rs = rowptr[r], re = rowptr[r + 1];
    int32_t first = 0x7fffffff;  // smallest owned condition column of row r
    for (int32_t j = rs + lane; j < re; j += LPR) {
      const int32_t cc = colidx[j];
      if (cc < n_rows && map[cc] != 0 && cc < first) first = cc;
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) first = min(first, __shfl_xor_sync(gmask, first, o));
    if (first != c) continue;
    double acc = 0.0;
    for (int32_t j = rs + lane; j < re; j += LPR) {
      const int32_t m = map[colidx[j]];
      if (m != 0) {
        acc += vals[j] * val[m - 1];
        vals[j] = 0.0;
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(gmask, acc, o);
    if (lane == 0) rhs[r] -= acc;
  }
}

// Rows whose ONLY condition columns are ghost columns (multi-GPU: the condition node belongs to a
// neighbour rank, so this rank has no row for it).  Columns are sorted, ghosts (>= n_rows) sit at the end
// of a row: one thread per row looks at the last entry and leaves unless the row has a ghost tail.
__global__ void __launch_bounds__(256) k_bc_ghost_rows(int32_t n_rows, const int32_t *__restrict__ rowptr,
                                                      const int32_t *__restrict__ colidx, double *__restrict__ vals,
                                                      double *__restrict__ rhs, const double *__restrict__ val,
                                                      const int32_t *__restrict__ map) {
  const int32_t r = blockIdx.x * 256 + threadIdx.x;
  if (r >= n_rows) return;
  const int32_t s = rowptr[r], e = rowptr[r + 1];
  if (e == s || colidx[e - 1] < n_rows || map[r] != 0) return;
  bool any = false;
  for (int32_t j = e - 1; j >= s && colidx[j] >= n_rows; --j) any |= map[colidx[j]] != 0;
  if (!any) return;
  for (int32_t j = s; j < e && colidx[j] < n_rows; ++j)
    if (map[colidx[j]] != 0) return;  // has an owned condition column: k_bc_rows eliminates the whole row
  double acc = 0.0;
  for (int32_t j = s; j < e; ++j) {
    const int32_t m = map[colidx[j]];
    if (m != 0) {
      acc += vals[j] * val[m - 1];
      vals[j] = 0.0;
    }
  }
  rhs[r] -= acc;
}

// The full sweep the restricted kernels replace (kept for arbitrary, structurally unsymmetric CSR input:
// FE_B200_BC_SWEEP=1, and as the cross-check of the tests).
template <int LPR>
__global__ void __launch_bounds__(256) k_bc_apply(int32_t n_rows, const int32_t *__restrict__ rowptr,
                                                 const int32_t *__restrict__ colidx, double *__restrict__ vals,
                                                 double *__restrict__ rhs, const double *__restrict__ val,
                                                 const int32_t *__restrict__ map) {
  const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) / LPR;
  const int lane = threadIdx.x % LPR;
  double acc = 0.0;
  int32_t row_bc = 0;
  if (row < n_rows) {
    row_bc = map[row];
    const int32_t s = rowptr[row], e = rowptr[row + 1];
    for (int32_t j = s + lane; j < e; j += LPR) {
      const int32_t c = colidx[j];
      if (row_bc) {
        vals[j] = (c == row) ? 1.0 : 0.0;
      } else if (const int32_t m = map[c]) {
        acc += vals[j] * val[m - 1];
        vals[j] = 0.0;
      }
    }
  }
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (row < n_rows && lane == 0) rhs[row] = row_bc ? val[row_bc - 1] : (rhs[row] - acc);
}

__global__ void k_scatter_add(int32_t n, const int32_t *__restrict__ dof, const double *__restrict__ val,
                              double *__restrict__ rhs) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) rhs[dof[i]] += val[i];
}

// ---------------------------------------------------------------------------------------
// host driver shared by fe_pcg / fe_pcg_fixed / fe_dist_pcg
// ---------------------------------------------------------------------------------------
__global__ void k_pcg_clear_converged(PcgState *st) { st->converged = 0; }

struct PcgLaunch {
  fe_ctx *ctx;
  cudaStream_t s;
  int32_t n_rows;
  const int32_t *rowptr, *colidx;
  const double *vals, *b;
  double *x, *r, *q, *dinv, *p, *partials;
  PcgState *st;
  const HaloPlan *halo;
  bool dist, p2p = false;
  int lpr, vgrid;
  StreamPlan sp;

  // r = b - A x, p = D^-1 r, sums[1..3] = (r.z, r.r, b.b); converged flag from the TRUE residual
  // the streamed SpMV of the peer-memory transport carries the halo exchange itself
  bool fused() const { return dist && p2p && sp.on && halo->n_nbr > 0; }
  int exchange(double *vec) {
    if (!dist || fused()) return FE_OK;
    return p2p ? halo_exchange_p2p(ctx, s, halo, vec, n_rows) : halo_exchange(ctx, s, halo, vec, n_rows);
  }

  int true_residual_start() {
    int rc;
    k_copy<<<reducing_grid(ctx, n_rows, 1024), 256, 0, s>>>(n_rows, x, p);
    FE_LAUNCH_CHECK(ctx);
    if ((rc = exchange(p))) return rc;
    if ((rc = launch_spmv<false>(ctx, s, lpr, n_rows, rowptr, colidx, vals, p, q, partials, st, &sp,
                                 fused() ? halo : nullptr)))
      return rc;
    k_pcg_init<<<vgrid, kRedBlock, 0, s>>>(n_rows, b, q, dinv, r, p, partials, st);
    FE_LAUNCH_CHECK(ctx);
    if (dist && !p2p && (rc = allreduce_sum(ctx, s, st->sums + 1, 3))) return rc;
    k_pcg_init_finish<<<1, 32, 0, s>>>(st);
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }

  int iteration(int parity) {
    int rc;
    if ((rc = exchange(p))) return rc;
    if ((rc = launch_spmv<true>(ctx, s, lpr, n_rows, rowptr, colidx, vals, p, q, partials, st, &sp,
                                fused() ? halo : nullptr)))
      return rc;
    if (dist && !p2p && (rc = allreduce_sum(ctx, s, st->sums + 0, 1))) return rc;
    k_pcg_update<<<vgrid, kRedBlock, 0, s>>>(n_rows, parity, p, q, dinv, x, r, partials, st);
    FE_LAUNCH_CHECK(ctx);
    if (dist && !p2p && (rc = allreduce_sum(ctx, s, st->sums + 1, 2))) return rc;
    k_pcg_pupdate<<<vgrid, 256, 0, s>>>(n_rows, parity, r, dinv, p, st);
    FE_LAUNCH_CHECK(ctx);
    return FE_OK;
  }
};

// One CUDA graph holding `len` (even) iterations; single-GPU only.  Cached in the ctx as long
// as the same buffers are used (bench / time-stepping loops call the solver repeatedly).
static int get_chunk_graph(PcgLaunch &L, int len, cudaGraphExec_t *out) {
  fe_ctx *ctx = L.ctx;
  // Every argument the captured launches bake in: a second mesh with the same row count whose
  // buffers land on recycled addresses must not hit a graph holding the old ring capacity, grid,
  // block pattern or halo list.
  const void *key[16] = {L.rowptr, L.colidx, L.vals, L.x, L.r, L.st, (void *)(intptr_t)L.n_rows,
                         (void *)(intptr_t)(len * 256 + (L.lpr & 31) + (L.sp.on ? 32 : 0) + (L.p2p ? 64 : 0) + (L.sp.scalar ? 128 : 0)),
                         (void *)(intptr_t)L.sp.cap, (void *)(intptr_t)L.sp.smem, (void *)(intptr_t)L.sp.grid,
                         (void *)L.sp.bptr, (void *)L.sp.bidx, (void *)(L.halo ? L.halo->send_idx : nullptr),
                         (void *)(intptr_t)ctx->bp_built_token, (void *)(intptr_t)(L.vgrid * 16 + L.sp.passes + (L.sp.bs3 ? 8 : 0))};
  if (ctx->pcg_graph && memcmp(key, ctx->pcg_graph_key, sizeof(key)) == 0) {
    *out = (cudaGraphExec_t)ctx->pcg_graph;
    return FE_OK;
  }
  if (ctx->pcg_graph) {
    cudaGraphExecDestroy((cudaGraphExec_t)ctx->pcg_graph);
    ctx->pcg_graph = nullptr;
  }
  const int64_t launches_before = ctx->launches;
  FE_CUDA(cudaStreamBeginCapture(L.s, cudaStreamCaptureModeRelaxed));
  int rc = FE_OK;
  for (int k = 0; k < len && rc == FE_OK; ++k) rc = L.iteration(k & 1);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(L.s, &graph);
  ctx->launches = launches_before;  // captured, not launched
  if (rc != FE_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(FE_ERR_CUDA, "cudaStreamEndCapture failed: %s", cudaGetErrorString(e));
  cudaGraphExec_t exec = nullptr;
  e = cudaGraphInstantiate(&exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(FE_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(e));
  ctx->pcg_graph = exec;
  memcpy(ctx->pcg_graph_key, key, sizeof(key));
  *out = exec;
  return FE_OK;
}

// ---------------------------------------------------------------------------------------
// persistent-kernel driver (pcg_persist.cuh): one cooperative launch runs init + iterations; the host
// only verifies the true residual once the recurrence reports convergence and relaunches (restart
// from x) if it has drifted -- the same policy as the three-kernel path below.
// ---------------------------------------------------------------------------------------
static int persist_grid(fe_ctx *ctx, const void *kernel, size_t smem, int n_tiles) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kStreamThreads, smem) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (per_sm > 2) per_sm = 2;
  int cap = per_sm * ctx->num_sms;
  if (const char *e = getenv("FE_B200_PERSIST_GRID")) {  // debugging: fewer, longer-running CTAs
    const int v = atoi(e);
    if (v >= 1 && v < cap) cap = v;
  }
  return n_tiles < cap ? n_tiles : cap;
}

static int persist_solve(PcgLaunch &L, double *work, int32_t n_cols, double rtol, int32_t maxit, bool fixed,
                         int32_t *iters_out, double *relres_out) {
  fe_ctx *ctx = L.ctx;
  cudaStream_t s = L.s;
  const int32_t n_nodes = L.n_rows / 2;
  const int n_tiles = (n_nodes + kStreamTile - 1) / kStreamTile;
  const bool multi_gpu = L.dist && L.p2p;
  const void *kernel = multi_gpu ? (const void *)k_pcg_persist<true> : (const void *)k_pcg_persist<false>;
  FE_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L.sp.smem));
  const int grid = persist_grid(ctx, kernel, L.sp.smem, n_tiles);
  if (grid < 1) return fail(FE_ERR_CUDA, "pcg: the persistent kernel does not fit an SM (%zu B of shared memory)", L.sp.smem);
  const size_t flag_bytes = ((size_t)(1 + 2 * ctx->num_sms) * sizeof(unsigned) + 255) / 256 * 256;
  const size_t need = flag_bytes + (size_t)(2 * ctx->num_sms + 2) * kPQ * sizeof(double);
  const size_t cap_before = ctx->scratch_p.bytes;
  int rc = ctx->scratch_p.reserve(need);
  if (rc) return rc;
  if (ctx->scratch_p.bytes != cap_before) FE_CUDA(cudaMemsetAsync(ctx->scratch_p.ptr, 0, ctx->scratch_p.bytes, s));
  PersistArgs a;
  memset(&a, 0, sizeof(a));
  a.n_nodes = n_nodes;
  a.cap = L.sp.cap;
  a.bptr = L.sp.bptr;
  a.bidx = L.sp.bidx;
  a.vals = L.vals;
  a.b = L.b;
  a.dinv = L.dinv;
  a.x = L.x;
  a.r = L.r;
  a.w = L.q;
  a.u = L.p;
  a.p = work + 3 * (int64_t)L.n_rows + n_cols;
  a.s = a.p + L.n_rows;
  a.flags = (unsigned *)ctx->scratch_p.ptr;
  a.partials = (double *)((char *)ctx->scratch_p.ptr + flag_bytes);
  a.st = L.st;
  const bool multi = L.dist && L.p2p && L.halo->n_nbr > 0;
  a.pp = (L.dist && L.p2p) ? ctx->p2p_dev : nullptr;
  a.hd = (HaloDev *)ctx->p2p_halo.ptr;
  a.send_idx = multi ? L.halo->send_idx : nullptr;
  PcgState *h = (PcgState *)ctx->pinned;
  auto run = [&](int32_t it_end) -> int {
    a.it_end = it_end;
    void *params[] = {&a};
    FE_CUDA(cudaLaunchCooperativeKernel(kernel, dim3(grid), dim3(kStreamThreads), params, L.sp.smem, s));
    ctx->launches++;
    FE_CUDA(cudaMemcpyAsync(h, L.st, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
    FE_CUDA(cudaStreamSynchronize(s));
    if (h->breakdown == 2) cudaMemsetAsync(ctx->scratch_p.ptr, 0, ctx->scratch_p.bytes, s);  // barrier flags are stale
    if (getenv("FE_B200_PERSIST_PROF") && h->iters > 0)
      fprintf(stderr, "[k_pcg_persist rank %d grid %d] %d its, us/it: spmv %.2f | barrier+reduce %.2f | update %.2f | barrier+halo %.2f | init %.1f us\n",
              ctx->rank, grid, h->iters, 1e-3 * h->prof[0] / h->iters, 1e-3 * h->prof[1] / h->iters, 1e-3 * h->prof[2] / h->iters,
              1e-3 * h->prof[3] / h->iters, 1e-3 * h->prof[4]);
    return FE_OK;
  };
  constexpr int kMaxRestarts = 12;
  int restarts = 0;
  bool stagnated = false;
  double prev_true_rr = -1.0;
  if ((rc = run(maxit))) return rc;
  while (!fixed && !h->breakdown && h->converged != 2) {
    // recurrence converged (1) or maxit reached (0): a launch that stops at its first reduction
    // recomputes r = b - A x and leaves the TRUE residual in sums[2]
    if ((rc = run(h->iters))) return rc;
    if (h->breakdown || h->converged == 2 || h->iters >= maxit) break;
    if (prev_true_rr >= 0.0 && h->sums[2] > 0.25 * prev_true_rr) stagnated = true;
    prev_true_rr = h->sums[2];
    if (stagnated || ++restarts > kMaxRestarts) break;
    if ((rc = run(maxit))) return rc;
  }
  if (iters_out) *iters_out = h->iters;
  const double rel = (h->sums[3] > 0.0) ? sqrt(h->sums[2] / h->sums[3]) : 0.0;
  if (relres_out) *relres_out = rel;
  if (h->breakdown)
    return fail(FE_ERR_BREAKDOWN, "pcg: breakdown after %d iterations (matrix not SPD or singular: p.Ap = %g%s)", h->iters,
                h->sums[0], h->breakdown == 2 ? "; a grid / peer wait timed out"
                : ((h->sums[0] != h->sums[0] && a.pp) ? "; NaN on the peer-memory transport also means a rank did not deliver "
                                                        "within the spin-wait limit" : ""));
  if (!fixed && h->converged != 2) {
    const double slack = 100.0 * rtol > 1e-8 ? 100.0 * rtol : 1e-8;
    if (!(stagnated && rel <= slack))
      return fail(FE_ERR_NOT_CONVERGED, "pcg: not converged after %d iterations (relres %.3e > %.3e%s)", h->iters, rel, rtol,
                  stagnated ? "; restarts stopped reducing the true residual" : "");
  }
  return FE_OK;
}

int pcg_drive(fe_ctx *ctx, cudaStream_t s, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
              const int32_t *colidx, const double *vals, const double *b, double *x, double *work,
              const HaloPlan *halo, int block_dim, double rtol, int32_t maxit, bool fixed, int32_t *iters_out,
              double *relres_out) {
  FE_REQUIRE(ctx && rowptr && colidx && vals && b && x && work, "pcg: NULL argument");
  FE_REQUIRE(n_rows >= 0 && n_cols >= n_rows, "pcg: bad sizes %d x %d", n_rows, n_cols);
  FE_REQUIRE(maxit >= 0, "pcg: negative iteration count");
  FE_REQUIRE(block_dim >= 1 && block_dim <= 3, "pcg: block_dim must be 1, 2 or 3");
  FE_CUDA(cudaSetDevice(ctx->device));
  // The legacy / per-thread default streams cannot be captured into a CUDA graph: run the
  // solve on the ctx's own stream, ordered after the caller's stream.  The driver ends with a
  // host synchronisation, so work the caller submits afterwards is ordered behind the solve.
  if (s == nullptr || s == cudaStreamLegacy || s == cudaStreamPerThread) {
    FE_CUDA(cudaEventRecord((cudaEvent_t)ctx->work_event, s));
    FE_CUDA(cudaStreamWaitEvent((cudaStream_t)ctx->work_stream, (cudaEvent_t)ctx->work_event, 0));
    s = (cudaStream_t)ctx->work_stream;
  }
  int rc = ctx->scratch_b.reserve(sizeof(PcgState) + 256 + (size_t)kMaxPartials * 3 * sizeof(double));
  if (rc) return rc;
  static_assert(sizeof(PcgState) <= 256, "PcgState too large");

  PcgLaunch L;
  L.ctx = ctx;
  L.s = s;
  L.n_rows = n_rows;
  L.rowptr = rowptr;
  L.colidx = colidx;
  L.vals = vals;
  L.b = b;
  L.x = x;
  // workspace: r | q | dinv | p (p has the ghost tail)
  L.r = work;
  L.q = work + n_rows;
  L.dinv = work + 2 * (int64_t)n_rows;
  L.p = work + 3 * (int64_t)n_rows;
  L.st = (PcgState *)ctx->scratch_b.ptr;
  L.partials = (double *)((char *)ctx->scratch_b.ptr + 256);
  L.halo = halo;
  L.dist = halo != nullptr && ctx->nranks > 1;
  // NCCL-free transport: peer-memory halo stores + all-reduce fused into the kernels
  const bool p2p = L.dist && ctx->p2p_dev != nullptr && halo->peer_dst_off != nullptr &&
                   getenv("FE_B200_NO_P2P") == nullptr;
  L.p2p = p2p;
  PcgState *st = L.st;

  int32_t h_rowptr_end = 0;
  FE_CUDA(cudaMemcpyAsync(&h_rowptr_end, rowptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  k_pcg_state_init<<<1, 1, 0, s>>>(st, fixed ? -1.0 : rtol * rtol, p2p ? ctx->p2p_dev : nullptr);
  FE_LAUNCH_CHECK(ctx);
  FE_CUDA(cudaStreamSynchronize(s));
  L.lpr = spmv_lpr(n_rows, h_rowptr_end, block_dim);
  L.vgrid = reducing_grid(ctx, n_rows, kRedBlock * 4);
  unsigned char *diag_slot = nullptr;  // per-node position of the diagonal block (block pattern path)
  if (L.lpr == kBlock2 && n_rows > 0 && getenv("FE_B200_NO_STREAM") == nullptr) {
    // node-level pattern for the TMA-streamed SpMV (rebuilt per solve: one pass over colidx)
    const int32_t n_nodes = n_rows / 2;
    const int64_t nnzb = h_rowptr_end / 4;
    const size_t bptr_bytes = ((size_t)(n_nodes + 1 + 128) * 4 + 255) / 256 * 256;  // +128: tile slice over-read
    const size_t bidx_bytes = ((size_t)(nnzb + 8) * 4 + 255) / 256 * 256;
    if ((rc = ctx->scratch_c.reserve(bptr_bytes + bidx_bytes + 256 + (size_t)n_nodes))) return rc;
    L.sp.bptr = (int32_t *)ctx->scratch_c.ptr;
    L.sp.bidx = (int32_t *)((char *)ctx->scratch_c.ptr + bptr_bytes);
    int *max_deg = (int *)((char *)ctx->scratch_c.ptr + bptr_bytes + bidx_bytes);
    diag_slot = (unsigned char *)ctx->scratch_c.ptr + bptr_bytes + bidx_bytes + 256;
    const bool cached = ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx &&
                        ctx->bp_built_token == ctx->bp_token && ctx->bp_built_rows == n_rows;
    int h_max_deg = ctx->bp_max_deg;
    if (!cached) {
      ctx->bp_built_token = 0;
      FE_CUDA(cudaMemsetAsync(max_deg, 0, sizeof(int), s));
      k_block_pattern<<<grid_for(((int64_t)n_nodes + 1) * 8, 256), 256, 0, s>>>(n_nodes, rowptr, colidx, L.sp.bptr,
                                                                                L.sp.bidx, max_deg, diag_slot);
      FE_LAUNCH_CHECK(ctx);
      FE_CUDA(cudaMemcpyAsync(&h_max_deg, max_deg, sizeof(int), cudaMemcpyDeviceToHost, s));
      FE_CUDA(cudaStreamSynchronize(s));
      if (ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx) {
        ctx->bp_built_token = ctx->bp_token;
        ctx->bp_built_rows = n_rows;
        ctx->bp_max_deg = h_max_deg;
      }
    }
    const int T = kStreamTile;  // fixed tile; meshes whose valence makes it too large use k_spmv_b2
    const int cap = (T * (h_max_deg > 0 ? h_max_deg : 1) + 3) & ~3;
    const size_t smem = stream_smem_bytes(T, cap);
    if (smem <= 110 * 1024) {
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      L.sp.on = true;
      L.sp.T = T;
      L.sp.cap = cap;
      L.sp.smem = smem;
      const int n_tiles = (n_nodes + T - 1) / T;
      L.sp.grid = n_tiles < 2 * ctx->num_sms ? n_tiles : 2 * ctx->num_sms;
    }
  }

  if (block_dim == 3 && n_rows > 0 && n_rows % 3 == 0 && h_rowptr_end % 9 == 0 && getenv("FE_B200_NO_STREAM") == nullptr &&
      getenv("FE_B200_NO_BLOCK3") == nullptr && ((uintptr_t)vals & 15) == 0) {
    // 3 DOF per node: node-level pattern (4 index bytes per 3x3 block) for k_spmv_stream3
    const int32_t n_nodes = n_rows / 3;
    const int64_t nnzb = h_rowptr_end / 9;
    const size_t bptr_bytes = ((size_t)(n_nodes + 1 + 128) * 4 + 255) / 256 * 256;  // +128: tile slice over-read
    const size_t bidx_bytes = ((size_t)(nnzb + 8) * 4 + 255) / 256 * 256;
    if ((rc = ctx->scratch_c.reserve(bptr_bytes + bidx_bytes + 256))) return rc;
    L.sp.bptr = (int32_t *)ctx->scratch_c.ptr;
    L.sp.bidx = (int32_t *)((char *)ctx->scratch_c.ptr + bptr_bytes);
    int *flags3 = (int *)((char *)ctx->scratch_c.ptr + bptr_bytes + bidx_bytes);  // [0] tile max, [1] bad
    const bool cached = ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx &&
                        ctx->bp_built_token == ctx->bp_token && ctx->bp_built_rows == -3 - n_rows;
    int h3[2] = {ctx->bp_max_deg, 0};
    if (!cached) {
      ctx->bp_built_token = 0;
      FE_CUDA(cudaMemsetAsync(flags3, 0, 2 * sizeof(int), s));
      k_block_pattern3<<<grid_for(((int64_t)n_nodes + 1) * 16, 256), 256, 0, s>>>(n_nodes, rowptr, colidx, L.sp.bptr, L.sp.bidx,
                                                                                 flags3 + 1);
      FE_LAUNCH_CHECK(ctx);
      const int n_tiles = (n_nodes + kBlock3Tile - 1) / kBlock3Tile;
      k_tile_max_entries<<<grid_for(n_tiles, 256), 256, 0, s>>>(n_nodes, kBlock3Tile, L.sp.bptr, flags3);
      FE_LAUNCH_CHECK(ctx);
      FE_CUDA(cudaMemcpyAsync(h3, flags3, 2 * sizeof(int), cudaMemcpyDeviceToHost, s));
      FE_CUDA(cudaStreamSynchronize(s));
      if (!h3[1] && ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx) {
        ctx->bp_built_token = ctx->bp_token;
        ctx->bp_built_rows = -3 - n_rows;  // (3-DOF entry: distinct from the scalar (-2 - n) and the 2-DOF (n) ones)
        ctx->bp_max_deg = h3[0];
      }
    }
    const int cap = ((h3[0] > 0 ? h3[0] : 1) + 3) & ~3;
    const size_t smem = stream3_smem_bytes(cap);
    if (!h3[1] && smem <= 110 * 1024) {  // (a CSR that is not 3x3-blocked falls through to the scalar kernel)
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream3<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream3<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream3<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      FE_CUDA(cudaFuncSetAttribute(k_spmv_stream3<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      L.sp.on = true;
      L.sp.bs3 = true;
      L.sp.cap = cap;
      L.sp.smem = smem;
      const int n_tiles = (n_nodes + kBlock3Tile - 1) / kBlock3Tile;
      L.sp.grid = n_tiles < 2 * ctx->num_sms ? n_tiles : 2 * ctx->num_sms;
    }
  }

  if (!L.sp.on && L.lpr != kBlock2 && n_rows > 0 && getenv("FE_B200_NO_STREAM") == nullptr &&
      ((uintptr_t)vals & 15) == 0 && ((uintptr_t)colidx & 15) == 0 && ((uintptr_t)rowptr & 15) == 0) {
    // scalar CSR: the streamed kernel works on rowptr / colidx as they are; only the largest tile
    // (in entries) is needed to size the ring
    if ((rc = ctx->scratch_c.reserve(256))) return rc;
    int *tile_max = (int *)ctx->scratch_c.ptr;
    // rows per tile from the mean row length: 240 rows of 7 entries, 120 of <= 24, 60 of 45 (tetrahedra)
    const double avg_row = (double)h_rowptr_end / n_rows;
    const int passes = avg_row <= 12.0 ? 4 : (avg_row <= 24.0 ? 2 : 1);
    const int kStream1Tile = stream1_tile(passes);
    const bool cached = ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx &&
                        ctx->bp_built_token == ctx->bp_token && ctx->bp_built_rows == -2 - n_rows;
    int h_tile_max = ctx->bp_max_deg;
    if (!cached) {
      ctx->bp_built_token = 0;
      const int n_tiles = (n_rows + kStream1Tile - 1) / kStream1Tile;
      FE_CUDA(cudaMemsetAsync(tile_max, 0, sizeof(int), s));
      k_tile_max_entries<<<grid_for(n_tiles, 256), 256, 0, s>>>(n_rows, kStream1Tile, rowptr, tile_max);
      FE_LAUNCH_CHECK(ctx);
      FE_CUDA(cudaMemcpyAsync(&h_tile_max, tile_max, sizeof(int), cudaMemcpyDeviceToHost, s));
      FE_CUDA(cudaStreamSynchronize(s));
      if (ctx->bp_token != 0 && ctx->bp_rowptr == rowptr && ctx->bp_colidx == colidx) {
        ctx->bp_built_token = ctx->bp_token;
        ctx->bp_built_rows = -2 - n_rows;  // (negative: scalar-path entry, never equal to a block-path one)
        ctx->bp_max_deg = h_tile_max;
      }
    }
    const int cap = ((h_tile_max > 0 ? h_tile_max : 1) + 3) & ~3;
    const size_t smem = stream1_smem_bytes(cap, passes);
    if (smem <= 110 * 1024) {
#define FE_STREAM1_ATTR(P)                                                                                                   \
  do {                                                                                                                       \
    FE_CUDA(cudaFuncSetAttribute(k_spmv_stream1<true, false, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    FE_CUDA(cudaFuncSetAttribute(k_spmv_stream1<false, false, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));  \
    FE_CUDA(cudaFuncSetAttribute(k_spmv_stream1<true, true, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));    \
    FE_CUDA(cudaFuncSetAttribute(k_spmv_stream1<false, true, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
  } while (0)
      if (passes == 4)
        FE_STREAM1_ATTR(4);
      else if (passes == 2)
        FE_STREAM1_ATTR(2);
      else
        FE_STREAM1_ATTR(1);
#undef FE_STREAM1_ATTR
      L.sp.on = true;
      L.sp.scalar = true;
      L.sp.passes = passes;
      L.sp.cap = cap;
      L.sp.smem = smem;
      const int n_tiles = (n_rows + kStream1Tile - 1) / kStream1Tile;
      L.sp.grid = n_tiles < 2 * ctx->num_sms ? n_tiles : 2 * ctx->num_sms;
    }
  }

  if (n_rows > 0 && diag_slot) {
    k_dinv_b2<<<grid_for(n_rows / 2, 256), 256, 0, s>>>(n_rows / 2, L.sp.bptr, diag_slot, vals, L.dinv, &st->breakdown);
    FE_LAUNCH_CHECK(ctx);
  } else if (n_rows > 0) {
    k_extract_dinv<<<grid_for((int64_t)n_rows * 8, 256), 256, 0, s>>>(n_rows, rowptr, colidx, vals, L.dinv, st);
    FE_LAUNCH_CHECK(ctx);
  }
  // One persistent cooperative kernel for the whole solve (pcg_persist.cuh: 2 DOF per node, streamed
  // pattern, one GPU or the peer-memory transport).  Measured (S16M, profiles/r02_*): it trades 16 n
  // bytes of extra vector traffic per iteration (the s = A p recurrence of the single-reduction CG)
  // for one cross-GPU round trip and three kernel boundaries -- a loss on 1-2 GPUs (0.682 vs 0.641,
  // 0.363 vs 0.337 ms/iteration), a gain once the per-rank share is small (8 GPUs: 0.099 vs 0.103).
  // Default: 4 ranks and more; FE_B200_PERSIST=1 / 0 forces it on / off.
  const bool aligned16 = (((uintptr_t)x | (uintptr_t)b | (uintptr_t)work | (uintptr_t)vals) & 15) == 0;
  bool want_persist = L.dist && ctx->nranks >= 4;
  if (const char *e = getenv("FE_B200_PERSIST")) want_persist = atoi(e) != 0;
  if (getenv("FE_B200_NO_PERSIST")) want_persist = false;
  if (want_persist && L.sp.on && !L.sp.scalar && !L.sp.bs3 && aligned16 && (!L.dist || L.p2p))
    return persist_solve(L, work, n_cols, rtol, maxit, fixed, iters_out, relres_out);

  if ((rc = L.true_residual_start())) return rc;

  PcgState *h = (PcgState *)ctx->pinned;
  constexpr int kChunk = 50;   // iterations between convergence polls (even: graph parity)
  constexpr int kMaxRestarts = 12;
  // (the peer-memory transport keeps its sequence numbers on the device, so it can be captured too)
  const bool use_graph = (!L.dist || L.p2p) && getenv("FE_B200_NO_GRAPH") == nullptr;
  int it = 0, local = 0, restarts = 0;
  bool done = (maxit == 0), stagnated = false;
  double prev_true_rr = -1.0;
  FE_CUDA(cudaMemcpyAsync(h, st, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
  FE_CUDA(cudaStreamSynchronize(s));
  if (h->converged || h->breakdown) done = true;
  while (!done) {
    // ---- run until the recurrence says converged (or maxit / breakdown)
    while (it < maxit) {
      const int chunk = (maxit - it < kChunk) ? (maxit - it) : kChunk;
      if (use_graph && chunk == kChunk && (local & 1) == 0) {
        cudaGraphExec_t exec;
        if ((rc = get_chunk_graph(L, kChunk, &exec))) return rc;
        FE_CUDA(cudaGraphLaunch(exec, s));
        ctx->launches += (3 + ((L.p2p && !L.fused()) ? 1 : 0)) * kChunk;
        it += kChunk;
        local += kChunk;
      } else {
        for (int k = 0; k < chunk; ++k, ++it, ++local)
          if ((rc = L.iteration(local & 1))) return rc;
      }
      if (fixed && it < maxit) continue;  // no polling in throughput mode
      FE_CUDA(cudaMemcpyAsync(h, st, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
      FE_CUDA(cudaStreamSynchronize(s));
      if (h->converged || h->breakdown) break;
    }
    if (fixed || h->breakdown || !h->converged) break;
    // ---- the recurrence residual drifts from b - A x over tens of thousands of iterations:
    //      recompute the true residual and, if it is not below the tolerance, restart from x.
    k_pcg_clear_converged<<<1, 1, 0, s>>>(st);
    FE_LAUNCH_CHECK(ctx);
    if ((rc = L.true_residual_start())) return rc;
    local = 0;
    FE_CUDA(cudaMemcpyAsync(h, st, sizeof(PcgState), cudaMemcpyDeviceToHost, s));
    FE_CUDA(cudaStreamSynchronize(s));
    if (h->converged || it >= maxit) {
      done = true;
    } else {
      // attainable accuracy: a restart that no longer reduces the true residual cannot help
      if (prev_true_rr >= 0.0 && h->sums[2] > 0.25 * prev_true_rr) stagnated = true;
      prev_true_rr = h->sums[2];
      if (stagnated || ++restarts > kMaxRestarts) done = true;
    }
  }
  if (iters_out) *iters_out = h->iters;
  if (relres_out) *relres_out = (h->sums[3] > 0.0) ? sqrt(h->sums[2] / h->sums[3]) : 0.0;
  if (h->breakdown)
    return fail(FE_ERR_BREAKDOWN, "pcg: breakdown after %d iterations (matrix not SPD or singular: p.Ap = %g%s)",
                h->iters, h->sums[0],
                (h->sums[0] != h->sums[0] && L.p2p) ? "; NaN on the peer-memory transport also means a rank did not "
                                                      "deliver within the spin-wait limit" : "");
  if (!fixed && !h->converged) {
    // Attainable accuracy: restarts that no longer reduce the true residual are accepted only within
    // a documented slack of the tolerance (fe_b200.h); anything worse is reported as not converged.
    const double rel = (h->sums[3] > 0.0) ? sqrt(h->sums[2] / h->sums[3]) : 0.0;
    const double slack = 100.0 * rtol > 1e-8 ? 100.0 * rtol : 1e-8;  // 1e-8: the residual north_star asks for
    if (!(stagnated && rel <= slack))
      return fail(FE_ERR_NOT_CONVERGED, "pcg: not converged after %d iterations (relres %.3e > %.3e%s)", h->iters, rel,
                  rtol, stagnated ? "; restarts stopped reducing the true residual" : "");
  }
  return FE_OK;
}

}  // namespace fe

using namespace fe;

extern "C" {

int fe_dirichlet_apply(fe_ctx *ctx, void *stream, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
                       const int32_t *colidx, double *vals, double *rhs, int32_t n_bc, const int32_t *bc_dof,
                       const double *bc_val) {
  FE_REQUIRE(ctx && rowptr && colidx && vals && rhs, "fe_dirichlet_apply: NULL argument");
  FE_REQUIRE(n_rows >= 0 && n_cols >= n_rows && n_bc >= 0, "fe_dirichlet_apply: bad sizes");
  if (n_bc == 0 || n_rows == 0) return FE_OK;
  FE_REQUIRE(bc_dof && bc_val, "fe_dirichlet_apply: NULL condition arrays");
  cudaStream_t s = as_stream(stream);
  // column -> condition map, all zero between calls (zeroed when (re)allocated, un-scattered at the end)
  const size_t map_bytes = ((size_t)n_cols * sizeof(int32_t) + 255) / 256 * 256;
  const size_t cap_before = ctx->bc_map.bytes;  // (a re-allocation may return the old address: compare capacities)
  int rc = ctx->bc_map.reserve(map_bytes + 256);
  if (rc) return rc;
  if (ctx->bc_map.bytes != cap_before) FE_CUDA(cudaMemsetAsync(ctx->bc_map.ptr, 0, ctx->bc_map.bytes, s));
  int32_t *map = (int32_t *)ctx->bc_map.ptr;
  int *bad = (int *)((char *)ctx->bc_map.ptr + map_bytes);
  k_bc_scatter<<<grid_for(n_bc, 256), 256, 0, s>>>(n_bc, n_cols, bc_dof, map, bad);
  FE_LAUNCH_CHECK(ctx);
  int hbad = 0;
  FE_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s));
  FE_CUDA(cudaStreamSynchronize(s));
  if (hbad) {
    cudaMemsetAsync(ctx->bc_map.ptr, 0, ctx->bc_map.bytes, s);
    return fail(FE_ERR_ARG, "fe_dirichlet_apply: a condition DOF is outside [0, %d)", n_cols);
  }
  if (getenv("FE_B200_BC_SWEEP")) {
    k_bc_apply<8><<<grid_for((int64_t)n_rows * 8, 256), 256, 0, s>>>(n_rows, rowptr, colidx, vals, rhs, bc_val, map);
    FE_LAUNCH_CHECK(ctx);
  } else {
    k_bc_rows<<<grid_for((int64_t)n_bc * 8, 256), 256, 0, s>>>(n_bc, n_rows, bc_dof, bc_val, rowptr, colidx, vals, rhs, map);
    FE_LAUNCH_CHECK(ctx);
    if (n_cols > n_rows) {  // (a rank of a partition: conditions on ghost nodes reach rows of this rank)
      k_bc_ghost_rows<<<grid_for(n_rows, 256), 256, 0, s>>>(n_rows, rowptr, colidx, vals, rhs, bc_val, map);
      FE_LAUNCH_CHECK(ctx);
    }
  }
  k_bc_unscatter<<<grid_for(n_bc, 256), 256, 0, s>>>(n_bc, n_cols, bc_dof, map);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_scatter_add(fe_ctx *ctx, void *stream, int32_t n, const int32_t *dof, const double *val, double *rhs) {
  FE_REQUIRE(ctx && (n == 0 || (dof && val && rhs)), "fe_scatter_add: NULL argument");
  if (n <= 0) return FE_OK;
  k_scatter_add<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(n, dof, val, rhs);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_spmv(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr, const int32_t *colidx,
            const double *vals, const double *x, double *y, int32_t block_dim) {
  FE_REQUIRE(ctx && rowptr && colidx && vals && x && y, "fe_spmv: NULL argument");
  if (n_rows <= 0) return FE_OK;
  cudaStream_t s = as_stream(stream);
  int32_t nnz = 0;
  FE_CUDA(cudaMemcpyAsync(&nnz, rowptr + n_rows, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  FE_CUDA(cudaStreamSynchronize(s));
  return launch_spmv<false>(ctx, s, spmv_lpr(n_rows, nnz, block_dim), n_rows, rowptr, colidx, vals, x, y, nullptr, nullptr);
}

int fe_pcg_cache_pattern(fe_ctx *ctx, const int32_t *rowptr, const int32_t *colidx, int64_t token) {
  FE_REQUIRE(ctx, "fe_pcg_cache_pattern: NULL ctx");
  ctx->bp_rowptr = rowptr;
  ctx->bp_colidx = colidx;
  ctx->bp_token = (rowptr && colidx) ? token : 0;
  return FE_OK;
}

// r | w (= q) | dinv | u (= p of the three-kernel path; ghost tail) | p | s  -- the last two only for the
// single-reduction recurrence of the persistent kernel
int64_t fe_pcg_work_len(int32_t n_rows, int32_t n_cols) { return 5 * (int64_t)n_rows + (int64_t)n_cols; }

int fe_pcg(fe_ctx *ctx, void *stream, int32_t n, const int32_t *rowptr, const int32_t *colidx, const double *vals,
           const double *b, double *x, double *work, int32_t block_dim, double rtol, int32_t maxit, int32_t *iters,
           double *relres) {
  return pcg_drive(ctx, as_stream(stream), n, n, rowptr, colidx, vals, b, x, work, nullptr, block_dim, rtol, maxit,
                   false, iters, relres);
}

int fe_pcg_fixed(fe_ctx *ctx, void *stream, int32_t n, const int32_t *rowptr, const int32_t *colidx,
                 const double *vals, const double *b, double *x, double *work, int32_t block_dim, int32_t iters) {
  int32_t done = 0;
  double rel = 0.0;
  return pcg_drive(ctx, as_stream(stream), n, n, rowptr, colidx, vals, b, x, work, nullptr, block_dim, 0.0, iters,
                   true, &done, &rel);
}

}  // extern "C"
