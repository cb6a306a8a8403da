// Block-vector products for the modal path.
//
// The reference hands K and M (same sparsity pattern: elementary_mass_matrix has the layout of
// elementary_matrix, elements.py:513-536 / :466-511) to scipy.sparse.linalg.eigsh, whose ARPACK
// loop multiplies one vector at a time (analysis.py:779-782).  The LOBPCG driver in
// finite_elements_b200/modal.py works on blocks of m vectors instead, so the matrix is streamed
// from HBM once per block: Y_A = A X and, in the same pass over the shared pattern, Y_B = B X.
//
// Layout: X is double[n_cols][m] row-major (a torch (n, m) tensor), so the m values a matrix entry
// needs are contiguous.  G lanes own one row; a lane owns CPL = 4, 2 or 1 adjacent columns of the
// block (the widest that divides m), so per matrix entry it issues one broadcast load of the
// (col, a, b) triple and ONE 32/16/8-byte load of X for CPL FMAs -- the first version (one column per
// lane) was bound by load instructions, not by memory (ncu: 19 % DRAM, 68 % L1).  The rows of X
// touched by neighbouring matrix rows overlap almost completely (L1/L2 resident).  No reduction
// across lanes is needed, and every output element is produced by one thread in a fixed order
// (deterministic).
// HBM bytes per call: (8 or 16) nnz + 4 nnz + 4 n + 8 m n_cols (X once) + 8 m n (per output).
#include "common.cuh"

namespace fe {

// 256-bit global accesses (one sector per lane and instruction; two 128-bit halves count the sector twice in L1)
__device__ __forceinline__ void ld4_nc(const double *p, double *x) {
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p));
}
__device__ __forceinline__ void ld4(const double *p, double *x) {
  asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(x[0]), "=d"(x[1]), "=d"(x[2]), "=d"(x[3]) : "l"(p) : "memory");
}
__device__ __forceinline__ void st4(double *p, const double *x) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(x[0]), "d"(x[1]), "d"(x[2]), "d"(x[3]) : "memory");
}

// CPL columns per lane (1, 2 or 4: one 8-, 16- or 32-byte load of X per matrix entry), G lanes per
// row.  EPI selects what happens to y = A x in the epilogue:
//   0  Y_A = y (and Y_B = B x when PAIR)
//   1  one step of the Chebyshev iteration (modal.py: chebyshev_preconditioner), all row-local:
//        z += x_row;  r -= y;  x_out = c1 x_row + c2 dinv[row] r
//      (x_out is a different buffer: other rows are still gathering from x)
template <int CPL, int G, bool PAIR, int EPI>
__global__ void __launch_bounds__(256) k_spmm(int32_t n_rows, int32_t m, const int32_t *__restrict__ rowptr,
                                             const int32_t *__restrict__ colidx, const double *__restrict__ va,
                                             const double *__restrict__ vb, const double *__restrict__ X,
                                             double *__restrict__ YA, double *__restrict__ YB,
                                             const double *__restrict__ dinv, double *__restrict__ r,
                                             double *__restrict__ z, double c1, double c2, bool vec32) {
  const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  if (row >= n_rows) return;
  const int32_t s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (int c0 = 0; c0 < m; c0 += G * CPL) {  // blocks wider than G * CPL: one more pass over the row
    const int col = c0 + lane * CPL;
    const bool on = col < m;                 // m % CPL == 0: a lane is fully on or fully off
    const double *xc = X + (on ? col : 0);
    double accA[CPL], accB[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) accA[q] = accB[q] = 0.0;
    auto load_x = [&](int32_t c, double *x) {
      const double *src = xc + (int64_t)c * m;
      if (CPL == 4 && vec32) {
        ld4_nc(src, x);
      } else if (CPL == 4) {
        const double2 lo = __ldg(reinterpret_cast<const double2 *>(src));
        const double2 hi = __ldg(reinterpret_cast<const double2 *>(src) + 1);
        x[0] = lo.x, x[1] = lo.y, x[2 % CPL] = hi.x, x[3 % CPL] = hi.y;
      } else if (CPL == 2) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(src));
        x[0] = v.x, x[1 % CPL] = v.y;
      } else {
        x[0] = __ldg(src);
      }
    };
    int32_t j = s;
    if (on) {
      for (; j + 2 <= e; j += 2) {  // two entries in flight
        int32_t c[2];
        double a[2], b[2], x[2][CPL];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          c[u] = __ldg(colidx + j + u);
          a[u] = __ldg(va + j + u);
          b[u] = PAIR ? __ldg(vb + j + u) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) load_x(c[u], x[u]);
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
          for (int q = 0; q < CPL; ++q) {
            accA[q] += a[u] * x[u][q];
            if (PAIR) accB[q] += b[u] * x[u][q];
          }
      }
      for (; j < e; ++j) {
        double x[CPL];
        load_x(__ldg(colidx + j), x);
        const double a = __ldg(va + j), b = PAIR ? __ldg(vb + j) : 0.0;
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          accA[q] += a * x[q];
          if (PAIR) accB[q] += b * x[q];
        }
      }
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        const int64_t o = row * m + col + q;
        if (EPI == 0) {
          YA[o] = accA[q];
          if (PAIR) YB[o] = accB[q];
        } else {
          const double xr = __ldg(X + o);
          const double rr = r[o] - accA[q];
          z[o] += xr;
          r[o] = rr;
          YA[o] = c1 * xr + c2 * (__ldg(dinv + row) * rr);
        }
      }
    }
  }
}

// Node-blocked form for 2 DOF per node (block_dim == 2, the CSR of an fe_plan with dim == 2): G lanes own a NODE,
// i.e. rows 2i and 2i+1, which share one column list of (2m, 2m+1) pairs.  Per 2x2 block a lane loads the column
// index once, the four matrix values as two 16-byte broadcasts, and the two rows of X once for BOTH output rows --
// the row-per-group kernel above fetched every row of X twice (ncu r01: the fused Chebyshev step at 0.52 of its
// roofline, L1 wavefronts of the 96-byte row gathers).  Entries are added in the same order as above (k_r0 x_2m, then
// k_r1 x_2m+1, block after block): bit-identical results.
template <int CPL, int G, bool PAIR, int EPI>
__global__ void __launch_bounds__(256) k_spmm_b2(int32_t n_nodes, int32_t m, const int32_t *__restrict__ rowptr,
                                                const int32_t *__restrict__ colidx, const double *__restrict__ va,
                                                const double *__restrict__ vb, const double *__restrict__ X,
                                                double *__restrict__ YA, double *__restrict__ YB,
                                                const double *__restrict__ dinv, double *__restrict__ r,
                                                double *__restrict__ z, double c1, double c2, bool vec32) {
  const int64_t node = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  if (node >= n_nodes) return;
  const int32_t s0 = __ldg(rowptr + 2 * node), s1 = __ldg(rowptr + 2 * node + 1);
  const int nb = (s1 - s0) >> 1;  // 2x2 blocks in the node's rows
  for (int c0 = 0; c0 < m; c0 += G * CPL) {
    const int col = c0 + lane * CPL;
    if (col >= m) continue;  // m % CPL == 0: a lane is fully on or fully off
    const double *xc = X + col;
    double a0[CPL], a1[CPL], b0[CPL], b1[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) a0[q] = a1[q] = b0[q] = b1[q] = 0.0;
    auto load_x = [&](int64_t row, double *x) {
      const double *src = xc + row * m;
      if (CPL == 4 && vec32) {
        ld4_nc(src, x);
      } else if (CPL == 4) {
        const double2 lo = __ldg(reinterpret_cast<const double2 *>(src));
        const double2 hi = __ldg(reinterpret_cast<const double2 *>(src) + 1);
        x[0] = lo.x, x[1] = lo.y, x[2 % CPL] = hi.x, x[3 % CPL] = hi.y;
      } else if (CPL == 2) {
        const double2 v = __ldg(reinterpret_cast<const double2 *>(src));
        x[0] = v.x, x[1 % CPL] = v.y;
      } else {
        x[0] = __ldg(src);
      }
    };
    auto block = [&](int jb, int32_t c, const double *x0, const double *x1) {
      // (rows start on multiples of 4 entries and hold pairs: the value pairs are 16-byte aligned)
      const double2 k0 = __ldg(reinterpret_cast<const double2 *>(va + s0) + jb);
      const double2 k1 = __ldg(reinterpret_cast<const double2 *>(va + s1) + jb);
#pragma unroll
      for (int q = 0; q < CPL; ++q) {
        a0[q] = fma(k0.x, x0[q], a0[q]);
        a0[q] = fma(k0.y, x1[q], a0[q]);
        a1[q] = fma(k1.x, x0[q], a1[q]);
        a1[q] = fma(k1.y, x1[q], a1[q]);
      }
      if (PAIR) {
        const double2 m0 = __ldg(reinterpret_cast<const double2 *>(vb + s0) + jb);
        const double2 m1 = __ldg(reinterpret_cast<const double2 *>(vb + s1) + jb);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          b0[q] = fma(m0.x, x0[q], b0[q]);
          b0[q] = fma(m0.y, x1[q], b0[q]);
          b1[q] = fma(m1.x, x0[q], b1[q]);
          b1[q] = fma(m1.y, x1[q], b1[q]);
        }
      }
      (void)c;
    };
    int jb = 0;
    for (; jb + 2 <= nb; jb += 2) {  // two blocks in flight
      const int32_t ca = __ldg(colidx + s0 + 2 * jb), cb = __ldg(colidx + s0 + 2 * jb + 2);
      double xa0[CPL], xa1[CPL], xb0[CPL], xb1[CPL];
      load_x(ca, xa0);
      load_x((int64_t)ca + 1, xa1);
      load_x(cb, xb0);
      load_x((int64_t)cb + 1, xb1);
      block(jb, ca, xa0, xa1);
      block(jb + 1, cb, xb0, xb1);
    }
    if (jb < nb) {
      const int32_t ca = __ldg(colidx + s0 + 2 * jb);
      double xa0[CPL], xa1[CPL];
      load_x(ca, xa0);
      load_x((int64_t)ca + 1, xa1);
      block(jb, ca, xa0, xa1);
    }
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
      const int64_t row = 2 * node + rr;
      const int64_t o0 = row * m + col;
      const double *ya = rr ? a1 : a0;
      if (CPL == 4 && vec32) {  // whole 32-byte pieces: one access per array instead of four
        if (EPI == 0) {
          st4(YA + o0, ya);
          if (PAIR) st4(YB + o0, rr ? b1 : b0);
        } else {
          double xr[4], rv[4], zv[4], out[4];
          ld4_nc(X + o0, xr);
          ld4(r + o0, rv);
          ld4(z + o0, zv);
          const double di = __ldg(dinv + row);
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            rv[q] -= ya[q];
            zv[q] += xr[q];
            out[q] = c1 * xr[q] + c2 * (di * rv[q]);
          }
          st4(z + o0, zv);
          st4(r + o0, rv);
          st4(YA + o0, out);
        }
      } else {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
          const int64_t o = o0 + q;
          if (EPI == 0) {
            YA[o] = ya[q];
            if (PAIR) YB[o] = rr ? b1[q] : b0[q];
          } else {
            const double xr = __ldg(X + o);
            const double rn = r[o] - ya[q];
            z[o] += xr;
            r[o] = rn;
            YA[o] = c1 * xr + c2 * (__ldg(dinv + row) * rn);
          }
        }
      }
    }
  }
}

__global__ void __launch_bounds__(256) k_csr_diag(int32_t n_rows, const int32_t *__restrict__ rowptr,
                                                 const int32_t *__restrict__ colidx,
                                                 const double *__restrict__ vals, double *__restrict__ diag) {
  const int32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  double d = 0.0;
  for (int32_t j = rowptr[row]; j < rowptr[row + 1]; ++j)
    if (colidx[j] == row) d = vals[j];
  diag[row] = d;
}

static void spmm_shape(int m, int *cpl, int *g) {
  *cpl = (m % 4 == 0) ? 4 : ((m % 2 == 0) ? 2 : 1);
  const int lanes = (m + *cpl - 1) / *cpl;
  *g = lanes <= 1 ? 1 : (lanes <= 2 ? 2 : (lanes <= 4 ? 4 : (lanes <= 8 ? 8 : (lanes <= 16 ? 16 : 32))));
}

template <bool PAIR, int EPI>
static void spmm_launch(cudaStream_t s, int32_t n_rows, int32_t m, const int32_t *rowptr, const int32_t *colidx,
                        const double *va, const double *vb, const double *x, double *ya, double *yb,
                        const double *dinv, double *r, double *z, double c1, double c2, bool block2) {
  int cpl, g;
  spmm_shape(m, &cpl, &g);
  const int grid = grid_for((int64_t)(block2 ? n_rows / 2 : n_rows) * g, 256);
  // 256-bit accesses need 32-byte aligned blocks (row pitch 8 m bytes with m % 4 == 0 keeps every piece aligned)
  const bool vec32 = m % 4 == 0 && (((uintptr_t)x | (uintptr_t)ya | (uintptr_t)yb | (uintptr_t)r | (uintptr_t)z) & 31) == 0;
#define FE_SPMM(CPL, G)                                                                                                  \
  do {                                                                                                                   \
    if (block2)                                                                                                          \
      k_spmm_b2<CPL, G, PAIR, EPI><<<grid, 256, 0, s>>>(n_rows / 2, m, rowptr, colidx, va, vb, x, ya, yb, dinv, r, z, c1, c2, \
                                                        vec32);                                                          \
    else                                                                                                                 \
      k_spmm<CPL, G, PAIR, EPI><<<grid, 256, 0, s>>>(n_rows, m, rowptr, colidx, va, vb, x, ya, yb, dinv, r, z, c1, c2,    \
                                                     vec32);                                                             \
  } while (0)
#define FE_SPMM_G(CPL)                    \
  switch (g) {                            \
    case 1: FE_SPMM(CPL, 1); break;       \
    case 2: FE_SPMM(CPL, 2); break;       \
    case 4: FE_SPMM(CPL, 4); break;       \
    case 8: FE_SPMM(CPL, 8); break;       \
    case 16: FE_SPMM(CPL, 16); break;     \
    default: FE_SPMM(CPL, 32); break;     \
  }
  if (cpl == 4) {
    FE_SPMM_G(4)
  } else if (cpl == 2) {
    FE_SPMM_G(2)
  } else {
    FE_SPMM_G(1)
  }
#undef FE_SPMM_G
#undef FE_SPMM
}

}  // namespace fe

using namespace fe;

extern "C" {

int fe_spmm_pair(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr, const int32_t *colidx,
                 const double *vals_a, const double *vals_b, const double *x, double *y_a, double *y_b, int32_t m,
                 int32_t block_dim) {
  FE_REQUIRE(ctx && rowptr && colidx && vals_a && x && y_a, "fe_spmm_pair: NULL argument");
  FE_REQUIRE((vals_b == nullptr) == (y_b == nullptr), "fe_spmm_pair: vals_b and y_b go together");
  FE_REQUIRE(m >= 1 && m <= 1024, "fe_spmm_pair: block width %d outside [1, 1024]", m);
  FE_REQUIRE(((uintptr_t)x & 15) == 0, "fe_spmm_pair: x must be 16-byte aligned");
  if (n_rows <= 0) return FE_OK;
  cudaStream_t s = as_stream(stream);
  const bool b2 = block_dim == 2 && n_rows % 2 == 0 && ((uintptr_t)vals_a & 15) == 0 && ((uintptr_t)vals_b & 15) == 0;
  if (vals_b)
    spmm_launch<true, 0>(s, n_rows, m, rowptr, colidx, vals_a, vals_b, x, y_a, y_b, nullptr, nullptr, nullptr, 0, 0, b2);
  else
    spmm_launch<false, 0>(s, n_rows, m, rowptr, colidx, vals_a, nullptr, x, y_a, nullptr, nullptr, nullptr, nullptr, 0,
                          0, b2);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_cheb_step(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr, const int32_t *colidx,
                 const double *vals, const double *dinv, const double *d_in, double *d_out, double *r, double *z,
                 double c1, double c2, int32_t m, int32_t block_dim) {
  FE_REQUIRE(ctx && rowptr && colidx && vals && dinv && d_in && d_out && r && z, "fe_cheb_step: NULL argument");
  FE_REQUIRE(d_in != d_out, "fe_cheb_step: d_out must not alias d_in (rows are still gathering from it)");
  FE_REQUIRE(m >= 1 && m <= 1024, "fe_cheb_step: block width %d outside [1, 1024]", m);
  FE_REQUIRE(((uintptr_t)d_in & 15) == 0, "fe_cheb_step: d_in must be 16-byte aligned");
  if (n_rows <= 0) return FE_OK;
  spmm_launch<false, 1>(as_stream(stream), n_rows, m, rowptr, colidx, vals, nullptr, d_in, d_out, nullptr, dinv, r, z,
                        c1, c2, block_dim == 2 && n_rows % 2 == 0 && ((uintptr_t)vals & 15) == 0);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_csr_diagonal(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr, const int32_t *colidx,
                    const double *vals, double *diag) {
  FE_REQUIRE(ctx && rowptr && colidx && vals && diag, "fe_csr_diagonal: NULL argument");
  if (n_rows <= 0) return FE_OK;
  k_csr_diag<<<grid_for(n_rows, 256), 256, 0, as_stream(stream)>>>(n_rows, rowptr, colidx, vals, diag);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

}  // extern "C"
