// Block-vector products for the modal path.
//
// The reference hands K and M (same sparsity pattern: elementary_mass_matrix has the layout of
// elementary_matrix, elements.py:513-536 / :466-511) to scipy.sparse.linalg.eigsh, whose ARPACK
// loop multiplies one vector at a time (analysis.py:779-782).  The LOBPCG driver in
// finite_elements_b200/modal.py works on blocks of m vectors instead, so the matrix is streamed
// from HBM once per block: Y_A = A X and, in the same pass over the shared pattern, Y_B = B X.
//
// Layout: X is double[n_cols][m] row-major (a torch (n, m) tensor), so the m values a matrix entry
// needs are contiguous.  G = 4/8/16/32 lanes own one row; lane l owns column l of the block.  Each
// step of the row loop reads one (col, a, b) triple -- the same address in all G lanes, i.e. one
// broadcast transaction -- and one coalesced m-wide row of X (L2/L1 resident: the rows of X
// touched by neighbouring matrix rows overlap almost completely).  No reduction across lanes is
// needed, and every output element is produced by one thread in a fixed order (deterministic).
// HBM bytes per call: (8 or 16) nnz + 4 nnz + 4 n + 8 m n_cols (X once) + 8 m n (per output).
#include "common.cuh"

namespace fe {

template <int G, bool PAIR>
__global__ void __launch_bounds__(256) k_spmm(int32_t n_rows, int32_t m, const int32_t *__restrict__ rowptr,
                                             const int32_t *__restrict__ colidx, const double *__restrict__ va,
                                             const double *__restrict__ vb, const double *__restrict__ X,
                                             double *__restrict__ YA, double *__restrict__ YB) {
  const int64_t row = ((int64_t)blockIdx.x * 256 + threadIdx.x) / G;
  const int lane = threadIdx.x % G;
  if (row >= n_rows) return;
  const int32_t s = __ldg(rowptr + row), e = __ldg(rowptr + row + 1);
  for (int c0 = 0; c0 < m; c0 += G) {  // blocks wider than G lanes: one more pass over the row
    const int col = c0 + lane;
    const bool on = col < m;
    const double *xc = X + (on ? col : 0);
    double accA = 0.0, accB = 0.0;
    int32_t j = s;
    for (; j + 4 <= e; j += 4) {  // four entries in flight
      int32_t c[4];
      double a[4], b[4], x[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        c[u] = __ldg(colidx + j + u);
        a[u] = __ldg(va + j + u);
        b[u] = PAIR ? __ldg(vb + j + u) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) x[u] = on ? __ldg(xc + (int64_t)c[u] * m) : 0.0;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        accA += a[u] * x[u];
        if (PAIR) accB += b[u] * x[u];
      }
    }
    for (; j < e; ++j) {
      const int32_t c = __ldg(colidx + j);
      const double x = on ? __ldg(xc + (int64_t)c * m) : 0.0;
      accA += __ldg(va + j) * x;
      if (PAIR) accB += __ldg(vb + j) * x;
    }
    if (on) {
      YA[row * m + col] = accA;
      if (PAIR) YB[row * m + col] = accB;
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

}  // namespace fe

using namespace fe;

extern "C" {

int fe_spmm_pair(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr, const int32_t *colidx,
                 const double *vals_a, const double *vals_b, const double *x, double *y_a, double *y_b, int32_t m) {
  FE_REQUIRE(ctx && rowptr && colidx && vals_a && x && y_a, "fe_spmm_pair: NULL argument");
  FE_REQUIRE((vals_b == nullptr) == (y_b == nullptr), "fe_spmm_pair: vals_b and y_b go together");
  FE_REQUIRE(m >= 1 && m <= 1024, "fe_spmm_pair: block width %d outside [1, 1024]", m);
  if (n_rows <= 0) return FE_OK;
  cudaStream_t s = as_stream(stream);
  const int g = m <= 4 ? 4 : (m <= 8 ? 8 : (m <= 16 ? 16 : 32));
  const int grid = grid_for((int64_t)n_rows * g, 256);
#define FE_SPMM(G)                                                                                        \
  do {                                                                                                    \
    if (vals_b)                                                                                           \
      k_spmm<G, true><<<grid, 256, 0, s>>>(n_rows, m, rowptr, colidx, vals_a, vals_b, x, y_a, y_b);        \
    else                                                                                                  \
      k_spmm<G, false><<<grid, 256, 0, s>>>(n_rows, m, rowptr, colidx, vals_a, nullptr, x, y_a, nullptr); \
  } while (0)
  switch (g) {
    case 4: FE_SPMM(4); break;
    case 8: FE_SPMM(8); break;
    case 16: FE_SPMM(16); break;
    default: FE_SPMM(32); break;
  }
#undef FE_SPMM
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
