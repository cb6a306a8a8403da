// Symbolic phase (once per mesh): CSR pattern + corner->slot map.
//
// Replaces analysis.py:714-735 (get_row_col_indices, 36 Python dict look-ups per element)
// and the pattern half of scipy's COO->CSR at analysis.py:661.  Instead of sorting the
// (3 dim)^2 * E scalar triplets (604 M keys at 16 M triangles) the pattern is built on the
// node graph:
//   1. counting sort of the 3E (element, vertex) corners keyed by node  (histogram + scan +
//      scatter = one radix pass with radix N);
//   2. per node, in one thread: order its corners by element id, gather the <= 3*valence
//      candidate neighbours, sort + unique them -> sorted node adjacency (incl. self);
//   3. scan of the row lengths; expansion of every node pair to a dim x dim block gives the
//      scalar CSR with sorted columns and explicit zeros kept = scipy's canonical form.
// The same pass records, for every corner, where its three blocks land in the node's row
// (k0,k1,k2) and whether it is the first contributor, which is all fe_assemble needs.
#include "plan.cuh"

namespace fe {

constexpr int kMaxDegree = 255;  // positions are stored in 8 bits

struct PlanFlags {
  int bad_node;    // connectivity index outside [0, n_nodes)
  int max_degree;  // max node valence incl. self
  int too_dense;   // valence > kMaxDegree
};

__global__ void k_count_corners(int64_t n3, const int32_t *__restrict__ conn, int32_t n_nodes, int32_t n_owned,
                                int32_t *__restrict__ cnt, PlanFlags *__restrict__ flags) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const int32_t nd = conn[i];
  if (nd < 0 || nd >= n_nodes) {
    flags->bad_node = 1;
    return;
  }
  if (nd < n_owned) atomicAdd(cnt + nd, 1);
}

__global__ void k_fill_corners(int64_t n3, const int32_t *__restrict__ conn, int32_t n_nodes, int32_t n_owned,
                               const int32_t *__restrict__ corner_ptr, int32_t *__restrict__ cursor,
                               int32_t *__restrict__ corner_tmp) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const int32_t nd = conn[i];
  if (nd < 0 || nd >= n_owned) return;
  const int32_t pos = corner_ptr[nd] + atomicAdd(cursor + nd, 1);
  const int64_t e = i / 3;
  corner_tmp[pos] = (int32_t)((e << 2) | (i - 3 * e));
}

__device__ __forceinline__ void insertion_sort(int32_t *a, int n) {
  for (int i = 1; i < n; ++i) {
    const int32_t key = a[i];
    int j = i - 1;
    while (j >= 0 && a[j] > key) {
      a[j + 1] = a[j];
      --j;
    }
    a[j + 1] = key;
  }
}

// One thread per owned node: canonical corner order, candidate neighbours, sort + unique.
__global__ void __launch_bounds__(128) k_node_adjacency(int32_t n_owned, const int32_t *__restrict__ conn,
                                                       const int32_t *__restrict__ corner_ptr,
                                                       int32_t *__restrict__ corner_tmp, int32_t *__restrict__ cand,
                                                       int32_t *__restrict__ ndeg, PlanFlags *__restrict__ flags) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
  int32_t *cs = corner_tmp + c0;
  const int nc = c1 - c0;
  insertion_sort(cs, nc);  // the atomic scatter left them in arbitrary order
  int32_t *cd = cand + 3 * (int64_t)c0;
  for (int k = 0; k < nc; ++k) {
    const int64_t e = cs[k] >> 2;
    cd[3 * k + 0] = conn[3 * e + 0];
    cd[3 * k + 1] = conn[3 * e + 1];
    cd[3 * k + 2] = conn[3 * e + 2];
  }
  insertion_sort(cd, 3 * nc);
  int m = 0;
  for (int k = 0; k < 3 * nc; ++k)
    if (m == 0 || cd[k] != cd[m - 1]) cd[m++] = cd[k];
  ndeg[n] = m;
  atomicMax(&flags->max_degree, m);
  if (m > kMaxDegree) flags->too_dense = 1;
}

// One thread per owned node: publish the adjacency row and the corner records.
__global__ void __launch_bounds__(128) k_node_records(int32_t n_owned, const int32_t *__restrict__ conn,
                                                     const int32_t *__restrict__ corner_ptr,
                                                     const int32_t *__restrict__ corner_tmp,
                                                     const int32_t *__restrict__ cand,
                                                     const int32_t *__restrict__ adj_ptr, int32_t *__restrict__ adj,
                                                     int2 *__restrict__ corner_rec) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
  const int32_t a0 = adj_ptr[n];
  const int deg = adj_ptr[n + 1] - a0;
  if (deg > kMaxDegree) return;
  const int32_t *cd = cand + 3 * (int64_t)c0;
  int32_t *row = adj + a0;
  for (int k = 0; k < deg; ++k) row[k] = cd[k];
  uint32_t seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int32_t c = c0; c < c1; ++c) {
    const int32_t ev = corner_tmp[c];
    const int64_t e = ev >> 2;
    uint32_t packed = 0;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int32_t target = conn[3 * e + j];
      int lo = 0, hi = deg - 1;  // binary search in the sorted row
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cd[mid] < target)
          lo = mid + 1;
        else
          hi = mid;
      }
      packed |= (uint32_t)lo << (8 * j);
      const uint32_t bit = 1u << (lo & 31);
      if (!(seen[lo >> 5] & bit)) {
        seen[lo >> 5] |= bit;
        packed |= 1u << (24 + j);
      }
    }
    corner_rec[c] = make_int2(ev, (int)packed);
  }
}

__global__ void k_conn4(int64_t n_elems, const int32_t *__restrict__ conn, const int32_t *__restrict__ mat_id,
                        int4 *__restrict__ conn4) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  conn4[e] = make_int4(conn[3 * e], conn[3 * e + 1], conn[3 * e + 2], mat_id ? mat_id[e] : 0);
}

// Scalar CSR export: row (n, d) = columns {m*dim + c : m in adj(n), c < dim}.
__global__ void __launch_bounds__(128) k_export_csr(int32_t n_owned, int32_t dim,
                                                   const int32_t *__restrict__ adj_ptr,
                                                   const int32_t *__restrict__ adj, int32_t *__restrict__ rowptr,
                                                   int32_t *__restrict__ colidx) {
  // one warp per node: lanes stride over the dim*deg columns of each of the dim rows
  const int32_t n = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (n >= n_owned) return;
  const int32_t a0 = adj_ptr[n];
  const int deg = adj_ptr[n + 1] - a0;
  const int64_t base = (int64_t)a0 * dim * dim;
  const int rowlen = deg * dim;
  for (int d = 0; d < dim; ++d) {
    if (lane == 0) rowptr[n * dim + d] = (int32_t)(base + (int64_t)d * rowlen);
    for (int q = lane; q < rowlen; q += 32) colidx[base + (int64_t)d * rowlen + q] = adj[a0 + q / dim] * dim + q % dim;
  }
  if (n == n_owned - 1 && lane == 0) rowptr[n_owned * dim] = (int32_t)(base + (int64_t)dim * rowlen);
}

template <typename T>
static int dev_alloc(T **p, int64_t count, int64_t *bytes_acc) {
  size_t b = (size_t)(count > 0 ? count : 1) * sizeof(T);
  cudaError_t e = cudaMalloc((void **)p, b);
  if (e != cudaSuccess) return fail(FE_ERR_CUDA, "cudaMalloc(%zu) failed: %s", b, cudaGetErrorString(e));
  if (bytes_acc) *bytes_acc += (int64_t)b;
  return FE_OK;
}

}  // namespace fe

using namespace fe;

extern "C" {

int fe_plan_destroy(fe_plan *p) {
  if (!p) return FE_OK;
  cudaSetDevice(p->ctx->device);
  cudaFree(p->corner_ptr);
  cudaFree(p->corner_rec);
  cudaFree(p->adj_ptr);
  cudaFree(p->adj);
  cudaFree(p->conn4);
  delete p;
  return FE_OK;
}

int fe_plan_create(fe_ctx *ctx, void *stream, int32_t n_nodes, int32_t n_owned, int64_t n_elems, int32_t dim,
                   const int32_t *conn, const int32_t *mat_id, fe_plan **out) {
  FE_REQUIRE(ctx && out, "fe_plan_create: NULL ctx/out");
  FE_REQUIRE(dim == 1 || dim == 2, "fe_plan_create: dim must be 1 (magnetic) or 2 (elasticity), got %d", dim);
  FE_REQUIRE(n_nodes >= 0 && n_owned >= 0 && n_owned <= n_nodes, "fe_plan_create: bad node counts %d/%d", n_owned,
             n_nodes);
  FE_REQUIRE(n_elems >= 0 && (n_elems == 0 || conn), "fe_plan_create: bad connectivity");
  if (n_elems >= (int64_t(1) << 29)) return fail(FE_ERR_UNSUPPORTED, "fe_plan_create: more than 2^29 elements");
  if ((int64_t)n_nodes * dim >= (int64_t(1) << 31))
    return fail(FE_ERR_UNSUPPORTED, "fe_plan_create: DOF count overflows int32");
  cudaStream_t st = as_stream(stream);
  FE_CUDA(cudaSetDevice(ctx->device));

  fe_plan *p = new fe_plan();
  p->ctx = ctx;
  p->n_nodes = n_nodes;
  p->n_owned = n_owned;
  p->dim = dim;
  p->n_elems = n_elems;
  int rc = FE_OK;
  int32_t *cursor = nullptr, *corner_tmp = nullptr, *cand = nullptr, *ndeg = nullptr;
  PlanFlags *flags = nullptr;
  int64_t *totals = nullptr;
  const int64_t n3 = 3 * n_elems;
  PlanFlags hflags = {0, 0, 0};
  int64_t htot[2] = {0, 0};

#define PLAN_TRY(expr)            \
  do {                            \
    rc = (expr);                  \
    if (rc != FE_OK) goto done;   \
  } while (0)
#define PLAN_CUDA(call)                                                                              \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      rc = fail(FE_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      goto done;                                                                                     \
    }                                                                                                \
  } while (0)
#define PLAN_LAUNCHED()                                                                              \
  do {                                                                                               \
    ctx->launches++;                                                                                 \
    cudaError_t e__ = cudaGetLastError();                                                            \
    if (e__ != cudaSuccess) {                                                                        \
      rc = fail(FE_ERR_CUDA, "launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
      goto done;                                                                                     \
    }                                                                                                \
  } while (0)

  PLAN_TRY(dev_alloc(&p->corner_ptr, (int64_t)n_owned + 1, &p->bytes));
  PLAN_TRY(dev_alloc(&p->adj_ptr, (int64_t)n_owned + 1, &p->bytes));
  PLAN_TRY(dev_alloc(&p->conn4, n_elems, &p->bytes));
  PLAN_TRY(dev_alloc(&cursor, (int64_t)n_owned + 1, nullptr));
  PLAN_TRY(dev_alloc(&ndeg, (int64_t)n_owned + 1, nullptr));
  PLAN_TRY(dev_alloc(&flags, 1, nullptr));
  PLAN_TRY(dev_alloc(&totals, 2, nullptr));
  PLAN_CUDA(cudaMemsetAsync(cursor, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  PLAN_CUDA(cudaMemsetAsync(ndeg, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  PLAN_CUDA(cudaMemsetAsync(flags, 0, sizeof(PlanFlags), st));
  PLAN_CUDA(cudaMemsetAsync(totals, 0, 2 * sizeof(int64_t), st));

  if (n_elems > 0 && n_owned > 0) {
    // 1. histogram of corners per owned node (cursor doubles as the count array)
    k_count_corners<<<grid_for(n3, 256), 256, 0, st>>>(n3, conn, n_nodes, n_owned, cursor, flags);
    PLAN_LAUNCHED();
    k_conn4<<<grid_for(n_elems, 256), 256, 0, st>>>(n_elems, conn, mat_id, p->conn4);
    PLAN_LAUNCHED();
  }
  PLAN_TRY(exclusive_scan_i32(ctx, st, cursor, p->corner_ptr, n_owned, totals + 0));
  PLAN_CUDA(cudaMemcpyAsync(htot, totals, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PLAN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
  PLAN_CUDA(cudaStreamSynchronize(st));
  if (hflags.bad_node) {
    rc = fail(FE_ERR_ARG, "fe_plan_create: connectivity references a node outside [0, %d)", n_nodes);
    goto done;
  }
  p->n_corners = htot[0];
  PLAN_TRY(dev_alloc(&p->corner_rec, p->n_corners, &p->bytes));
  PLAN_TRY(dev_alloc(&corner_tmp, p->n_corners, nullptr));
  PLAN_TRY(dev_alloc(&cand, 3 * p->n_corners, nullptr));
  if (p->n_corners > 0) {
    // 2. scatter (counting sort), then per-node canonical order + sorted unique neighbours
    PLAN_CUDA(cudaMemsetAsync(cursor, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
    k_fill_corners<<<grid_for(n3, 256), 256, 0, st>>>(n3, conn, n_nodes, n_owned, p->corner_ptr, cursor, corner_tmp);
    PLAN_LAUNCHED();
    k_node_adjacency<<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, ndeg,
                                                            flags);
    PLAN_LAUNCHED();
  }
  // 3. row lengths -> block row pointer
  PLAN_TRY(exclusive_scan_i32(ctx, st, ndeg, p->adj_ptr, n_owned, totals + 1));
  PLAN_CUDA(cudaMemcpyAsync(htot, totals, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  PLAN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
  PLAN_CUDA(cudaStreamSynchronize(st));
  p->nnzb = htot[1];
  p->nnz = p->nnzb * dim * dim;
  p->max_degree = hflags.max_degree;
  if (hflags.too_dense) {
    rc = fail(FE_ERR_UNSUPPORTED, "fe_plan_create: a node has %d neighbours (limit %d)", hflags.max_degree,
              kMaxDegree);
    goto done;
  }
  if (p->nnz >= (int64_t(1) << 31)) {
    rc = fail(FE_ERR_UNSUPPORTED, "fe_plan_create: nnz = %lld overflows int32 CSR indices", (long long)p->nnz);
    goto done;
  }
  PLAN_TRY(dev_alloc(&p->adj, p->nnzb, &p->bytes));
  if (p->n_corners > 0) {
    k_node_records<<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, p->adj_ptr,
                                                          p->adj, p->corner_rec);
    PLAN_LAUNCHED();
  }
  PLAN_CUDA(cudaStreamSynchronize(st));

done:
  cudaFree(cursor);
  cudaFree(corner_tmp);
  cudaFree(cand);
  cudaFree(ndeg);
  cudaFree(flags);
  cudaFree(totals);
  if (rc != FE_OK) {
    fe_plan_destroy(p);
    return rc;
  }
  *out = p;
  return FE_OK;
#undef PLAN_TRY
#undef PLAN_CUDA
#undef PLAN_LAUNCHED
}

int64_t fe_plan_nnz(const fe_plan *p) { return p ? p->nnz : 0; }
int32_t fe_plan_n_rows(const fe_plan *p) { return p ? p->n_owned * p->dim : 0; }
int32_t fe_plan_max_degree(const fe_plan *p) { return p ? p->max_degree : 0; }
int64_t fe_plan_bytes(const fe_plan *p) { return p ? p->bytes : 0; }

int fe_plan_csr(const fe_plan *p, void *stream, int32_t *rowptr, int32_t *colidx) {
  FE_REQUIRE(p && rowptr && (colidx || p->nnz == 0), "fe_plan_csr: NULL argument");
  cudaStream_t st = as_stream(stream);
  if (p->n_owned == 0) {
    FE_CUDA(cudaMemsetAsync(rowptr, 0, sizeof(int32_t), st));
    return FE_OK;
  }
  k_export_csr<<<grid_for((int64_t)p->n_owned * 32, 128), 128, 0, st>>>(p->n_owned, p->dim, p->adj_ptr, p->adj,
                                                                       rowptr, colidx);
  FE_LAUNCH_CHECK(p->ctx);
  return FE_OK;
}

}  // extern "C"
