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
  int fan_irregular;  // some node's corners do not form simple fans (edge shared by > 2 elements, ...)
  int max_mat;     // largest mat_id (negative ids are reported through bad_mat)
  int bad_mat;
};

// ---------------------------------------------------------------------------------------
// Fan ordering: walk the corners of a node so that consecutive corners share an edge.  Every
// off-diagonal block of the node's row then receives its (at most two) contributions from
// consecutive steps and can be finished in registers by the assembly kernel.
//   seed record : .x = first neighbour of a chain, .y = k | FAN_SEED << 8 | k_self << 13
//   step record : .x = the corner's NEW neighbour, .y = k | flags << 8 | mat_id << 13
// k = position of .x in the node's sorted adjacency row.  A chain is "closed" when the fan
// wraps around an interior node: its first block is held back (FAN_HOLD_A) and completed by
// the last step (FAN_ADD_FIRST).
// ---------------------------------------------------------------------------------------
constexpr int kFanMaxCorners = 32;
constexpr int kFanMatBits = 19;
enum : uint32_t { FAN_SEED = 1, FAN_MULTI = 2, FAN_HOLD_A = 4, FAN_LAST = 8, FAN_ADD_FIRST = 16 };

__device__ __forceinline__ int row_position(const int32_t *row, int deg, int32_t target) {
  int lo = 0, hi = deg - 1;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (row[mid] < target)
      lo = mid + 1;
    else
      hi = mid;
  }
  return lo;
}

// Returns the number of records of node `self`, or -1 when its corners are not simple fans.
template <bool EMIT>
__device__ int fan_walk(int32_t self, int nc, const int32_t *__restrict__ cs, const int32_t *__restrict__ conn,
                        const int32_t *__restrict__ mat_id, const int32_t *row, int deg, int2 *out) {
  if (nc > kFanMaxCorners) return -1;
  int32_t va[kFanMaxCorners], vb[kFanMaxCorners];
  for (int i = 0; i < nc; ++i) {
    const int64_t e = cs[i] >> 2;
    const int v = cs[i] & 3;
    va[i] = conn[3 * e + (v + 1) % 3];
    vb[i] = conn[3 * e + (v + 2) % 3];
    if (va[i] == self || vb[i] == self || va[i] == vb[i]) return -1;  // degenerate element
    if (mat_id && (uint32_t)mat_id[e] >= (1u << kFanMatBits)) return -1;
  }
  auto occurrences = [&](int32_t m) {
    int c = 0;
    for (int j = 0; j < nc; ++j) c += (va[j] == m) + (vb[j] == m);
    return c;
  };
  for (int i = 0; i < nc; ++i)
    if (occurrences(va[i]) > 2 || occurrences(vb[i]) > 2) return -1;
  const uint32_t full = (nc == 32) ? 0xffffffffu : ((1u << nc) - 1u);
  uint32_t used = 0;
  int nrec = 0, chains = 0;
  const int k_self = EMIT ? row_position(row, deg, self) : 0;
  while (used != full) {
    ++chains;
    int start = -1;
    int32_t sv = 0;
    for (int i = 0; i < nc && start < 0; ++i) {
      if (used & (1u << i)) continue;
      if (occurrences(va[i]) == 1) {
        start = i;
        sv = va[i];
      } else if (occurrences(vb[i]) == 1) {
        start = i;
        sv = vb[i];
      }
    }
    bool closed = false;
    if (start < 0) {
      for (int i = 0; i < nc; ++i)
        if (!(used & (1u << i))) {
          start = i;
          break;
        }
      sv = va[start];
      closed = true;
    }
    if (EMIT)
      out[nrec] = make_int2(sv, (int)((uint32_t)row_position(row, deg, sv) | (FAN_SEED << 8) | ((uint32_t)k_self << 13)));
    ++nrec;
    int32_t cur = sv;
    int j = start;
    bool first = true;
    while (true) {
      const int32_t next = (va[j] == cur) ? vb[j] : va[j];
      used |= 1u << j;
      int jn = -1;
      for (int i = 0; i < nc; ++i)
        if (!(used & (1u << i)) && (va[i] == next || vb[i] == next)) {
          jn = i;
          break;
        }
      const bool last = jn < 0;
      if (last && closed && next != sv) return -1;
      if (EMIT) {
        const uint32_t fl = ((first && closed) ? FAN_HOLD_A : 0u) |
                            (last ? FAN_LAST : 0u) | ((last && closed) ? FAN_ADD_FIRST : 0u);
        const uint32_t mid = mat_id ? (uint32_t)mat_id[cs[j] >> 2] : 0u;
        out[nrec] = make_int2(next, (int)((uint32_t)row_position(row, deg, next) | (fl << 8) | (mid << 13)));
      }
      ++nrec;
      if (last) break;
      first = false;
      cur = next;
      j = jn;
    }
  }
  if (EMIT && chains > 1) out[0].y |= (int)(FAN_MULTI << 8);  // several fans around this node: the kernels' general loop
  return nrec;
}

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
                               int32_t *__restrict__ corner_tmp, int npe = 3) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n3) return;
  const int32_t nd = conn[i];
  if (nd < 0 || nd >= n_owned) return;
  const int32_t pos = corner_ptr[nd] + atomicAdd(cursor + nd, 1);
  const int64_t e = i / npe;
  corner_tmp[pos] = (int32_t)((e << 2) | (i - npe * e));
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
                                                       int32_t *__restrict__ ndeg, const int32_t *__restrict__ mat_id,
                                                       int32_t *__restrict__ nfan, PlanFlags *__restrict__ flags) {
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
  if (m > *reinterpret_cast<volatile int *>(&flags->max_degree)) atomicMax(&flags->max_degree, m);  // rarely taken
  if (m > kMaxDegree) flags->too_dense = 1;
  const int nf = fan_walk<false>(n, nc, cs, conn, mat_id, nullptr, 0, nullptr);
  nfan[n] = nf < 0 ? 0 : nf;
  if (nf < 0) flags->fan_irregular = 1;
}

// One thread per owned node: publish the adjacency row and the corner records.
__global__ void __launch_bounds__(128) k_node_records(int32_t n_owned, const int32_t *__restrict__ conn,
                                                     const int32_t *__restrict__ corner_ptr,
                                                     const int32_t *__restrict__ corner_tmp,
                                                     const int32_t *__restrict__ cand,
                                                     const int32_t *__restrict__ adj_ptr, int32_t *__restrict__ adj,
                                                     int2 *__restrict__ corner_rec,
                                                     const int32_t *__restrict__ mat_id,
                                                     const int32_t *__restrict__ fan_ptr, int2 *__restrict__ fan_rec) {
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
  if (fan_rec) fan_walk<true>(n, c1 - c0, corner_tmp + c0, conn, mat_id, cd, deg, fan_rec + fan_ptr[n]);
}

__global__ void k_conn4(int64_t n_elems, const int32_t *__restrict__ conn, const int32_t *__restrict__ mat_id,
                        int4 *__restrict__ conn4, PlanFlags *__restrict__ flags) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int32_t mid = mat_id ? mat_id[e] : 0;
  if (mid < 0) flags->bad_mat = 1;
  if (mid > *reinterpret_cast<volatile int *>(&flags->max_mat)) atomicMax(&flags->max_mat, mid);  // rarely taken
  conn4[e] = make_int4(conn[3 * e], conn[3 * e + 1], conn[3 * e + 2], mid);
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

__global__ void k_fan_tile_max(int32_t n_owned, const int32_t *__restrict__ fan_ptr, int *__restrict__ out) {
  const int32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const int32_t n0 = t * 32;  // assemble.cu: kFanChunk (one warp's 32-node chunk)
  if (n0 >= n_owned) return;
  const int32_t n1 = min(n0 + 32, n_owned);
  const int cnt = fan_ptr[n1] - fan_ptr[n0];
  if (cnt > *reinterpret_cast<volatile int *>(out)) atomicMax(out, cnt);
}

// 8-byte fan records -> 4-byte words + one header word per node (layout in plan.cuh).  One thread per node.
__global__ void __launch_bounds__(128) k_fan_compact(int32_t n_owned, const int32_t *__restrict__ fan_ptr,
                                                    const int2 *__restrict__ rec, uint32_t *__restrict__ rec4,
                                                    uint32_t *__restrict__ hdr, int *__restrict__ bad) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t f0 = fan_ptr[n], f1 = fan_ptr[n + 1];
  uint32_t kself = 0;
  int mat0 = -1, mat1 = -1, cur = -1, n_seeds = 0;
  for (int32_t f = f0; f < f1; ++f) {
    const int2 r = rec[f];
    const uint32_t y = (uint32_t)r.y;
    const uint32_t k = y & 255, fl = (y >> 8) & 31, hi = y >> 13;
    uint32_t f4 = (fl & FAN_SEED ? FAN4_SEED : 0u) | (fl & FAN_LAST ? FAN4_LAST : 0u) |
                  (fl & FAN_ADD_FIRST ? FAN4_ADD_FIRST : 0u);
    const int32_t delta = r.x - n;
    if (delta < -(1 << (kFan4FieldBits - 1)) || delta >= (1 << (kFan4FieldBits - 1))) *bad = 1;
    const uint32_t field = (uint32_t)delta & ((1u << kFan4FieldBits) - 1u);
    if (fl & FAN_SEED) {
      kself = hi;
      ++n_seeds;
    } else {
      const int mid = (int)hi;
      if (mat0 < 0) mat0 = cur = mid;
      if (mid != cur) {
        if (mat1 < 0 && mid != mat0) mat1 = mid;
        if (mid != mat0 && mid != mat1) *bad = 1;  // a third material around one node
        f4 |= FAN4_MATSW;
        cur = mid;
      }
      if (mid >= 4096) *bad = 1;
    }
    rec4[f] = k | (f4 << 8) | (field << kFan4Shift);
  }
  if (n_seeds > 1) rec4[f0] |= FAN4_MULTI << 8;  // (= FAN_MULTI of the 8-byte record)
  hdr[n] = kself | ((uint32_t)(mat0 < 0 ? 0 : mat0) << 8) | ((uint32_t)(mat1 < 0 ? 0 : mat1) << 20);
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
  cudaFree(p->fan_ptr);
  cudaFree(p->fan_rec);
  cudaFree(p->fan_rec4);
  cudaFree(p->fan_hdr);
  cudaFree(p->tile_eptr);
  cudaFree(p->tile_nptr);
  cudaFree(p->tile_desc);
  cudaFree(p->tile_elist);
  cudaFree(p->tile_erec);
  cudaFree(p->tile_nodes);
  cudaFree(p->contrib16);
  cudaFree(p->tet_kself);
  cudaFree(p->corner_elem);
  cudaFree(p->contrib_ptr);
  cudaFree(p->contrib);
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
  int32_t *cursor = nullptr, *corner_tmp = nullptr, *cand = nullptr, *ndeg = nullptr, *nfan = nullptr;
  PlanFlags *flags = nullptr;
  int64_t *totals = nullptr;
  const int64_t n3 = 3 * n_elems;
  PlanFlags hflags = {0, 0, 0, 0, 0, 0};
  int64_t htot[3] = {0, 0, 0};

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
  PLAN_TRY(dev_alloc(&p->adj_ptr, (int64_t)n_owned + 1 + 136, &p->bytes));  // +136: TMA tile-slice over-read
  PLAN_TRY(dev_alloc(&p->conn4, n_elems, &p->bytes));
  PLAN_TRY(dev_alloc(&cursor, (int64_t)n_owned + 1, nullptr));
  PLAN_TRY(dev_alloc(&ndeg, (int64_t)n_owned + 1, nullptr));
  PLAN_TRY(dev_alloc(&flags, 1, nullptr));
  PLAN_TRY(dev_alloc(&totals, 3, nullptr));
  PLAN_TRY(dev_alloc(&nfan, (int64_t)n_owned + 1, nullptr));
  PLAN_TRY(dev_alloc(&p->fan_ptr, (int64_t)n_owned + 1 + 136, &p->bytes));
  PLAN_CUDA(cudaMemsetAsync(cursor, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  PLAN_CUDA(cudaMemsetAsync(ndeg, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  PLAN_CUDA(cudaMemsetAsync(flags, 0, sizeof(PlanFlags), st));
  PLAN_CUDA(cudaMemsetAsync(totals, 0, 3 * sizeof(int64_t), st));
  PLAN_CUDA(cudaMemsetAsync(nfan, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));

  if (n_elems > 0 && n_owned > 0) {
    // 1. histogram of corners per owned node (cursor doubles as the count array)
    k_count_corners<<<grid_for(n3, 256), 256, 0, st>>>(n3, conn, n_nodes, n_owned, cursor, flags);
    PLAN_LAUNCHED();
    k_conn4<<<grid_for(n_elems, 256), 256, 0, st>>>(n_elems, conn, mat_id, p->conn4, flags);
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
  if (hflags.bad_mat) {
    rc = fail(FE_ERR_ARG, "fe_plan_create: negative material index");
    goto done;
  }
  p->max_mat_id = hflags.max_mat;
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
                                                            mat_id, nfan, flags);
    PLAN_LAUNCHED();
  }
  // 3. row lengths -> block row pointer
  PLAN_TRY(exclusive_scan_i32(ctx, st, ndeg, p->adj_ptr, n_owned, totals + 1));
  PLAN_TRY(exclusive_scan_i32(ctx, st, nfan, p->fan_ptr, n_owned, totals + 2));
  PLAN_CUDA(cudaMemcpyAsync(htot, totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
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
  p->fan_ok = !hflags.fan_irregular;
  p->n_fan = p->fan_ok ? htot[2] : 0;
  if (p->fan_ok) PLAN_TRY(dev_alloc(&p->fan_rec, p->n_fan + 2, &p->bytes));  // +2: 16-byte staging over-read
  if (p->n_corners > 0) {
    k_node_records<<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, p->adj_ptr,
                                                          p->adj, p->corner_rec, mat_id, p->fan_ptr,
                                                          p->fan_ok ? p->fan_rec : nullptr);
    PLAN_LAUNCHED();
  }
  if (p->fan_ok && n_owned > 0) {
    PLAN_CUDA(cudaMemsetAsync(&flags->max_degree, 0, sizeof(int), st));  // reuse as the tile maximum
    k_fan_tile_max<<<grid_for((n_owned + 31) / 32, 128), 128, 0, st>>>(n_owned, p->fan_ptr,
                                                                                 &flags->max_degree);
    PLAN_LAUNCHED();
    PLAN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
  }
  PLAN_CUDA(cudaStreamSynchronize(st));
  if (p->fan_ok) p->fan_tile_max = hflags.max_degree;
  if (p->fan_ok && n_owned > 0 && p->n_fan > 0) {  // compact (4-byte) copy of the fan records when they fit
    PLAN_TRY(dev_alloc(&p->fan_rec4, p->n_fan + 8, &p->bytes));  // +8: 16-byte staging over-read
    PLAN_TRY(dev_alloc(&p->fan_hdr, (int64_t)n_owned + 136, &p->bytes));  // +136: slice over-read
    PLAN_CUDA(cudaMemsetAsync(&flags->bad_node, 0, sizeof(int), st));
    k_fan_compact<<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, p->fan_ptr, p->fan_rec, p->fan_rec4, p->fan_hdr,
                                                         &flags->bad_node);
    PLAN_LAUNCHED();
    PLAN_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
    PLAN_CUDA(cudaStreamSynchronize(st));
    p->fan_compact_ok = hflags.bad_node == 0;
  }

done:
  cudaFree(cursor);
  cudaFree(corner_tmp);
  cudaFree(cand);
  cudaFree(ndeg);
  cudaFree(nfan);
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


}  // extern "C"

// ---------------------------------------------------------------------------------------
// Linear tetrahedra (3 DOF per node): the same node-graph construction with 4 corners per element
// (ElasticityTetrahedralElement3D, elements.py:663-876; get_row_col_indices analysis.py:714-735 for
// 12 x 12 element matrices).  Instead of the triangles' fan records the plan lists, per off-diagonal
// block (node, neighbour), the elements that contain both -- k_tet_assemble_slots gives every block
// of the global matrix to one lane.
// ---------------------------------------------------------------------------------------
namespace fe {

__global__ void __launch_bounds__(128) k_tet_node_adjacency(int32_t n_owned, const int32_t *__restrict__ conn,
                                                           const int32_t *__restrict__ corner_ptr,
                                                           int32_t *__restrict__ corner_tmp, int32_t *__restrict__ cand,
                                                           int32_t *__restrict__ ndeg, int32_t *__restrict__ corner_elem,
                                                           PlanFlags *__restrict__ flags) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
  int32_t *cs = corner_tmp + c0;
  const int nc = c1 - c0;
  insertion_sort(cs, nc);  // (element << 2 | vertex): ascending element id
  int32_t *cd = cand + 4 * (int64_t)c0;
  for (int k = 0; k < nc; ++k) {
    const int64_t e = cs[k] >> 2;
    corner_elem[c0 + k] = (int32_t)e;
    if (k > 0 && (cs[k - 1] >> 2) == e) flags->fan_irregular = 1;  // the element lists this node twice
    for (int j = 0; j < 4; ++j) cd[4 * k + j] = conn[4 * e + j];
  }
  insertion_sort(cd, 4 * nc);
  int m = 0;
  for (int k = 0; k < 4 * nc; ++k)
    if (m == 0 || cd[k] != cd[m - 1]) cd[m++] = cd[k];
  ndeg[n] = m;
  if (m > *reinterpret_cast<volatile int *>(&flags->max_degree)) atomicMax(&flags->max_degree, m);
  if (m > kMaxDegree) flags->too_dense = 1;
}

// One thread per owned node, slot-major: for every neighbour slot the corners whose element holds that
// neighbour, in ascending element order.  FILL = false counts, FILL = true writes the codes.
template <bool FILL>
__global__ void __launch_bounds__(128) k_tet_contrib(int32_t n_owned, const int32_t *__restrict__ conn,
                                                    const int32_t *__restrict__ corner_ptr,
                                                    const int32_t *__restrict__ corner_tmp,
                                                    const int32_t *__restrict__ cand, const int32_t *__restrict__ adj_ptr,
                                                    int32_t *__restrict__ adj, int32_t *__restrict__ cnt,
                                                    const int32_t *__restrict__ contrib_ptr, int32_t *__restrict__ contrib) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
  const int32_t a0 = adj_ptr[n];
  const int deg = adj_ptr[n + 1] - a0;
  if (deg > kMaxDegree) return;
  const int32_t *cd = cand + 4 * (int64_t)c0;
  for (int k = 0; k < deg; ++k) {
    const int32_t m = cd[k];
    if (!FILL) adj[a0 + k] = m;
    int32_t count = 0;
    if (m != n) {
      int32_t w = FILL ? contrib_ptr[a0 + k] : 0;
      for (int32_t c = c0; c < c1; ++c) {
        const int32_t code = corner_tmp[c];
        const int64_t e = code >> 2;
        const int v = code & 3;
        for (int j = 0; j < 4; ++j) {
          if (j == v || conn[4 * e + j] != m) continue;
          if (FILL) contrib[w++] = (int32_t)((e << 4) | (v << 2) | j);
          ++count;
        }
      }
    }
    if (!FILL) cnt[a0 + k] = count;
  }
}

// ---------------------------------------------------------------------------------------
// Tiles of the staged assembly (plan.cuh): one CTA per tile of kTetStageNodes owned nodes.
//  * elements: a corner (node i, element e) of the tile's corner range is the element's FIRST appearance
//    when i is the smallest tile node of e; first appearances are numbered by an exclusive scan
//    -> tile-local element index (order: by node, then ascending element id).
//  * nodes: the tile's adjacency range, sorted (bitonic, shared memory) and made unique.
//  * contribution codes: the element of every code is looked up in the corner list of its smallest tile
//    node and replaced by the tile-local index.
// FILL = false counts (tile_ecnt, tile_ncnt, maxima), FILL = true writes the lists.
// ---------------------------------------------------------------------------------------
constexpr int kTetStageCornerCap = 2048;  // corners per tile (16 nodes: <= 128 elements per node on average)
constexpr int kTetStageAdjCap = 2048;     // adjacency entries per tile
struct TetStageFlags {
  int bad, max_elems, max_nodes, max_contrib, max_adj;
};

__device__ __forceinline__ int tile_first_node(const int32_t *__restrict__ conn, int64_t e, int32_t n0, int32_t n1) {
  int32_t imin = INT32_MAX;
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    const int32_t m = conn[4 * e + v];
    if (m >= n0 && m < n1 && m < imin) imin = m;
  }
  return imin;
}

// exclusive scan of the 0/1 flags in s[0..n) by 256 threads; s[c] becomes prefix | flag << 31; returns the total
__device__ int tile_scan_flags(int32_t *s, int n, int32_t *part) {
  const int tid = threadIdx.x;
  const int per = (n + 255) / 256, beg = min(tid * per, n), end = min(beg + per, n);
  int sum = 0;
  for (int c = beg; c < end; ++c) sum += s[c];
  part[tid] = sum;
  __syncthreads();
  if (tid == 0) {
    int run = 0;
    for (int t = 0; t < 256; ++t) {
      const int v = part[t];
      part[t] = run;
      run += v;
    }
    part[256] = run;
  }
  __syncthreads();
  int run = part[tid];
  for (int c = beg; c < end; ++c) {
    const int f = s[c];
    s[c] = run | (f ? (int32_t)0x80000000 : 0);
    run += f;
  }
  __syncthreads();
  return part[256];
}

__device__ __forceinline__ int lower_bound_i32(const int32_t *a, int n, int32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

template <bool FILL>
__global__ void __launch_bounds__(256) k_tet_stage_tiles(
    int32_t n_owned, const int32_t *__restrict__ conn, const int32_t *__restrict__ corner_ptr,
    const int32_t *__restrict__ corner_elem, const int32_t *__restrict__ adj_ptr, const int32_t *__restrict__ adj,
    const int32_t *__restrict__ contrib_ptr, const int32_t *__restrict__ contrib, int32_t *__restrict__ tile_ecnt,
    int32_t *__restrict__ tile_ncnt, const int32_t *__restrict__ tile_eptr, const int32_t *__restrict__ tile_nptr,
    int32_t *__restrict__ tile_elist, ushort4 *__restrict__ tile_erec, int32_t *__restrict__ tile_nodes,
    uint16_t *__restrict__ contrib16, uint8_t *__restrict__ kself, int4 *__restrict__ tile_desc,
    TetStageFlags *__restrict__ flags) {
  __shared__ int32_t s_pre[kTetStageCornerCap];
  __shared__ int32_t s_sort[kTetStageAdjCap];
  __shared__ int32_t s_uflag[kTetStageAdjCap];
  __shared__ int32_t s_uniq[kTetStageAdjCap];
  __shared__ int32_t s_cptr[kTetStageNodes + 1];
  __shared__ int32_t s_part[257];
  const int tid = threadIdx.x;
  const int32_t tile = blockIdx.x, n0 = tile * kTetStageNodes, n1 = min(n0 + kTetStageNodes, n_owned);
  const int nt = n1 - n0;
  if (tid <= nt) s_cptr[tid] = corner_ptr[n0 + tid];
  __syncthreads();
  const int32_t c0 = s_cptr[0];
  const int nc = s_cptr[nt] - c0;
  const int32_t a0 = adj_ptr[n0];
  const int na = adj_ptr[n1] - a0;
  const int32_t q0 = contrib_ptr[a0];
  const int nq = contrib_ptr[a0 + na] - q0;
  if (nc > kTetStageCornerCap || na > kTetStageAdjCap) {  // (uniform across the CTA)
    if (tid == 0) flags->bad = 1;
    return;
  }
  // ---- elements: first appearances within the tile
  for (int c = tid; c < nc; c += 256) {
    int lo = 0, hi = nt - 1;  // node of the corner: last i with s_cptr[i] <= c0 + c
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_cptr[mid] <= c0 + c) lo = mid; else hi = mid - 1;
    }
    s_pre[c] = tile_first_node(conn, corner_elem[c0 + c], n0, n1) == n0 + lo;
  }
  __syncthreads();
  const int ne = tile_scan_flags(s_pre, nc, s_part);
  // ---- nodes: sorted unique adjacency of the tile
  int npad = 1;
  while (npad < na) npad <<= 1;
  for (int k = tid; k < npad; k += 256) s_sort[k] = k < na ? adj[a0 + k] : INT32_MAX;
  __syncthreads();
  for (int size = 2; size <= npad; size <<= 1)
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int k = tid; k < npad; k += 256) {
        const int partner = k ^ stride;
        if (partner > k) {
          const bool up = (k & size) == 0;
          const int32_t x = s_sort[k], y = s_sort[partner];
          if ((x > y) == up) {
            s_sort[k] = y;
            s_sort[partner] = x;
          }
        }
      }
      __syncthreads();
    }
  for (int k = tid; k < na; k += 256) s_uflag[k] = (k == 0 || s_sort[k] != s_sort[k - 1]);
  __syncthreads();
  const int nn = tile_scan_flags(s_uflag, na, s_part);
  for (int k = tid; k < na; k += 256)
    if (s_uflag[k] < 0) s_uniq[s_uflag[k] & 0x7fffffff] = s_sort[k];
  __syncthreads();
  if (!FILL) {
    if (tid == 0) {
      tile_ecnt[tile] = ne;
      tile_ncnt[tile] = nn;
      atomicMax(&flags->max_elems, ne);
      atomicMax(&flags->max_nodes, nn);
      atomicMax(&flags->max_contrib, nq);
      atomicMax(&flags->max_adj, na);
      if (ne > 4095 || nn > 65535) flags->bad = 1;
    }
    return;
  }
  const int32_t eb = tile_eptr[tile], nb = tile_nptr[tile];
  if (tid == 0) {
    tile_desc[2 * tile] = make_int4(a0, na, eb, ne);
    tile_desc[2 * tile + 1] = make_int4(nb, nn, q0, nq);
  }
  for (int k = tid; k < nn; k += 256) tile_nodes[nb + k] = s_uniq[k];
  for (int c = tid; c < nc; c += 256) {
    if (s_pre[c] >= 0) continue;
    const int le = s_pre[c] & 0x7fffffff;
    const int64_t e = corner_elem[c0 + c];
    tile_elist[eb + le] = (int32_t)e;
    ushort4 r;
    r.x = (unsigned short)lower_bound_i32(s_uniq, nn, conn[4 * e + 0]);
    r.y = (unsigned short)lower_bound_i32(s_uniq, nn, conn[4 * e + 1]);
    r.z = (unsigned short)lower_bound_i32(s_uniq, nn, conn[4 * e + 2]);
    r.w = (unsigned short)lower_bound_i32(s_uniq, nn, conn[4 * e + 3]);
    tile_erec[eb + le] = r;
  }
  if (tid < nt) {
    const int32_t i = n0 + tid, b0 = adj_ptr[i], d = adj_ptr[i + 1] - b0;
    kself[i] = (uint8_t)(d > 0 ? lower_bound_i32(adj + b0, d, i) : 0);
  }
  for (int q = tid; q < nq; q += 256) {
    const int32_t code = contrib[q0 + q];
    const int32_t e = code >> 4;
    const int li = tile_first_node(conn, e, n0, n1) - n0;
    const int32_t cb = s_cptr[li];
    const int cpos = cb + lower_bound_i32(corner_elem + cb, s_cptr[li + 1] - cb, e) - c0;
    contrib16[q0 + q] = (uint16_t)(((s_pre[cpos] & 0x7fffffff) << 4) | (code & 15));
  }
}

}  // namespace fe

extern "C" {

int fe_tet_plan_create(fe_ctx *ctx, void *stream, int32_t n_nodes, int32_t n_owned, int64_t n_elems, const int32_t *conn,
                       fe_plan **out) {
  FE_REQUIRE(ctx && out, "fe_tet_plan_create: NULL ctx/out");
  FE_REQUIRE(n_nodes >= 0 && n_owned >= 0 && n_owned <= n_nodes, "fe_tet_plan_create: bad node counts %d/%d", n_owned, n_nodes);
  FE_REQUIRE(n_elems >= 0 && (n_elems == 0 || conn), "fe_tet_plan_create: bad connectivity");
  if (n_elems >= (int64_t(1) << 27)) return fail(FE_ERR_UNSUPPORTED, "fe_tet_plan_create: more than 2^27 elements");
  if ((int64_t)n_nodes * 3 >= (int64_t(1) << 31)) return fail(FE_ERR_UNSUPPORTED, "fe_tet_plan_create: DOF count overflows int32");
  cudaStream_t st = as_stream(stream);
  FE_CUDA(cudaSetDevice(ctx->device));
  fe_plan *p = new fe_plan();
  p->ctx = ctx;
  p->n_nodes = n_nodes;
  p->n_owned = n_owned;
  p->dim = 3;
  p->npe = 4;
  p->n_elems = n_elems;
  int rc = FE_OK;
  int32_t *cursor = nullptr, *corner_tmp = nullptr, *cand = nullptr, *ndeg = nullptr, *cnt = nullptr, *tcnt = nullptr;
  PlanFlags *flags = nullptr;
  TetStageFlags *sflags = nullptr;
  int64_t *totals = nullptr, *stot = nullptr;
  const int64_t n4 = 4 * n_elems;
  PlanFlags hflags = {0, 0, 0, 0, 0, 0};
  int64_t htot[3] = {0, 0, 0};
#define TP_TRY(expr)            \
  do {                          \
    rc = (expr);                \
    if (rc != FE_OK) goto done; \
  } while (0)
#define TP_CUDA(call)                                                                                       \
  do {                                                                                                      \
    cudaError_t e__ = (call);                                                                               \
    if (e__ != cudaSuccess) {                                                                               \
      rc = fail(FE_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
      goto done;                                                                                            \
    }                                                                                                       \
  } while (0)
#define TP_LAUNCHED()                                                                                      \
  do {                                                                                                     \
    ctx->launches++;                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                                  \
    if (e__ != cudaSuccess) {                                                                              \
      rc = fail(FE_ERR_CUDA, "launch failed at %s:%d: %s", __FILE__, __LINE__, cudaGetErrorString(e__));   \
      goto done;                                                                                           \
    }                                                                                                      \
  } while (0)
  TP_TRY(dev_alloc(&p->corner_ptr, (int64_t)n_owned + 1, &p->bytes));
  TP_TRY(dev_alloc(&p->adj_ptr, (int64_t)n_owned + 1 + 136, &p->bytes));
  TP_TRY(dev_alloc(&cursor, (int64_t)n_owned + 1, nullptr));
  TP_TRY(dev_alloc(&ndeg, (int64_t)n_owned + 1, nullptr));
  TP_TRY(dev_alloc(&flags, 1, nullptr));
  TP_TRY(dev_alloc(&totals, 3, nullptr));
  TP_CUDA(cudaMemsetAsync(cursor, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  TP_CUDA(cudaMemsetAsync(ndeg, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
  TP_CUDA(cudaMemsetAsync(flags, 0, sizeof(PlanFlags), st));
  TP_CUDA(cudaMemsetAsync(totals, 0, 3 * sizeof(int64_t), st));
  if (n_elems > 0 && n_owned > 0) {
    k_count_corners<<<grid_for(n4, 256), 256, 0, st>>>(n4, conn, n_nodes, n_owned, cursor, flags);
    TP_LAUNCHED();
  }
  TP_TRY(exclusive_scan_i32(ctx, st, cursor, p->corner_ptr, n_owned, totals + 0));
  TP_CUDA(cudaMemcpyAsync(htot, totals, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TP_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
  TP_CUDA(cudaStreamSynchronize(st));
  if (hflags.bad_node) {
    rc = fail(FE_ERR_ARG, "fe_tet_plan_create: connectivity references a node outside [0, %d)", n_nodes);
    goto done;
  }
  p->n_corners = htot[0];
  TP_TRY(dev_alloc(&p->corner_elem, p->n_corners, &p->bytes));
  TP_TRY(dev_alloc(&corner_tmp, p->n_corners, nullptr));
  TP_TRY(dev_alloc(&cand, 4 * p->n_corners, nullptr));
  if (p->n_corners > 0) {
    TP_CUDA(cudaMemsetAsync(cursor, 0, ((size_t)n_owned + 1) * sizeof(int32_t), st));
    k_fill_corners<<<grid_for(n4, 256), 256, 0, st>>>(n4, conn, n_nodes, n_owned, p->corner_ptr, cursor, corner_tmp, 4);
    TP_LAUNCHED();
    k_tet_node_adjacency<<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, ndeg,
                                                                p->corner_elem, flags);
    TP_LAUNCHED();
  }
  TP_TRY(exclusive_scan_i32(ctx, st, ndeg, p->adj_ptr, n_owned, totals + 1));
  TP_CUDA(cudaMemcpyAsync(htot, totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TP_CUDA(cudaMemcpyAsync(&hflags, flags, sizeof(PlanFlags), cudaMemcpyDeviceToHost, st));
  TP_CUDA(cudaStreamSynchronize(st));
  p->nnzb = htot[1];
  p->nnz = p->nnzb * 9;
  p->max_degree = hflags.max_degree;
  p->tet_degenerate = hflags.fan_irregular != 0;
  if (hflags.too_dense) {
    rc = fail(FE_ERR_UNSUPPORTED, "fe_tet_plan_create: a node has %d neighbours (limit %d)", hflags.max_degree, kMaxDegree);
    goto done;
  }
  if (p->nnz >= (int64_t(1) << 31)) {
    rc = fail(FE_ERR_UNSUPPORTED, "fe_tet_plan_create: nnz = %lld overflows int32 CSR indices", (long long)p->nnz);
    goto done;
  }
  TP_TRY(dev_alloc(&p->adj, p->nnzb, &p->bytes));
  TP_TRY(dev_alloc(&cnt, p->nnzb + 1, nullptr));
  TP_TRY(dev_alloc(&p->contrib_ptr, p->nnzb + 1, &p->bytes));
  TP_CUDA(cudaMemsetAsync(cnt, 0, ((size_t)p->nnzb + 1) * sizeof(int32_t), st));
  if (p->n_corners > 0) {
    k_tet_contrib<false><<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, p->adj_ptr,
                                                              p->adj, cnt, nullptr, nullptr);
    TP_LAUNCHED();
  }
  TP_TRY(exclusive_scan_i32(ctx, st, cnt, p->contrib_ptr, p->nnzb, totals + 2));
  TP_CUDA(cudaMemcpyAsync(htot, totals, 3 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  TP_CUDA(cudaStreamSynchronize(st));
  p->n_contrib = htot[2];
  if (p->n_contrib >= (int64_t(1) << 31)) {
    rc = fail(FE_ERR_UNSUPPORTED, "fe_tet_plan_create: %lld block contributions overflow int32", (long long)p->n_contrib);
    goto done;
  }
  TP_TRY(dev_alloc(&p->contrib, p->n_contrib, &p->bytes));
  if (p->n_contrib > 0) {
    k_tet_contrib<true><<<grid_for(n_owned, 128), 128, 0, st>>>(n_owned, conn, p->corner_ptr, corner_tmp, cand, p->adj_ptr,
                                                             p->adj, nullptr, p->contrib_ptr, p->contrib);
    TP_LAUNCHED();
  }
  TP_CUDA(cudaStreamSynchronize(st));
  // ---- tiles of the staged assembly variant
  if (!p->tet_degenerate && n_owned > 0 && p->n_contrib > 0) {
    TetStageFlags hs = {0, 0, 0, 0, 0};
    int64_t hst[2] = {0, 0};
    p->n_tiles = (n_owned + kTetStageNodes - 1) / kTetStageNodes;
    TP_TRY(dev_alloc(&sflags, 1, nullptr));
    TP_TRY(dev_alloc(&stot, 2, nullptr));
    TP_TRY(dev_alloc(&tcnt, 2 * ((int64_t)p->n_tiles + 1), nullptr));
    TP_TRY(dev_alloc(&p->tile_eptr, (int64_t)p->n_tiles + 1, &p->bytes));
    TP_TRY(dev_alloc(&p->tile_nptr, (int64_t)p->n_tiles + 1, &p->bytes));
    TP_CUDA(cudaMemsetAsync(sflags, 0, sizeof(TetStageFlags), st));
    TP_CUDA(cudaMemsetAsync(tcnt, 0, 2 * ((size_t)p->n_tiles + 1) * sizeof(int32_t), st));
    k_tet_stage_tiles<false><<<p->n_tiles, 256, 0, st>>>(n_owned, conn, p->corner_ptr, p->corner_elem, p->adj_ptr, p->adj,
                                                        p->contrib_ptr, p->contrib, tcnt, tcnt + p->n_tiles + 1, nullptr,
                                                        nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, sflags);
    TP_LAUNCHED();
    TP_TRY(exclusive_scan_i32(ctx, st, tcnt, p->tile_eptr, p->n_tiles, stot + 0));
    TP_TRY(exclusive_scan_i32(ctx, st, tcnt + p->n_tiles + 1, p->tile_nptr, p->n_tiles, stot + 1));
    TP_CUDA(cudaMemcpyAsync(hst, stot, 2 * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaMemcpyAsync(&hs, sflags, sizeof(TetStageFlags), cudaMemcpyDeviceToHost, st));
    TP_CUDA(cudaStreamSynchronize(st));
    if (!hs.bad && hst[0] < (int64_t(1) << 31) && hst[1] < (int64_t(1) << 31)) {
      TP_TRY(dev_alloc(&p->tile_desc, 2 * (int64_t)p->n_tiles, &p->bytes));
      TP_TRY(dev_alloc(&p->tile_elist, hst[0], &p->bytes));
      TP_TRY(dev_alloc(&p->tile_erec, hst[0], &p->bytes));
      TP_TRY(dev_alloc(&p->tile_nodes, hst[1], &p->bytes));
      TP_TRY(dev_alloc(&p->contrib16, p->n_contrib + 8, &p->bytes));
      TP_TRY(dev_alloc(&p->tet_kself, (int64_t)n_owned + 16, &p->bytes));  // +16: word-wise staging over-read
      k_tet_stage_tiles<true><<<p->n_tiles, 256, 0, st>>>(n_owned, conn, p->corner_ptr, p->corner_elem, p->adj_ptr, p->adj,
                                                         p->contrib_ptr, p->contrib, nullptr, nullptr, p->tile_eptr,
                                                         p->tile_nptr, p->tile_elist, p->tile_erec, p->tile_nodes,
                                                         p->contrib16, p->tet_kself, p->tile_desc, sflags);
      TP_LAUNCHED();
      TP_CUDA(cudaStreamSynchronize(st));
      p->tile_elems_max = hs.max_elems;
      p->tile_nodes_max = hs.max_nodes;
      p->tile_contrib_max = hs.max_contrib;
      p->tile_adj_max = hs.max_adj;
      p->tet_stage_ok = true;
    }
  }
done:
  cudaFree(sflags);
  cudaFree(stot);
  cudaFree(tcnt);
  cudaFree(cursor);
  cudaFree(corner_tmp);
  cudaFree(cand);
  cudaFree(ndeg);
  cudaFree(cnt);
  cudaFree(flags);
  cudaFree(totals);
  if (rc != FE_OK) {
    fe_plan_destroy(p);
    return rc;
  }
  *out = p;
  return FE_OK;
#undef TP_TRY
#undef TP_CUDA
#undef TP_LAUNCHED
}

int64_t fe_plan_nnz(const fe_plan *p) { return p ? p->nnz : 0; }
int32_t fe_plan_n_rows(const fe_plan *p) { return p ? p->n_owned * p->dim : 0; }
int32_t fe_plan_max_degree(const fe_plan *p) { return p ? p->max_degree : 0; }
int64_t fe_plan_bytes(const fe_plan *p) { return p ? p->bytes : 0; }
int32_t fe_plan_fan_record_bytes(const fe_plan *p) { return (p && p->fan_ok) ? (p->fan_compact_ok ? 4 : 8) : 0; }

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
