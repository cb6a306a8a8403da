// ctx / error plumbing, the int32 scan used by the symbolic phase, the material table and
// the per-element kernels (fe_elem_matrices, fe_source_factors).
#include <stdarg.h>

#include "elem.cuh"

namespace fe {

static thread_local char g_err[512];

char *last_error_buf() { return g_err; }

int fail(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

// ---------------------------------------------------------------------------------------
// exclusive scan (three passes; block sums scanned by one CTA in fixed order)
// ---------------------------------------------------------------------------------------
constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__global__ void __launch_bounds__(kScanBlock) k_scan_tile_sums(const int32_t *__restrict__ in, int64_t n,
                                                              int64_t *__restrict__ tile_sums) {
  __shared__ long long warp_part[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile;
  long long s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + (int64_t)k * kScanBlock + threadIdx.x;
    if (i < n) s += in[i];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) warp_part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    long long t = 0;
    for (int w = 0; w < kScanBlock / 32; ++w) t += warp_part[w];
    tile_sums[blockIdx.x] = t;
  }
}

__global__ void __launch_bounds__(1024) k_scan_tile_offsets(int64_t *__restrict__ tile_sums, int64_t n_tiles,
                                                           int64_t *__restrict__ total) {
  // single CTA: sequential over chunks of 1024 tiles, in-chunk Hillis-Steele scan
  __shared__ long long buf[1024];
  __shared__ long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t c = 0; c < n_tiles; c += 1024) {
    int64_t i = c + threadIdx.x;
    long long v = (i < n_tiles) ? tile_sums[i] : 0;
    buf[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      long long t = (threadIdx.x >= o) ? buf[threadIdx.x - o] : 0;
      __syncthreads();
      buf[threadIdx.x] += t;
      __syncthreads();
    }
    long long incl = buf[threadIdx.x];
    if (i < n_tiles) tile_sums[i] = carry + incl - v;  // exclusive
    __syncthreads();
    if (threadIdx.x == 1023) carry += incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}

__global__ void __launch_bounds__(kScanBlock) k_scan_apply(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                          int64_t n, const int64_t *__restrict__ tile_off,
                                                          const int64_t *__restrict__ total) {
  // each thread owns kScanItems CONSECUTIVE items so that a serial in-thread scan works
  __shared__ long long warp_part[kScanBlock / 32];
  const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
  int32_t v[kScanItems];
  long long s = 0;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    v[k] = (i < n) ? in[i] : 0;
    s += v[k];
  }
  // exclusive scan of s across the block
  long long incl = s;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) warp_part[w] = incl;
  __syncthreads();
  long long woff = 0;
  for (int k = 0; k < w; ++k) woff += warp_part[k];
  long long run = tile_off[blockIdx.x] + woff + incl - s;
#pragma unroll
  for (int k = 0; k < kScanItems; ++k) {
    int64_t i = base + k;
    if (i < n) out[i] = (int32_t)run;
    run += v[k];
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (int32_t)(*total);
}

int exclusive_scan_i32(fe_ctx *ctx, cudaStream_t st, const int32_t *in, int32_t *out, int64_t n,
                       int64_t *total_dev) {
  const int64_t n_tiles = n > 0 ? (n + kScanTile - 1) / kScanTile : 1;  // n == 0 still writes out[0] = 0
  int rc = ctx->scratch_a.reserve((size_t)(n_tiles + 1) * sizeof(int64_t));
  if (rc) return rc;
  int64_t *tile_sums = (int64_t *)ctx->scratch_a.ptr;
  k_scan_tile_sums<<<(int)n_tiles, kScanBlock, 0, st>>>(in, n, tile_sums);
  FE_LAUNCH_CHECK(ctx);
  k_scan_tile_offsets<<<1, 1024, 0, st>>>(tile_sums, n_tiles, total_dev);
  FE_LAUNCH_CHECK(ctx);
  k_scan_apply<<<(int)n_tiles, kScanBlock, 0, st>>>(in, out, n, tile_sums, total_dev);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

// ---------------------------------------------------------------------------------------
// material table
// ---------------------------------------------------------------------------------------
__global__ void k_material_table(int kind, int n_mat, const double *__restrict__ mat, MatRow *__restrict__ tab) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_mat) return;
  const double p0 = mat[4 * g + 0], p1 = mat[4 * g + 1], p2 = mat[4 * g + 2], p3 = mat[4 * g + 3];
  MatRow r;
  if (kind == FE_ELAST_PSTRESS || kind == FE_ELAST_PSTRAIN) {
    // elements.py:426-429 (strain) / :444-447 (stress)
    const double a = (kind == FE_ELAST_PSTRAIN) ? (p0 * p1) / ((1 + p1) * (1 - 2 * p1)) : (p0 * p1) / (1 - p1 * p1);
    const double b = p0 / (2 * (1 + p1));
    r.p0 = (a + 2 * b) * p2;  // c * thickness
    r.p1 = a * p2;
    r.p2 = b * p2;
    r.p3 = p2;
  } else if (kind == FE_MAGNETIC) {
    r.p0 = 1.0 / p0;  // elements.py:108
    r.p1 = r.p2 = r.p3 = 0.0;
  } else {  // FE_MASS, elements.py:530-531
    r.p0 = (p3 * p2) / 12.0;
    r.p1 = r.p2 = r.p3 = 0.0;
  }
  tab[g] = r;
}

int build_material_table(fe_ctx *ctx, cudaStream_t st, int kind, const double *mat, int n_mat, MatRow **tab_out,
                         Scratch *where) {
  FE_REQUIRE(kind >= FE_ELAST_PSTRESS && kind <= FE_MASS, "unknown kind %d", kind);
  FE_REQUIRE(mat != nullptr && n_mat > 0, "material table is empty");
  int rc = where->reserve((size_t)n_mat * sizeof(MatRow));
  if (rc) return rc;
  MatRow *tab = (MatRow *)where->ptr;
  k_material_table<<<grid_for(n_mat, 128), 128, 0, st>>>(kind, n_mat, mat, tab);
  FE_LAUNCH_CHECK(ctx);
  *tab_out = tab;
  return FE_OK;
}

// ---------------------------------------------------------------------------------------
// element dump: one thread per element, Ke in FP64 registers, row-major [E][(3 dim)^2]
// ---------------------------------------------------------------------------------------
template <int KIND_CLASS>  // 0 elasticity, 1 mass, 2 magnetic
__global__ void __launch_bounds__(128) k_elem_matrices(int64_t n_elems, const double2 *__restrict__ coords,
                                                      const int32_t *__restrict__ conn,
                                                      const int32_t *__restrict__ mat_id,
                                                      const MatRow *__restrict__ tab, double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int a = conn[3 * e + 0], b = conn[3 * e + 1], c = conn[3 * e + 2];
  const TriGeom g = tri_geom(__ldg(coords + a), __ldg(coords + b), __ldg(coords + c));
  const MatRow m = tab[mat_id ? mat_id[e] : 0];
  if (KIND_CLASS == 2) {
    double *o = out + 9 * e;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      double r[3];
      mag_row(g, m, v, r);
      o[3 * v + 0] = r[0];
      o[3 * v + 1] = r[1];
      o[3 * v + 2] = r[2];
    }
  } else {
    double *o = out + 36 * e;
#pragma unroll
    for (int v = 0; v < 3; ++v) {
      Blk2 r[3];
      if (KIND_CLASS == 0)
        elast_row_blocks(g, m, v, r);
      else
        mass_row_blocks(g, m, v, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        o[(2 * v) * 6 + 2 * j] = r[j].k00;
        o[(2 * v) * 6 + 2 * j + 1] = r[j].k01;
        o[(2 * v + 1) * 6 + 2 * j] = r[j].k10;
        o[(2 * v + 1) * 6 + 2 * j + 1] = r[j].k11;
      }
    }
  }
}

// elements.py:18-53: |det| * (a_i + .5 b_i x2 + .5 c_i y2 + .5 b_i x3 + .5 c_i y3)
__global__ void __launch_bounds__(128) k_source_factors(int64_t n_sel, const int32_t *__restrict__ sel,
                                                       const double2 *__restrict__ coords,
                                                       const int32_t *__restrict__ conn, double *__restrict__ out,
                                                       double *__restrict__ out_area) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_sel) return;
  const int64_t e = sel ? sel[i] : i;
  const double2 p[3] = {__ldg(coords + conn[3 * e + 0]), __ldg(coords + conn[3 * e + 1]),
                        __ldg(coords + conn[3 * e + 2])};
  const TriGeom g = tri_geom(p[0], p[1], p[2]);
  const double det = fabs(g.cross);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    const int k1 = (k + 1) % 3, k2 = (k + 2) % 3;
    // true divisions (not a reciprocal): a_k cancels against the b_k, c_k terms (SURVEY a-6)
    const double ak = (p[k1].x * p[k2].y - p[k2].x * p[k1].y) / g.cross;
    const double bk = g.beta[k] / g.cross, ck = g.gamma[k] / g.cross;
    out[3 * i + k] = det * (ak + 0.5 * bk * p[1].x + 0.5 * ck * p[1].y + 0.5 * bk * p[2].x + 0.5 * ck * p[2].y);
  }
  if (out_area) out_area[i] = 0.5 * det;
}


// ---------------------------------------------------------------------------------------
// element post-processing: one thread per element, gathers the element's
// nodal solution and evaluates
//   elasticity: strain = B u_e, stress = D strain (results.py:809-830), energy = 1/2 u_e^T Ke u_e
//               = 1/2 t A strain . stress (results.py:769-781, elements.py:275-292)  -> out[e][7]
//   magnetic  : B = (sum c_i A_i, -sum b_i A_i) (results.py:121-152)                    -> out[e][2]
// ---------------------------------------------------------------------------------------
template <bool MAGNETIC>
__global__ void __launch_bounds__(128) k_elem_post(int64_t n_elems, const double2 *__restrict__ coords,
                                                  const int32_t *__restrict__ conn,
                                                  const int32_t *__restrict__ mat_id,
                                                  const MatRow *__restrict__ tab, const double *__restrict__ u,
                                                  double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int n0 = conn[3 * e + 0], n1 = conn[3 * e + 1], n2 = conn[3 * e + 2];
  const TriGeom g = tri_geom(__ldg(coords + n0), __ldg(coords + n1), __ldg(coords + n2));
  if (MAGNETIC) {
    const double a[3] = {u[n0], u[n1], u[n2]};
    double bx = 0.0, by = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      bx += (g.gamma[i] / g.cross) * a[i];  // c_i A_i
      by -= (g.beta[i] / g.cross) * a[i];   // -b_i A_i
    }
    out[2 * e + 0] = bx;
    out[2 * e + 1] = by;
  } else {
    const MatRow m = tab[mat_id ? mat_id[e] : 0];  // (c t, a t, b t, t)
    const double2 *u2 = reinterpret_cast<const double2 *>(u);
    const double2 d[3] = {u2[n0], u2[n1], u2[n2]};
    const double inv = 1.0 / g.det;
    double exx = 0.0, eyy = 0.0, gxy = 0.0;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      exx += g.beta[i] * d[i].x;
      eyy += g.gamma[i] * d[i].y;
      gxy += g.gamma[i] * d[i].x + g.beta[i] * d[i].y;
    }
    exx *= inv;
    eyy *= inv;
    gxy *= inv;
    const double it = 1.0 / m.p3;  // D = (c, a, b) = m.p0..2 / thickness
    const double c = m.p0 * it, a = m.p1 * it, b = m.p2 * it;
    const double sxx = c * exx + a * eyy, syy = a * exx + c * eyy, sxy = b * gxy;
    double *o = out + 7 * e;
    o[0] = exx;
    o[1] = eyy;
    o[2] = gxy;
    o[3] = sxx;
    o[4] = syy;
    o[5] = sxy;
    o[6] = 0.5 * m.p3 * (0.5 * fabs(g.cross)) * (exx * sxx + eyy * syy + gxy * sxy);
  }
}

}  // namespace fe

using namespace fe;

extern "C" {

int fe_version(void) { return FE_B200_VERSION; }

const char *fe_last_error(void) { return fe::last_error_buf(); }

int fe_ctx_create(int device, fe_ctx **out) {
  FE_REQUIRE(out != nullptr, "fe_ctx_create: out is NULL");
  int count = 0;
  FE_CUDA(cudaGetDeviceCount(&count));
  FE_REQUIRE(device >= 0 && device < count, "fe_ctx_create: device %d out of range (%d devices)", device, count);
  FE_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  FE_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(FE_ERR_UNSUPPORTED, "libfe_b200 needs sm_100a (B200); device %d is sm_%d%d", device, prop.major,
                prop.minor);
  fe_ctx *c = new fe_ctx();
  c->device = device;
  c->num_sms = prop.multiProcessorCount;
  FE_CUDA(cudaMallocHost(&c->pinned, 4096));
  cudaStream_t ws;
  cudaEvent_t we;
  FE_CUDA(cudaStreamCreateWithFlags(&ws, cudaStreamNonBlocking));
  FE_CUDA(cudaEventCreateWithFlags(&we, cudaEventDisableTiming));
  c->work_stream = ws;
  c->work_event = we;
  *out = c;
  return FE_OK;
}

int fe_ctx_destroy(fe_ctx *ctx) {
  if (!ctx) return FE_OK;
  cudaSetDevice(ctx->device);
  ctx->scratch_a.release();
  ctx->scratch_b.release();
  ctx->scratch_c.release();
  ctx->scratch_p.release();
  ctx->bc_map.release();
  ctx->scratch_g.release();
  ctx->halo_send.release();
  ctx->halo_recv.release();
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->pcg_graph) cudaGraphExecDestroy((cudaGraphExec_t)ctx->pcg_graph);
  if (ctx->work_stream) cudaStreamDestroy((cudaStream_t)ctx->work_stream);
  if (ctx->work_event) cudaEventDestroy((cudaEvent_t)ctx->work_event);
  extern __attribute__((visibility("hidden"))) void fe_dist_teardown(fe_ctx *);
  fe_dist_teardown(ctx);
  delete ctx;
  return FE_OK;
}

int64_t fe_ctx_launch_count(const fe_ctx *ctx) { return ctx ? ctx->launches : 0; }

int fe_elem_matrices(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords, const int32_t *conn,
                     const int32_t *mat_id, const double *mat, int32_t n_mat, double *out) {
  FE_REQUIRE(ctx && coords && conn && out, "fe_elem_matrices: NULL argument");
  FE_REQUIRE(n_elems >= 0, "fe_elem_matrices: negative element count");
  if (n_elems == 0) return FE_OK;
  cudaStream_t st = as_stream(stream);
  MatRow *tab = nullptr;
  int rc = build_material_table(ctx, st, kind, mat, n_mat, &tab, &ctx->scratch_b);
  if (rc) return rc;
  const int grid = grid_for(n_elems, 128);
  const double2 *xy = reinterpret_cast<const double2 *>(coords);
  if (kind == FE_MAGNETIC)
    k_elem_matrices<2><<<grid, 128, 0, st>>>(n_elems, xy, conn, mat_id, tab, out);
  else if (kind == FE_MASS)
    k_elem_matrices<1><<<grid, 128, 0, st>>>(n_elems, xy, conn, mat_id, tab, out);
  else
    k_elem_matrices<0><<<grid, 128, 0, st>>>(n_elems, xy, conn, mat_id, tab, out);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_source_factors(fe_ctx *ctx, void *stream, int64_t n_sel, const int32_t *elem_sel, const double *coords,
                      const int32_t *conn, double *out, double *out_area) {
  FE_REQUIRE(ctx && coords && conn && out, "fe_source_factors: NULL argument");
  if (n_sel <= 0) return FE_OK;
  k_source_factors<<<grid_for(n_sel, 128), 128, 0, as_stream(stream)>>>(
      n_sel, elem_sel, reinterpret_cast<const double2 *>(coords), conn, out, out_area);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_elem_post(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords, const int32_t *conn,
                 const int32_t *mat_id, const double *mat, int32_t n_mat, const double *u, double *out) {
  FE_REQUIRE(ctx && coords && conn && u && out, "fe_elem_post: NULL argument");
  FE_REQUIRE(kind == FE_ELAST_PSTRESS || kind == FE_ELAST_PSTRAIN || kind == FE_MAGNETIC,
             "fe_elem_post: kind %d has no post-processing", kind);
  if (n_elems <= 0) return FE_OK;
  cudaStream_t st = as_stream(stream);
  MatRow *tab = nullptr;
  int rc = build_material_table(ctx, st, kind, mat, n_mat, &tab, &ctx->scratch_b);
  if (rc) return rc;
  const double2 *xy = reinterpret_cast<const double2 *>(coords);
  if (kind == FE_MAGNETIC)
    k_elem_post<true><<<grid_for(n_elems, 128), 128, 0, st>>>(n_elems, xy, conn, mat_id, tab, u, out);
  else
    k_elem_post<false><<<grid_for(n_elems, 128), 128, 0, st>>>(n_elems, xy, conn, mat_id, tab, u, out);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

}  // extern "C"
