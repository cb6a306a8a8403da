// Numeric assembly: deterministic row-owner gather, no atomics.
//
// Replaces analysis.py:324-339 / :357-365 (the per-element Python loop that emits 36 COO
// triplets) and scipy's duplicate summation at analysis.py:661.  One thread owns one node,
// i.e. `dim` consecutive CSR rows whose values are contiguous in `vals`.  It walks the
// node's corners (elements incident to it, ascending element id), recomputes the `dim`
// rows of Ke that belong to this node from the three vertex coordinates (FP64 registers,
// elem.cuh) and adds the three dim x dim blocks into the slots recorded by the plan.
// Every slot is written by exactly one thread in a fixed order => bit-reproducible.
//
// HBM traffic per element (2 DOF/node, valence 7): vals 112 B written once; conn4 16 B;
// coords 8 B; corner records 24 B; row pointers ~4 B.  Re-reads of conn4/coords by the
// three owners of an element are L2 hits (they sit one grid line apart).
//
// Variants
//   1  k_assemble_global : accumulates straight into vals (global RMW, first-touch flags
//                          avoid the memset).  Works for any valence.
//   2  k_assemble_tile   : a CTA stages its nodes' rows in shared memory, transposed
//                          ([slot][thread], conflict-free accumulation), then streams the
//                          tile out with fully coalesced 128-bit-per-lane row writes.
#include "elem.cuh"
#include "plan.cuh"

namespace fe {

constexpr int kTile = 128;  // nodes (= threads) per CTA

struct CornerCtx {
  int v;        // local vertex of this node in the element
  int k[3];     // row positions of the element's three vertices
  bool first[3];
};

__device__ __forceinline__ CornerCtx decode(int2 rec) {
  CornerCtx c;
  c.v = rec.x & 3;
  const uint32_t y = (uint32_t)rec.y;
  c.k[0] = y & 255;
  c.k[1] = (y >> 8) & 255;
  c.k[2] = (y >> 16) & 255;
  c.first[0] = (y >> 24) & 1;
  c.first[1] = (y >> 25) & 1;
  c.first[2] = (y >> 26) & 1;
  return c;
}

// ---------------------------------------------------------------------------------------
// variant 1: global accumulation
// ---------------------------------------------------------------------------------------
template <int KC>  // 0 elasticity, 1 mass, 2 magnetic
__global__ void __launch_bounds__(kTile) k_assemble_global(int32_t n_owned, const int32_t *__restrict__ corner_ptr,
                                                          const int2 *__restrict__ corner_rec,
                                                          const int32_t *__restrict__ adj_ptr,
                                                          const int4 *__restrict__ conn4,
                                                          const double2 *__restrict__ coords,
                                                          const MatRow *__restrict__ tab, double *__restrict__ vals) {
  const int32_t n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_owned) return;
  const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
  const int32_t a0 = adj_ptr[n];
  const int deg = adj_ptr[n + 1] - a0;
  constexpr int DIM = (KC == 2) ? 1 : 2;
  double *row0 = vals + (int64_t)a0 * DIM * DIM;
  double *row1 = row0 + DIM * deg;
  for (int32_t c = c0; c < c1; ++c) {
    const int2 rec = corner_rec[c];
    const CornerCtx cc = decode(rec);
    const int4 cn = __ldg(conn4 + (rec.x >> 2));
    const TriGeom g = tri_geom(__ldg(coords + cn.x), __ldg(coords + cn.y), __ldg(coords + cn.z));
    const MatRow m = tab[cn.w];
    if (KC == 2) {
      double r[3];
      mag_row(g, m, cc.v, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double *s = row0 + cc.k[j];
        *s = cc.first[j] ? r[j] : (*s + r[j]);
      }
    } else {
      Blk2 r[3];
      if (KC == 0)
        elast_row_blocks(g, m, cc.v, r);
      else
        mass_row_blocks(g, m, cc.v, r);
#pragma unroll
      for (int j = 0; j < 3; ++j) {
        double2 *s0 = reinterpret_cast<double2 *>(row0 + 2 * cc.k[j]);
        double2 *s1 = reinterpret_cast<double2 *>(row1 + 2 * cc.k[j]);
        if (cc.first[j]) {
          *s0 = make_double2(r[j].k00, r[j].k01);
          *s1 = make_double2(r[j].k10, r[j].k11);
        } else {
          double2 u0 = *s0, u1 = *s1;
          *s0 = make_double2(u0.x + r[j].k00, u0.y + r[j].k01);
          *s1 = make_double2(u1.x + r[j].k10, u1.y + r[j].k11);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------
// variant 2: shared-memory staged tile
// ---------------------------------------------------------------------------------------
// smem: int32 a_tile[kTile + 1] (block row pointers of the tile) | double acc[slots][kTile + 1]
// slot s of node i lives at acc[s * (kTile + 1) + i]; s = d * DIM * deg + DIM * k + c is also
// the offset of the entry inside the node's contiguous segment of `vals`.
template <int KC>
__global__ void __launch_bounds__(kTile) k_assemble_tile(int32_t n_owned, const int32_t *__restrict__ corner_ptr,
                                                        const int2 *__restrict__ corner_rec,
                                                        const int32_t *__restrict__ adj_ptr,
                                                        const int4 *__restrict__ conn4,
                                                        const double2 *__restrict__ coords,
                                                        const MatRow *__restrict__ tab, double *__restrict__ vals) {
  constexpr int DIM = (KC == 2) ? 1 : 2;
  constexpr int LD = kTile + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *a_tile = reinterpret_cast<int32_t *>(smem_raw);
  double *acc = reinterpret_cast<double *>(smem_raw + ((kTile + 1) * sizeof(int32_t) + 15) / 16 * 16);

  const int tid = threadIdx.x;
  const int32_t n0 = blockIdx.x * kTile;
  const int32_t n = n0 + tid;
  const int n_in_tile = min(kTile, n_owned - n0);
  if (tid <= n_in_tile) a_tile[tid] = adj_ptr[n0 + tid];
  if (tid == 0 && n_in_tile == kTile) a_tile[kTile] = adj_ptr[n0 + kTile];
  __syncthreads();

  if (n < n_owned) {
    const int32_t c0 = corner_ptr[n], c1 = corner_ptr[n + 1];
    const int deg = a_tile[tid + 1] - a_tile[tid];
    double *my = acc + tid;
    const int r1 = DIM * deg;  // slot offset of the node's second row
    for (int32_t c = c0; c < c1; ++c) {
      const int2 rec = corner_rec[c];
      const CornerCtx cc = decode(rec);
      const int4 cn = __ldg(conn4 + (rec.x >> 2));
      const TriGeom g = tri_geom(__ldg(coords + cn.x), __ldg(coords + cn.y), __ldg(coords + cn.z));
      const MatRow m = tab[cn.w];
      if (KC == 2) {
        double r[3];
        mag_row(g, m, cc.v, r);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double *s = my + cc.k[j] * LD;
          *s = cc.first[j] ? r[j] : (*s + r[j]);
        }
      } else {
        Blk2 r[3];
        if (KC == 0)
          elast_row_blocks(g, m, cc.v, r);
        else
          mass_row_blocks(g, m, cc.v, r);
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          double *s00 = my + (2 * cc.k[j]) * LD;
          double *s01 = s00 + LD;
          double *s10 = my + (r1 + 2 * cc.k[j]) * LD;
          double *s11 = s10 + LD;
          if (cc.first[j]) {
            *s00 = r[j].k00;
            *s01 = r[j].k01;
            *s10 = r[j].k10;
            *s11 = r[j].k11;
          } else {
            *s00 += r[j].k00;
            *s01 += r[j].k01;
            *s10 += r[j].k10;
            *s11 += r[j].k11;
          }
        }
      }
    }
  }
  __syncthreads();

  // stream the tile out: one warp per node row segment, lanes over consecutive entries
  const int lane = tid & 31, w = tid >> 5;
  for (int i = w; i < n_in_tile; i += kTile / 32) {
    const int32_t a0 = a_tile[i];
    const int len = (a_tile[i + 1] - a0) * DIM * DIM;
    double *dst = vals + (int64_t)a0 * DIM * DIM;
    for (int q = lane; q < len; q += 32) dst[q] = acc[q * LD + i];
  }
}

// ---------------------------------------------------------------------------------------
// variant 3: fan-ordered traversal + shared-memory staged tile (the default)
// ---------------------------------------------------------------------------------------
// The plan orders each node's corners around the node (plan.cu: fan_walk), so a step shares
// its "previous" neighbour with the step before it.  Per step: ONE 8-byte record and ONE
// 16-byte coordinate load (the other two vertices are already in registers), no connectivity
// read; the block towards the previous neighbour is completed in registers (carry + this
// element) and stored once; the diagonal block never leaves registers until the end.  Shared
// memory therefore sees every value exactly once, as a 128-bit store, and the tile is then
// streamed to HBM as full contiguous rows.  Element geometry is evaluated in (self, prev,
// next) vertex order -- Ke is invariant under relabelling, rounding differs in the last ulp
// from the element-order kernels (tests: 1e-14 between variants).
struct FanFlags {
  static constexpr uint32_t SEED = 1, ADD_CARRY = 2, HOLD_A = 4, LAST = 8, ADD_FIRST = 16;
};

template <int KC>
__global__ void __launch_bounds__(kTile) k_assemble_fan(int32_t n_owned, const int32_t *__restrict__ fan_ptr,
                                                       const int2 *__restrict__ fan_rec,
                                                       const int32_t *__restrict__ adj_ptr,
                                                       const double2 *__restrict__ coords,
                                                       const MatRow *__restrict__ tab, double *__restrict__ vals) {
  constexpr int DIM = (KC == 2) ? 1 : 2;
  constexpr int LD = kTile + 1;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  int32_t *a_tile = reinterpret_cast<int32_t *>(smem_raw);
  unsigned char *acc_raw = smem_raw + ((kTile + 1) * sizeof(int32_t) + 15) / 16 * 16;

  const int tid = threadIdx.x;
  const int32_t n0 = blockIdx.x * kTile;
  const int32_t n = n0 + tid;
  const int n_in_tile = min(kTile, n_owned - n0);
  if (tid <= n_in_tile) a_tile[tid] = adj_ptr[n0 + tid];
  if (tid == 0 && n_in_tile == kTile) a_tile[kTile] = adj_ptr[n0 + kTile];
  __syncthreads();

  if (n < n_owned) {
    const int32_t f0 = fan_ptr[n], f1 = fan_ptr[n + 1];
    const int deg = a_tile[tid + 1] - a_tile[tid];
    const double2 ps = __ldg(coords + n);
    double2 pprev = ps;
    int kprev = 0, kself = 0, cur_mat = -1;
    MatRow m = {0.0, 0.0, 0.0, 0.0};
    if (KC == 2) {
      double *my = reinterpret_cast<double *>(acc_raw) + tid;
      double diag = 0.0, carry = 0.0, first = 0.0;
      for (int32_t f = f0; f < f1; ++f) {
        const int2 rec = ldg_nc_int2(fan_rec + f);
        const uint32_t y = (uint32_t)rec.y;
        const int k = y & 255;
        const uint32_t fl = (y >> 8) & 31;
        const double2 pc = __ldg(coords + rec.x);
        if (fl & FanFlags::SEED) {
          kself = y >> 13;
        } else {
          const int mid = y >> 13;
          if (mid != cur_mat) {
            m = tab[mid];
            cur_mat = mid;
          }
          const TriGeom g = tri_geom(ps, pprev, pc);
          double r[3];
          mag_row(g, m, 0, r);
          diag += r[0];
          double a = r[1];
          if (fl & FanFlags::ADD_CARRY) a += carry;
          if (fl & FanFlags::HOLD_A)
            first = a;
          else
            my[kprev * LD] = a;
          carry = r[2];
          if (fl & FanFlags::LAST) my[k * LD] = (fl & FanFlags::ADD_FIRST) ? carry + first : carry;
        }
        pprev = pc;
        kprev = k;
      }
      if (f1 > f0) my[kself * LD] = diag;
    } else {
      double2 *my = reinterpret_cast<double2 *>(acc_raw) + tid;
      Blk2 diag = {0.0, 0.0, 0.0, 0.0}, carry = diag, first = diag;
      for (int32_t f = f0; f < f1; ++f) {
        const int2 rec = ldg_nc_int2(fan_rec + f);
        const uint32_t y = (uint32_t)rec.y;
        const int k = y & 255;
        const uint32_t fl = (y >> 8) & 31;
        const double2 pc = __ldg(coords + rec.x);
        if (fl & FanFlags::SEED) {
          kself = y >> 13;
        } else {
          const int mid = y >> 13;
          if (mid != cur_mat) {
            m = tab[mid];
            cur_mat = mid;
          }
          const TriGeom g = tri_geom(ps, pprev, pc);
          Blk2 r[3];
          if (KC == 0)
            elast_row_blocks(g, m, 0, r);
          else
            mass_row_blocks(g, m, 0, r);
          diag.k00 += r[0].k00;
          diag.k01 += r[0].k01;
          diag.k10 += r[0].k10;
          diag.k11 += r[0].k11;
          Blk2 a = r[1];
          if (fl & FanFlags::ADD_CARRY) {
            a.k00 += carry.k00;
            a.k01 += carry.k01;
            a.k10 += carry.k10;
            a.k11 += carry.k11;
          }
          if (fl & FanFlags::HOLD_A) {
            first = a;
          } else {
            my[kprev * LD] = make_double2(a.k00, a.k01);
            my[(deg + kprev) * LD] = make_double2(a.k10, a.k11);
          }
          carry = r[2];
          if (fl & FanFlags::LAST) {
            Blk2 b = carry;
            if (fl & FanFlags::ADD_FIRST) {
              b.k00 += first.k00;
              b.k01 += first.k01;
              b.k10 += first.k10;
              b.k11 += first.k11;
            }
            my[k * LD] = make_double2(b.k00, b.k01);
            my[(deg + k) * LD] = make_double2(b.k10, b.k11);
          }
        }
        pprev = pc;
        kprev = k;
      }
      if (f1 > f0) {
        my[kself * LD] = make_double2(diag.k00, diag.k01);
        my[(deg + kself) * LD] = make_double2(diag.k10, diag.k11);
      }
    }
  }
  __syncthreads();

  const int lane = tid & 31, w = tid >> 5;
  if (KC == 2) {
    const double *acc = reinterpret_cast<const double *>(acc_raw);
    for (int i = w; i < n_in_tile; i += kTile / 32) {
      const int32_t a0 = a_tile[i];
      const int len = a_tile[i + 1] - a0;
      double *dst = vals + a0;
      for (int q = lane; q < len; q += 32) dst[q] = acc[q * LD + i];
    }
  } else {
    // half a warp per node: its 2 * valence double2 entries are contiguous in vals
    const double2 *acc = reinterpret_cast<const double2 *>(acc_raw);
    const int hl = lane & 15;
    for (int i = 2 * w + (lane >> 4); i < n_in_tile; i += 2 * (kTile / 32)) {
      const int32_t a0 = a_tile[i];
      const int len2 = 2 * (a_tile[i + 1] - a0);
      double2 *dst = reinterpret_cast<double2 *>(vals + 4 * (int64_t)a0);
      for (int q = hl; q < len2; q += 16) dst[q] = acc[q * LD + i];
    }
  }
}

static size_t tile_smem_bytes(int dim, int max_degree) {
  return ((kTile + 1) * sizeof(int32_t) + 15) / 16 * 16 + (size_t)dim * dim * max_degree * (kTile + 1) * sizeof(double);
}

}  // namespace fe

using namespace fe;

extern "C" int fe_assemble(fe_ctx *ctx, void *stream, const fe_plan *p, int kind, const double *coords,
                           const double *mat, int32_t n_mat, double *vals, int variant) {
  FE_REQUIRE(ctx && p && coords && (vals || p->nnz == 0), "fe_assemble: NULL argument");
  FE_REQUIRE(kind >= FE_ELAST_PSTRESS && kind <= FE_MASS, "fe_assemble: unknown kind %d", kind);
  const int dim = (kind == FE_MAGNETIC) ? 1 : 2;
  FE_REQUIRE(dim == p->dim, "fe_assemble: kind %d needs dim %d but the plan was built with dim %d", kind, dim, p->dim);
  if (p->n_owned == 0 || p->nnz == 0) return FE_OK;
  cudaStream_t st = as_stream(stream);
  MatRow *tab = nullptr;
  int rc = build_material_table(ctx, st, kind, mat, n_mat, &tab, &ctx->scratch_b);
  if (rc) return rc;
  const double2 *xy = reinterpret_cast<const double2 *>(coords);
  const int grid = grid_for(p->n_owned, kTile);
  const size_t smem = tile_smem_bytes(dim, p->max_degree);
  const size_t smem_limit = 200 * 1024;
  if (variant == 0) variant = (smem <= smem_limit) ? (p->fan_ok ? 3 : 2) : 1;
  if (variant == 3 && !p->fan_ok)
    return fail(FE_ERR_UNSUPPORTED, "fe_assemble: the fan variant needs a mesh whose node stars are simple fans");
  if (variant >= 2 && smem > smem_limit)
    return fail(FE_ERR_UNSUPPORTED, "fe_assemble: tile variant needs %zu B of shared memory (valence %d)", smem,
                p->max_degree);
  FE_REQUIRE(variant >= 1 && variant <= 3, "fe_assemble: unknown variant %d", variant);

#define FE_ASM_LAUNCH(KC)                                                                                       \
  do {                                                                                                          \
    if (variant == 1) {                                                                                         \
      k_assemble_global<KC><<<grid, kTile, 0, st>>>(p->n_owned, p->corner_ptr, p->corner_rec, p->adj_ptr,        \
                                                    p->conn4, xy, tab, vals);                                   \
    } else if (variant == 3) {                                                                                  \
      FE_CUDA(cudaFuncSetAttribute(k_assemble_fan<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_assemble_fan<KC><<<grid, kTile, smem, st>>>(p->n_owned, p->fan_ptr, p->fan_rec, p->adj_ptr, xy, tab,     \
                                                    vals);                                                      \
    } else {                                                                                                    \
      FE_CUDA(cudaFuncSetAttribute(k_assemble_tile<KC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      k_assemble_tile<KC><<<grid, kTile, smem, st>>>(p->n_owned, p->corner_ptr, p->corner_rec, p->adj_ptr,       \
                                                     p->conn4, xy, tab, vals);                                  \
    }                                                                                                           \
  } while (0)

  if (kind == FE_MAGNETIC)
    FE_ASM_LAUNCH(2);
  else if (kind == FE_MASS)
    FE_ASM_LAUNCH(1);
  else
    FE_ASM_LAUNCH(0);
#undef FE_ASM_LAUNCH
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}
