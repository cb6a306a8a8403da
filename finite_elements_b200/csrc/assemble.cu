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
#include "ptx.cuh"

namespace fe {


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
  extern __shared__ __align__(128) unsigned char smem_raw[];
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
// its "previous" neighbour with the step before it.
//  * The tile's fan records are one contiguous range of fan_rec: the CTA copies it to shared
//    memory with coalesced 128-bit streaming loads (each record leaves HBM exactly once).
//  * Per step: one LDS.64 record and ONE 16-byte coordinate gather, issued two steps ahead
//    (software pipeline) -- the other two vertices are already in registers, no connectivity.
//  * The block towards the previous neighbour is completed in registers (carry + this
//    element) and stored once; the diagonal block stays in registers until the end.  Shared
//    memory sees every value exactly once, as a 128-bit store into the exact image of the
//    tile's slice of `vals`, which one thread then hands to the TMA engine as a single bulk
//    store (cp.async.bulk.global.shared::cta) -- no per-lane write-out loop at all.
// Element geometry is evaluated in (self, prev, next) vertex order -- Ke is invariant under
// relabelling, rounding differs in the last ulp from the element-order kernels (tests: 1e-14
// between variants).
struct FanFlags {
  static constexpr uint32_t SEED = 1, MULTI = 2, HOLD_A = 4, LAST = 8, ADD_FIRST = 16;
};

// 1 / d to ~1 ulp without the slow-path branch of the compiler's division: MUFU.RCP64H seed (>= 20 bits)
// and two Newton steps.  |d| is an element's 2 x area: never subnormal or huge on a mesh that has a K.
__device__ __forceinline__ double fan_rcp(double d) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(d));
  double e = fma(-d, y, 1.0);
  y = fma(y, e, y);
  e = fma(-d, y, 1.0);
  return fma(y, e, y);
}

// Per-kind arithmetic of one fan step.  With e1 = prev - self and e2 = cur - self (the element is (self, prev,
// cur)) the reference's coefficients (elements.py:403-408) are
//   beta_self = e1.y - e2.y, gamma_self = e2.x - e1.x;  beta_prev = e2.y, gamma_prev = -e2.x;
//   beta_cur = -e1.y, gamma_cur = e1.x;  det = e1.x e2.y - e1.y e2.x  (= 2 A, signed)
// and area / det^2 = 1 / (2 |det|).  step() ADDS the element's block towards `prev` to pb (which arrives holding
// the previous element's share, the carry) and returns the block towards `cur` in cb.  The diagonal block is not
// evaluated: every block row of Ke sums to zero (rigid translation / constant potential), so K_ii = -(sum of
// the row's off-diagonal blocks); for the consistent mass matrix M_ii = +(that sum).
template <int KC>
struct FanOps;

template <>
struct FanOps<2> {  // magnetic: scalar entries, Ke_ij = (1/mu) (beta_i beta_j + gamma_i gamma_j) / (2 |det|)
  using Val = double;
  using Slot = double;
  static __device__ __forceinline__ Val zero() { return 0.0; }
  static __device__ __forceinline__ void step(const double2 &e1, const double2 &e2, const MatRow &m, Val &pb, Val &cb) {
    const double b0 = e1.y - e2.y, g0 = e2.x - e1.x;
    const double det = e1.x * e2.y - e1.y * e2.x;
    const double s = (0.5 * m.p0) * fabs(fan_rcp(det));
    const double sb = s * b0, sg = s * g0;
    pb = fma(sb, e2.y, pb);
    pb = fma(-sg, e2.x, pb);
    cb = sg * e1.x - sb * e1.y;
  }
  static __device__ __forceinline__ void add(Val &a, const Val &b) { a += b; }
  static __device__ __forceinline__ void diag_acc(Val &d, const Val &b) { d -= b; }
  static __device__ __forceinline__ void store(Slot *my, int /*deg*/, int k, const Val &v) { my[k] = v; }
  static __device__ __forceinline__ Val load(const Slot *my, int /*deg*/, int k) { return my[k]; }
};

template <int KC>
struct FanOps {  // elasticity (0) / mass (1): 2x2 blocks, stored as two double2 (one per row)
  using Val = Blk2;
  using Slot = double2;
  static __device__ __forceinline__ Val zero() { return Blk2{0.0, 0.0, 0.0, 0.0}; }
  static __device__ __forceinline__ void step(const double2 &e1, const double2 &e2, const MatRow &m, Val &pb, Val &cb) {
    const double det = e1.x * e2.y - e1.y * e2.x;
    if (KC == 1) {  // rho t / 12 * area on the off-diagonal blocks' diagonals
      const double s = (0.5 * m.p0) * fabs(det);
      pb.k00 += s;
      pb.k11 += s;
      cb = Blk2{s, 0.0, 0.0, s};
      return;
    }
    // Ke(self, j) = t A B_self^T D B_j (elements.py:466-511), t in the material row
    const double b0 = e1.y - e2.y, g0 = e2.x - e1.x;
    const double s = 0.5 * fabs(fan_rcp(det));
    const double tb = s * b0, tg = s * g0;
    const double cb_ = m.p0 * tb, cg = m.p0 * tg;  // c t
    const double ab = m.p1 * tb, ag = m.p1 * tg;   // a t
    const double sb = m.p2 * tb, sg = m.p2 * tg;   // b t (shear)
    // j = prev: beta = e2.y, gamma = -e2.x
    pb.k00 = fma(cb_, e2.y, pb.k00);
    pb.k00 = fma(-sg, e2.x, pb.k00);
    pb.k01 = fma(sg, e2.y, pb.k01);
    pb.k01 = fma(-ab, e2.x, pb.k01);
    pb.k10 = fma(ag, e2.y, pb.k10);
    pb.k10 = fma(-sb, e2.x, pb.k10);
    pb.k11 = fma(sb, e2.y, pb.k11);
    pb.k11 = fma(-cg, e2.x, pb.k11);
    // j = cur: beta = -e1.y, gamma = e1.x
    cb.k00 = sg * e1.x - cb_ * e1.y;
    cb.k01 = ab * e1.x - sg * e1.y;
    cb.k10 = sb * e1.x - ag * e1.y;
    cb.k11 = cg * e1.x - sb * e1.y;
  }
  static __device__ __forceinline__ void add(Val &a, const Val &b) {
    a.k00 += b.k00;
    a.k01 += b.k01;
    a.k10 += b.k10;
    a.k11 += b.k11;
  }
  static __device__ __forceinline__ void diag_acc(Val &d, const Val &b) {
    if (KC == 1) {
      d.k00 += b.k00;
      d.k11 += b.k11;
    } else {
      d.k00 -= b.k00;
      d.k01 -= b.k01;
      d.k10 -= b.k10;
      d.k11 -= b.k11;
    }
  }
  static __device__ __forceinline__ void store(Slot *my, int deg, int k, const Val &v) {
    my[k] = make_double2(v.k00, v.k01);
    my[deg + k] = make_double2(v.k10, v.k11);
  }
  static __device__ __forceinline__ Val load(const Slot *my, int deg, int k) {
    const double2 u = my[k], w = my[deg + k];
    return Blk2{u.x, u.y, w.x, w.y};
  }
};

constexpr int kFanChunk = 32;                           // nodes per chunk = one warp
constexpr int kFanPtrInts = (kFanChunk + 1 + 3) & ~3;  // 36: a pointer slice of a chunk, padded to 16 bytes

// Record format of the fan walk: the plan's 8-byte records, or their 4-byte form (plan.cuh) with the
// per-node header word in a third pointer slice of the ring stage.
template <bool R4>
struct FanRec;
template <>
struct FanRec<false> {
  using T = int2;
  static constexpr int kAlign = 2;  // records per 16 bytes
  static __device__ __forceinline__ int32_t nbr(T r, int32_t /*self*/, int32_t /*n_owned*/) { return r.x; }
  static __device__ __forceinline__ uint32_t k(T r) { return (uint32_t)r.y & 255; }
  static __device__ __forceinline__ bool seed(T r) { return (uint32_t)r.y & (FanFlags::SEED << 8); }
  static __device__ __forceinline__ bool last(T r) { return (uint32_t)r.y & (FanFlags::LAST << 8); }
  static __device__ __forceinline__ bool add_first(T r) { return (uint32_t)r.y & (FanFlags::ADD_FIRST << 8); }
  static __device__ __forceinline__ int kself(T r, uint32_t /*hdr*/) { return (uint32_t)r.y >> 13; }
  static __device__ __forceinline__ bool multi(T r) { return (uint32_t)r.y & (FanFlags::MULTI << 8); }
  static __device__ __forceinline__ int first_mat(uint32_t /*hdr*/) { return -1; }
  static __device__ __forceinline__ bool new_mat(T r, uint32_t /*hdr*/, int &cur) {
    const int mid = (uint32_t)r.y >> 13;
    const bool ch = mid != cur;
    cur = mid;
    return ch;
  }
};
template <>
struct FanRec<true> {
  using T = uint32_t;
  static constexpr int kAlign = 4;
  static __device__ __forceinline__ int32_t nbr(T r, int32_t self, int32_t n_owned) {
    return self + ((int32_t)r >> kFan4Shift);
  }
  static __device__ __forceinline__ uint32_t k(T r) { return r & 255; }
  static __device__ __forceinline__ bool seed(T r) { return r & (FAN4_SEED << 8); }
  static __device__ __forceinline__ bool last(T r) { return r & (FAN4_LAST << 8); }
  static __device__ __forceinline__ bool add_first(T r) { return r & (FAN4_ADD_FIRST << 8); }
  static __device__ __forceinline__ int kself(T /*r*/, uint32_t hdr) { return hdr & 255; }
  static __device__ __forceinline__ bool multi(T r) { return r & (FAN4_MULTI << 8); }
  static __device__ __forceinline__ int first_mat(uint32_t hdr) { return (hdr >> 8) & 4095; }
  static __device__ __forceinline__ bool new_mat(T r, uint32_t hdr, int &cur) {
    if (!(r & (FAN4_MATSW << 8))) return false;
    const int m0 = (hdr >> 8) & 4095, m1 = hdr >> 20;
    cur = (cur == m0) ? m1 : m0;
    return true;
  }
};

#include "fan_kernel.cuh"

static int fan_warp_slot_bytes(int dim, int max_degree) { return dim * dim * max_degree * kFanChunk * 8; }
static int fan_rec_cap(int fan_tile_max, bool r4) { return r4 ? ((fan_tile_max + 7) & ~3) : ((fan_tile_max + 3) & ~1); }
static size_t fan_smem_bytes(int dim, int max_degree, int fan_tile_max, bool r4) {
  return kFanWarps * fan_warp_bytes(fan_rec_cap(fan_tile_max, r4), fan_warp_slot_bytes(dim, max_degree), r4);
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
  FE_REQUIRE(mat && n_mat > p->max_mat_id, "fe_assemble: the mesh refers to material %d but the table has %d row(s)",
             p->max_mat_id, n_mat);
  if (p->n_owned == 0 || p->nnz == 0) return FE_OK;
  cudaStream_t st = as_stream(stream);
  MatRow *tab = nullptr;
  int rc = build_material_table(ctx, st, kind, mat, n_mat, &tab, &ctx->scratch_b);
  if (rc) return rc;
  const double2 *xy = reinterpret_cast<const double2 *>(coords);
  const int grid = grid_for(p->n_owned, kTile);
  // variant 3 = fan walk (4-byte records when the plan could build them), 4 = fan walk on the 8-byte records
  bool r4 = p->fan_compact_ok && variant != 4;
  size_t smem = tile_smem_bytes(dim, p->max_degree);
  const size_t smem_limit = 200 * 1024;
  if (r4 && fan_smem_bytes(dim, p->max_degree, p->fan_tile_max, true) > smem_limit) r4 = false;
  const size_t smem_fan = fan_smem_bytes(dim, p->max_degree, p->fan_tile_max, r4);
  const int rec_cap = fan_rec_cap(p->fan_tile_max, r4);
  if (variant == 0) variant = (p->fan_ok && smem_fan <= smem_limit) ? 3 : ((smem <= smem_limit) ? 2 : 1);
  if (variant == 4) variant = 3;
  if (variant == 3) smem = smem_fan;
  if (variant == 3 && !p->fan_ok)
    return fail(FE_ERR_UNSUPPORTED, "fe_assemble: the fan variant needs a mesh whose node stars are simple fans");
  if (variant >= 2 && smem > smem_limit)
    return fail(FE_ERR_UNSUPPORTED, "fe_assemble: tile variant needs %zu B of shared memory (valence %d)", smem,
                p->max_degree);
  FE_REQUIRE(variant >= 1 && variant <= 3, "fe_assemble: unknown variant %d", variant);

#define FE_FAN_LAUNCH(KC, R4, RECS)                                                                              \
  do {                                                                                                          \
    /* persistent grid: as many CTAs as are resident.  The attribute / occupancy calls cost the CPU ~20 us:    */ \
    /* remembered per instance and shared-memory size                                                          */ \
    static thread_local size_t cached_smem = ~(size_t)0;                                                        \
    static thread_local int cached_minb = 1, cached_dev = -1;                                                   \
    if (cached_smem != smem || cached_dev != ctx->device) {                                                     \
      int minb = 1;                                                                                             \
      FE_CUDA(cudaFuncSetAttribute(k_assemble_fan<KC, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      FE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&minb, k_assemble_fan<KC, R4>, kFanThreads, smem)); \
      cached_minb = minb < 1 ? 1 : minb;                                                                        \
      cached_smem = smem;                                                                                       \
      cached_dev = ctx->device;                                                                                 \
    }                                                                                                           \
    const int fgrid = grid < cached_minb * ctx->num_sms ? grid : cached_minb * ctx->num_sms;                    \
    k_assemble_fan<KC, R4><<<fgrid, kFanThreads, smem, st>>>(p->n_owned, p->fan_ptr, RECS, p->fan_hdr,          \
                                                             p->adj_ptr, xy, tab, vals, rec_cap,                \
                                                             fan_warp_slot_bytes(dim, p->max_degree));          \
  } while (0)
#define FE_ASM_LAUNCH(KC)                                                                                       \
  do {                                                                                                          \
    if (variant == 1) {                                                                                         \
      k_assemble_global<KC><<<grid, kTile, 0, st>>>(p->n_owned, p->corner_ptr, p->corner_rec, p->adj_ptr,        \
                                                    p->conn4, xy, tab, vals);                                   \
    } else if (variant == 3) {                                                                                  \
      if (r4)                                                                                                   \
        FE_FAN_LAUNCH(KC, true, p->fan_rec4);                                                                   \
      else                                                                                                      \
        FE_FAN_LAUNCH(KC, false, p->fan_rec);                                                                   \
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
#undef FE_FAN_LAUNCH
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}
