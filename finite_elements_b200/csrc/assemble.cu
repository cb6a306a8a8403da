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
#ifndef FE_FAN_DEFAULT_DESIGN
#define FE_FAN_DEFAULT_DESIGN 0  // 0 = register walk (fan_regwalk.cuh), 1 = all-asynchronous (k_assemble_fan)
#endif
#ifndef FE_FAN_MINB
#define FE_FAN_MINB 5  // resident CTAs (of 2 warps) per SM the register allocation targets; shared memory allows 5 at valence 7
#endif

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

// Persistent kernel in which every WARP is an independent software pipeline over 32-node
// chunks (chunk = global warp id, + total warps, ...); warps never synchronise with each other and
// the walk touches nothing but shared memory and registers.
//  * Input ring per warp (kFanStages = 3 stages, one mbarrier each), filled by the TMA engine: the pointer
//    slices of a chunk (adj_ptr, fan_ptr, fan_hdr: 36 words each), its own coordinates and its contiguous
//    record range.  A stage is refilled as soon as its chunk is done; the end points of the record range
//    a refill needs are fetched one chunk earlier with cp.async (global -> shared: no load is in flight
//    into a register across the loop's back edge -- ptxas waits for those at the branch).
//  * Neighbour coordinates: at the top of chunk c every lane walks the records of ITS node of chunk
//    c+1 (already in the ring) and issues one 16-byte cp.async per record into one of two coordinate
//    arrays -- a whole chunk ahead of their use, tracked by cp.async groups, not by register scoreboards
//    (register prefetch two steps ahead: 17 % of the stall samples on the first gathers of a chunk and
//    false scoreboard dependencies between the prefetch sets, ncu r02 captures L and N).
//  * Walk: the first record of a node is the seed of its fan; a node with a single fan (all but
//    boundary corners and bow-ties) runs a loop without flag tests, two steps per trip with the
//    carry / edge registers swapping roles; other nodes take the general loop.
//  * Output: each warp owns a private sub-tile, the exact image of its 32 nodes' slice of `vals`,
//    and copies it out itself with coalesced 128-bit stores (a TMA bulk store kept the sub-tile busy
//    until the store engine had drained it: 9 % of the samples).
// smem per warp: full[3] | end points int[2][2] | 3 x { a_slice[36], f_slice[36], (hdr[36]), self_xy[32], recs[rec_cap] }
//                | 2 x xy[rec_cap] | sub-tile
#ifndef FE_FAN_WARPS
#define FE_FAN_WARPS 2
#endif
constexpr int kFanWarps = FE_FAN_WARPS;  // independent warps per CTA
constexpr int kFanThreads = kFanWarps * 32;
constexpr int kFanChunk = 32;
constexpr int kFanPtrInts = (kFanChunk + 1 + 3) & ~3;  // 36
constexpr int kFanStages = 3;
constexpr int kFanHdrBytes = 64;  // barriers + end points

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

__host__ __device__ inline size_t fan_stage_bytes(int rec_cap, bool r4) {
  // pointer slices | own coordinates | records (16-byte aligned pieces)
  return (size_t)(r4 ? 3 : 2) * kFanPtrInts * sizeof(int32_t) + kFanChunk * 16 +
         ((size_t)rec_cap * (r4 ? 4 : 8) + 15) / 16 * 16;
}
__host__ __device__ inline size_t fan_warp_bytes(int rec_cap, int warp_slot_bytes, bool r4) {
  return (kFanHdrBytes + kFanStages * fan_stage_bytes(rec_cap, r4) + 2 * (size_t)rec_cap * 16 +
          (size_t)warp_slot_bytes + 127) / 128 * 128;
}

template <int KC, bool R4>
__global__ void __launch_bounds__(kFanThreads, FE_FAN_MINB) k_assemble_fan(
    int32_t n_owned, const int32_t *__restrict__ fan_ptr, const typename FanRec<R4>::T *__restrict__ fan_rec,
    const uint32_t *__restrict__ fan_hdr, const int32_t *__restrict__ adj_ptr, const double2 *__restrict__ coords,
    const MatRow *__restrict__ tab, double *__restrict__ vals, int rec_cap, int warp_slot_bytes) {
  using Ops = FanOps<KC>;
  using Val = typename Ops::Val;
  using Slot = typename Ops::Slot;
  using RO = FanRec<R4>;
  using Rec = typename RO::T;
  constexpr int SPB = (KC == 2) ? 1 : 2;  // Slots per node-level block
  constexpr int kPtrSlices = R4 ? 3 : 2;  // adj_ptr, fan_ptr (, fan_hdr)
  constexpr int kSelfOff = kPtrSlices * kFanPtrInts * 4, kRecOff = kSelfOff + kFanChunk * 16;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char *wbase = smem_raw + (size_t)warp * fan_warp_bytes(rec_cap, warp_slot_bytes, R4);
  uint64_t *full = reinterpret_cast<uint64_t *>(wbase);
  int32_t *ep = reinterpret_cast<int32_t *>(wbase + 32);  // [2][2] record-range end points (LDGSTS)
  const size_t stage_bytes = fan_stage_bytes(rec_cap, R4);
  unsigned char *stage0 = wbase + kFanHdrBytes;
  double2 *xy0 = reinterpret_cast<double2 *>(stage0 + kFanStages * stage_bytes);
  Slot *acc = reinterpret_cast<Slot *>(xy0 + 2 * rec_cap);

  const int n_chunks = (n_owned + kFanChunk - 1) / kFanChunk;
  const int chunk_stride = gridDim.x * kFanWarps;
  if (lane == 0) {
#pragma unroll
    for (int q = 0; q < kFanStages; ++q) ptx::mbar_init(&full[q], 1);
    ptx::mbar_init_fence();
  }
  __syncwarp();

  // ---- lane 0: the TMA loads of a chunk, given the end points [r0, r1) of its record range
  auto issue = [&](int chunk, int stage, int32_t r0, int32_t r1) {
    const int32_t n0 = chunk * kFanChunk;
    const int32_t base = r0 & ~(RO::kAlign - 1);  // 16-byte aligned start of the record copy
    const uint32_t rec_bytes = (uint32_t)((r1 - base + RO::kAlign - 1) / RO::kAlign) * 16u;
    const uint32_t self_bytes = (uint32_t)min(kFanChunk, n_owned - n0) * 16u;
    unsigned char *st = stage0 + stage * stage_bytes;
    ptx::mbar_expect_tx(&full[stage], (uint32_t)kPtrSlices * kFanPtrInts * 4u + self_bytes + rec_bytes);
    ptx::bulk_load(st, adj_ptr + n0, kFanPtrInts * 4u, &full[stage]);
    ptx::bulk_load(st + kFanPtrInts * 4, fan_ptr + n0, kFanPtrInts * 4u, &full[stage]);
    if (R4) ptx::bulk_load(st + 2 * kFanPtrInts * 4, fan_hdr + n0, kFanPtrInts * 4u, &full[stage]);
    ptx::bulk_load(st + kSelfOff, coords + n0, self_bytes, &full[stage]);
    if (rec_bytes) ptx::bulk_load(st + kRecOff, fan_rec + base, rec_bytes, &full[stage]);
  };
  auto request_endpoints = [&](int chunk, int slot) {
    if (chunk < n_chunks) {
      const int32_t n0 = chunk * kFanChunk;
      ptx::cp_async4(ep + 2 * slot, fan_ptr + n0);
      ptx::cp_async4(ep + 2 * slot + 1, fan_ptr + min(n0 + kFanChunk, n_owned));
    }
    ptx::cp_async_commit();  // a group of its own, OLDER than the gathers committed at the top of the next trip:
                             // that trip's wait_group<1> retires it before the end points are read
  };
  // ---- every lane: record range of its node in a (full) ring stage; the gathers of that node
  auto lane_range = [&](int chunk, int stage, int &f0, int &f1) {
    const int32_t *f_sl = reinterpret_cast<const int32_t *>(stage0 + stage * stage_bytes) + kFanPtrInts;
    f0 = f1 = 0;
    if (lane < min(kFanChunk, n_owned - chunk * kFanChunk)) {
      const int32_t base = f_sl[0] & ~(RO::kAlign - 1);
      f0 = f_sl[lane] - base;
      f1 = f_sl[lane + 1] - base;
    }
  };
  auto gather = [&](int chunk, int stage, double2 *xy) {
    int f0, f1;
    lane_range(chunk, stage, f0, f1);
    const Rec *rc = reinterpret_cast<const Rec *>(stage0 + stage * stage_bytes + kRecOff);
    const int32_t self = chunk * kFanChunk + lane;
    for (int i = f0; i < f1; ++i) ptx::cp_async16(xy + i, coords + RO::nbr(rc[i], self, n_owned));
  };

  int chunk = blockIdx.x * kFanWarps + warp;
  if (chunk < n_chunks) {
    if (lane == 0) {
      // the first kFanStages chunks of this warp: direct loads of the end points (start-up only)
      for (int q = 0; q < kFanStages; ++q) {
        const int c = chunk + q * chunk_stride;
        if (c < n_chunks) {
          const int32_t n0 = c * kFanChunk;
          issue(c, q, __ldg(fan_ptr + n0), __ldg(fan_ptr + min(n0 + kFanChunk, n_owned)));
        }
      }
      request_endpoints(chunk + kFanStages * chunk_stride, 0);
    }
    ptx::mbar_wait(&full[0], 0);
    gather(chunk, 0, xy0);
    ptx::cp_async_commit();  // group of chunk 0's gathers
  }

  // ring position j in [0, 6): stage = j % 3, barrier parity = j / 3, coordinate array = j & 1
  for (int j = 0; chunk < n_chunks; chunk += chunk_stride, j = (j == 5) ? 0 : j + 1) {
    const int stage = (j >= kFanStages) ? j - kFanStages : j;
    const int next = chunk + chunk_stride;
    // ---- the next chunk: its ring stage was refilled two chunks ago; put its gathers in flight
    if (next < n_chunks) {
      const int jn = (j == 5) ? 0 : j + 1;
      const int sn = (jn >= kFanStages) ? jn - kFanStages : jn;
      ptx::mbar_wait(&full[sn], (uint32_t)(jn >= kFanStages));
      gather(next, sn, xy0 + ((j & 1) ^ 1) * rec_cap);
    }
    ptx::cp_async_commit();  // one group of gathers per chunk

    // ---- this thread's node
    const int32_t n0 = chunk * kFanChunk;
    const int n_in = min(kFanChunk, n_owned - n0);
    const unsigned char *st = stage0 + stage * stage_bytes;
    const int32_t *a_sl = reinterpret_cast<const int32_t *>(st);
    const Rec *recs = reinterpret_cast<const Rec *>(st + kRecOff);
    const double2 *xy = xy0 + (j & 1) * rec_cap;
    const int32_t out_lo = a_sl[0];
    const int32_t out_len = a_sl[n_in] - out_lo;  // node-level block range of this chunk
    int f, fe, deg = 0;
    lane_range(chunk, stage, f, fe);
    uint32_t hdr = 0;
    double2 ps = make_double2(0.0, 0.0);
    Slot *my = acc;
    if (lane < n_in) {
      deg = a_sl[lane + 1] - a_sl[lane];
      my = acc + SPB * (a_sl[lane] - out_lo);
      ps = reinterpret_cast<const double2 *>(st + kSelfOff)[lane];
      if (R4) hdr = reinterpret_cast<const uint32_t *>(a_sl + 2 * kFanPtrInts)[lane];
    }
    ptx::cp_async_wait_group<1>();  // everything but the group just committed: this chunk's gathers have landed

    // ---- the fan walk
    if (f < fe) {
      int cur_mat = RO::first_mat(hdr);
      MatRow m = {0.0, 0.0, 0.0, 0.0};
      if (R4) m = tab[cur_mat];
      Val diag = Ops::zero();
      Rec rp = recs[f];  // the seed of the node's first fan
      int kself = RO::kself(rp, hdr);
      const double2 p0 = xy[f];
      double2 ea = make_double2(p0.x - ps.x, p0.y - ps.y);
      Val X = Ops::zero(), Y;
      if (!RO::multi(rp)) {
        // single fan: records f+1 .. fe-1 are its steps, the last one closes or ends it
        int i = f + 1;
        for (; i + 1 < fe; i += 2) {
          const Rec r1 = recs[i], r2 = recs[i + 1];
          const double2 p1 = xy[i], p2 = xy[i + 1];
          const double2 e1 = make_double2(p1.x - ps.x, p1.y - ps.y), e2 = make_double2(p2.x - ps.x, p2.y - ps.y);
          if (RO::new_mat(r1, hdr, cur_mat)) m = tab[cur_mat];
          Ops::step(ea, e1, m, X, Y);  // X: the finished block towards the previous neighbour
          Ops::store(my, deg, RO::k(rp), X);
          Ops::diag_acc(diag, X);
          if (RO::new_mat(r2, hdr, cur_mat)) m = tab[cur_mat];
          Ops::step(e1, e2, m, Y, X);
          Ops::store(my, deg, RO::k(r1), Y);
          Ops::diag_acc(diag, Y);
          rp = r2;
          ea = e2;
        }
        if (i < fe) {
          const Rec r1 = recs[i];
          const double2 p1 = xy[i];
          const double2 e1 = make_double2(p1.x - ps.x, p1.y - ps.y);
          if (RO::new_mat(r1, hdr, cur_mat)) m = tab[cur_mat];
          Ops::step(ea, e1, m, X, Y);
          Ops::store(my, deg, RO::k(rp), X);
          Ops::diag_acc(diag, X);
          X = Y;
          rp = r1;
        }
        // X: the last element's block towards the last neighbour (a closed fan's first block waits in its slot)
        if (fe - f > 1) {
          Ops::diag_acc(diag, X);
          if (RO::add_first(rp)) Ops::add(X, Ops::load(my, deg, RO::k(rp)));
          Ops::store(my, deg, RO::k(rp), X);
        }
      } else {
        // general walk: several fans around the node (boundary corners, bow-ties) or 8-byte records
        for (int i = f + 1; i < fe; ++i) {
          const Rec rc = recs[i];
          const double2 p = xy[i];
          const double2 e2 = make_double2(p.x - ps.x, p.y - ps.y);
          if (RO::seed(rc)) {  // a chain starts: its first neighbour, nothing carried
            kself = RO::kself(rc, hdr);
            X = Ops::zero();
          } else {
            if (RO::new_mat(rc, hdr, cur_mat)) m = tab[cur_mat];
            Ops::step(ea, e2, m, X, Y);
            Ops::store(my, deg, RO::k(rp), X);
            Ops::diag_acc(diag, X);
            X = Y;
            if (RO::last(rc)) {
              Ops::diag_acc(diag, Y);
              if (RO::add_first(rc)) Ops::add(Y, Ops::load(my, deg, RO::k(rc)));
              Ops::store(my, deg, RO::k(rc), Y);
            }
          }
          ea = e2;
          rp = rc;
        }
      }
      Ops::store(my, deg, kself, diag);
    }
    __syncwarp();

    // ---- the sub-tile is complete, the exact image of vals[dim^2 * out_lo ...): coalesced copy-out
    if (KC == 2) {
      const double *src = reinterpret_cast<const double *>(acc);  // 1 DOF per node: only 8-byte aligned
      double *dst = vals + out_lo;
      for (int q = lane; q < out_len; q += 32) dst[q] = src[q];
    } else {
      const double2 *src = reinterpret_cast<const double2 *>(acc);
      double2 *dst = reinterpret_cast<double2 *>(vals + 4 * (int64_t)out_lo);
      for (int q = lane; q < 2 * out_len; q += 32) dst[q] = src[q];
    }
    if (lane == 0) {
      // this warp is done with ring slot `stage`: refill it with the chunk kFanStages ahead (its end points were
      // requested a whole chunk ago, in ep[j & 1], and belong to a group the wait above has retired)
      const int nn = chunk + kFanStages * chunk_stride;
      if (nn < n_chunks) issue(nn, stage, ep[2 * (j & 1)], ep[2 * (j & 1) + 1]);
      request_endpoints(nn + chunk_stride, (j & 1) ^ 1);
    }
    __syncwarp();
  }
}

#include "fan_regwalk.cuh"

static int fan_warp_slot_bytes(int dim, int max_degree) { return dim * dim * max_degree * kFanChunk * 8; }
static int fan_rec_cap(int fan_tile_max, bool r4) {  // alignment slack + round-up of the 16-byte copy
  return r4 ? ((fan_tile_max + 7) & ~3) : ((fan_tile_max + 3) & ~1);
}
static size_t fan_smem_bytes(int dim, int max_degree, int fan_tile_max, bool r4) {
  return kFanWarps * fan_warp_bytes(fan_rec_cap(fan_tile_max, r4), fan_warp_slot_bytes(dim, max_degree), r4);
}

static int rw_rec_cap(int fan_tile_max, bool r4) { return r4 ? ((fan_tile_max + 7) & ~3) : ((fan_tile_max + 3) & ~1); }
static size_t rw_smem_bytes(int dim, int max_degree, int fan_tile_max, bool r4) {
  return kRwWarps * rw_warp_bytes(rw_rec_cap(fan_tile_max, r4), fan_warp_slot_bytes(dim, max_degree), r4);
}
// FE_B200_FAN_DESIGN=rw selects the register-walk kernel (fan_regwalk.cuh), =async the all-asynchronous one
static int fan_design() {
  static int d = -1;
  if (d < 0) {
    const char *e = getenv("FE_B200_FAN_DESIGN");
    d = (e && e[0] == 'a') ? 1 : ((e && e[0] == 'r') ? 0 : FE_FAN_DEFAULT_DESIGN);
  }
  return d;
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
  const bool rw = fan_design() == 0;
  size_t smem = tile_smem_bytes(dim, p->max_degree);
  const size_t smem_limit = 200 * 1024;
  auto fan_bytes = [&](bool four) {
    return rw ? rw_smem_bytes(dim, p->max_degree, p->fan_tile_max, four) : fan_smem_bytes(dim, p->max_degree, p->fan_tile_max, four);
  };
  if (r4 && fan_bytes(true) > smem_limit) r4 = false;
  const size_t smem_fan = fan_bytes(r4);
  const int rec_cap = rw ? rw_rec_cap(p->fan_tile_max, r4) : fan_rec_cap(p->fan_tile_max, r4);
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
    int minb = 1;                                                                                               \
    if (rw) {                                                                                                   \
      FE_CUDA(cudaFuncSetAttribute(k_assemble_fan_rw<KC, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      FE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&minb, k_assemble_fan_rw<KC, R4>, kRwThreads, smem)); \
      if (minb < 1) minb = 1;                                                                                   \
      const int fgrid = grid < minb * ctx->num_sms ? grid : minb * ctx->num_sms;                                \
      k_assemble_fan_rw<KC, R4><<<fgrid, kRwThreads, smem, st>>>(p->n_owned, p->fan_ptr, RECS, p->fan_hdr,      \
                                                                 p->adj_ptr, xy, tab, vals, rec_cap,            \
                                                                 fan_warp_slot_bytes(dim, p->max_degree));      \
    } else {                                                                                                    \
      FE_CUDA(cudaFuncSetAttribute(k_assemble_fan<KC, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
      FE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&minb, k_assemble_fan<KC, R4>, kFanThreads, smem)); \
      if (minb < 1) minb = 1;                                                                                   \
      const int cgrid = grid_for(p->n_owned, kFanThreads);                                                      \
      const int fgrid = cgrid < minb * ctx->num_sms ? cgrid : minb * ctx->num_sms;                              \
      k_assemble_fan<KC, R4><<<fgrid, kFanThreads, smem, st>>>(p->n_owned, p->fan_ptr, RECS, p->fan_hdr,        \
                                                               p->adj_ptr, xy, tab, vals, rec_cap,              \
                                                               fan_warp_slot_bytes(dim, p->max_degree));        \
    }                                                                                                           \
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
