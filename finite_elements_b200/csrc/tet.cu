// Linear tetrahedra, 3 DOF per node: the element arithmetic of
// ElasticityTetrahedralElement3D (elements.py:663-876) and the deterministic row-owner assembly
// of its global matrices.
//
//   B = 1/(6V) [[a_i,0,0],[0,b_i,0],[0,0,c_i],[b_i,a_i,0],[0,c_i,b_i],[c_i,0,a_i]]   elements.py:719-751
//   D = E/((1+nu)(1-2nu)) [[1-nu,nu,nu,.],[nu,1-nu,nu,.],[nu,nu,1-nu,.],[., (1-2nu)/2 I3]]   :773-797
//   Ke = V B^T D B                                                                     :809-828
//   Me = rho V / 20 ((1 + delta_ij) (x) I3)                                            :830-857
// With g_i = grad N_i = (a_i, b_i, c_i) / (6V), lam = D_12 and mu = D_44, the 3x3 block of nodes
// (i, j) of V B^T D B is  V [ lam g_i g_j^T + mu g_j g_i^T + mu (g_i . g_j) I ]  -- evaluated in
// registers, the 6x12 B is never formed.
//
// The symbolic phase is fe_tet_plan_create (plan.cu).  Numeric assembly, deterministic in every variant (no atomics,
// fixed summation order; two runs are bit-identical):
//   6  k_tet_assemble_pipe    (default) persistent CTAs, tiles of 16 nodes staged in shared memory, next tile's inputs
//                             by cp.async; one lane per 3x3 block
//   5  k_tet_assemble_staged  the same, one tile per CTA, no prefetch (bit-identical to 6)
//   4  k_tet_gradient_table + k_tet_assemble_table   per-element records in a global table, one lane per block
//   3  k_tet_assemble_slots   one lane per block, geometry rebuilt at every visit
//   2  k_tet_assemble_tile / 1  k_tet_assemble   round 1: one thread owns a node's three rows and visits its
//                             incident elements in ascending order (shared-memory tile / global accumulation)
// Measured on B200 (1.33 M tetrahedra): 2.90 ms (1) -> 1.23 / 0.94 ms (2) -> 0.36 (3) -> 0.30 (4) -> 0.235 (5) -> 0.225 ms (6).
#include "common.cuh"
#include "elem.cuh"
#include "plan.cuh"

namespace fe {

struct TetGeom {
  double g[4][3];  // grad N_i
  double vol;      // |det [1 x y z]| / 6  (volmdlr TetrahedralElement.volume)
};

__device__ __forceinline__ TetGeom tet_geom_xyz(double x0, double y0, double z0, double x1, double y1, double z1, double x2,
                                                double y2, double z2, double x3, double y3, double z3) {
  // edge vectors from vertex 0
  const double ax = x1 - x0, ay = y1 - y0, az = z1 - z0;
  const double bx = x2 - x0, by = y2 - y0, bz = z2 - z0;
  const double cx = x3 - x0, cy = y3 - y0, cz = z3 - z0;
  // rows of the inverse of J = [a; b; c] are grad N_1..3 (N_i(p_j) = delta_ij); grad N_0 = -(sum)
  const double c1x = by * cz - bz * cy, c1y = bz * cx - bx * cz, c1z = bx * cy - by * cx;  // b x c
  const double c2x = cy * az - cz * ay, c2y = cz * ax - cx * az, c2z = cx * ay - cy * ax;  // c x a
  const double c3x = ay * bz - az * by, c3y = az * bx - ax * bz, c3z = ax * by - ay * bx;  // a x b
  const double det = ax * c1x + ay * c1y + az * c1z;                                       // a . (b x c) = 6 V signed
  const double inv = 1.0 / det;
  TetGeom t;
  t.g[1][0] = c1x * inv, t.g[1][1] = c1y * inv, t.g[1][2] = c1z * inv;
  t.g[2][0] = c2x * inv, t.g[2][1] = c2y * inv, t.g[2][2] = c2z * inv;
  t.g[3][0] = c3x * inv, t.g[3][1] = c3y * inv, t.g[3][2] = c3z * inv;
#pragma unroll
  for (int d = 0; d < 3; ++d) t.g[0][d] = -(t.g[1][d] + t.g[2][d] + t.g[3][d]);
  t.vol = fabs(det) / 6.0;
  return t;
}

__device__ __forceinline__ TetGeom tet_geom(const double *__restrict__ coords, int n0, int n1, int n2, int n3) {
  const double *p0 = coords + 3 * (int64_t)n0, *p1 = coords + 3 * (int64_t)n1;
  const double *p2 = coords + 3 * (int64_t)n2, *p3 = coords + 3 * (int64_t)n3;
  return tet_geom_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
}

// (lam V, mu V) for the stiffness, (rho V / 20, -) for the mass
struct TetMat {
  double p0, p1;
};

__device__ __forceinline__ TetMat tet_material(int kind, const double *__restrict__ mat, int mid, double vol) {
  const double e_mod = mat[4 * mid + 0], nu = mat[4 * mid + 1], rho = mat[4 * mid + 3];
  TetMat m;
  if (kind == FE_ELAST_TET) {
    const double coeff = e_mod / ((1 + nu) * (1 - 2 * nu));  // elements.py:794
    m.p0 = coeff * nu * vol;
    m.p1 = coeff * ((1 - 2 * nu) / 2) * vol;
  } else {
    m.p0 = (rho * vol) / 20;  // elements.py:854
    m.p1 = 0.0;
  }
  return m;
}

// 3x3 block (v, j) of the element matrix, row-major in out[9]
__device__ __forceinline__ void tet_block(int kind, const TetGeom &t, const TetMat &m, int v, int j, double out[9]) {
  if (kind == FE_ELAST_TET) {
    const double dot = t.g[v][0] * t.g[j][0] + t.g[v][1] * t.g[j][1] + t.g[v][2] * t.g[j][2];
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int b = 0; b < 3; ++b)
        out[3 * a + b] = m.p0 * t.g[v][a] * t.g[j][b] + m.p1 * t.g[v][b] * t.g[j][a] + (a == b ? m.p1 * dot : 0.0);
  } else {
    const double d = (v == j) ? 2.0 * m.p0 : m.p0;
#pragma unroll
    for (int q = 0; q < 9; ++q) out[q] = (q % 4 == 0) ? d : 0.0;
  }
}

// The four 3x3 blocks of node `self`'s row of the element matrix, out[j] towards nbr[j] (nbr[0] =
// self).  The element is relabelled so that the owned node is local vertex 0 (cyclic shift of the
// connectivity: the element matrix does not depend on the labelling), which keeps every index into
// the gradients a compile-time constant -- no local-memory arrays in the assembly kernels.
struct TetRow {
  int nbr[4];
  double blk[4][9];
};

__device__ __forceinline__ bool tet_row(int kind, const double *__restrict__ coords, const int32_t *__restrict__ conn,
                                        const int32_t *__restrict__ mat_id, const double *__restrict__ mat, int32_t e,
                                        int32_t self, int skip_before, TetRow &r) {
  const int4 c = *reinterpret_cast<const int4 *>(conn + 4 * (int64_t)e);
  // local vertex of `self`: the skip_before-th match (an element may list a node twice)
  int v = -1, seen = 0;
  if (c.x == self && seen++ == skip_before && v < 0) v = 0;
  if (c.y == self && v < 0 && seen++ == skip_before) v = 1;
  if (c.z == self && v < 0 && seen++ == skip_before) v = 2;
  if (c.w == self && v < 0 && seen++ == skip_before) v = 3;
  if (v < 0) return false;
  r.nbr[0] = self;
  r.nbr[1] = v == 0 ? c.y : (v == 1 ? c.z : (v == 2 ? c.w : c.x));
  r.nbr[2] = v == 0 ? c.z : (v == 1 ? c.w : (v == 2 ? c.x : c.y));
  r.nbr[3] = v == 0 ? c.w : (v == 1 ? c.x : (v == 2 ? c.y : c.z));
  const TetGeom t = tet_geom(coords, r.nbr[0], r.nbr[1], r.nbr[2], r.nbr[3]);
  const TetMat m = tet_material(kind, mat, mat_id ? mat_id[e] : 0, t.vol);
#pragma unroll
  for (int j = 0; j < 4; ++j) tet_block(kind, t, m, 0, j, r.blk[j]);
  return true;
}

__global__ void __launch_bounds__(128) k_tet_elem_matrices(int kind, int64_t n_elems, const double *__restrict__ coords,
                                                          const int32_t *__restrict__ conn,
                                                          const int32_t *__restrict__ mat_id,
                                                          const double *__restrict__ mat, double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int4 c = *reinterpret_cast<const int4 *>(conn + 4 * e);
  const TetGeom t = tet_geom(coords, c.x, c.y, c.z, c.w);
  const TetMat m = tet_material(kind, mat, mat_id ? mat_id[e] : 0, t.vol);
  double *o = out + 144 * e;
  for (int v = 0; v < 4; ++v)
    for (int j = 0; j < 4; ++j) {
      double b[9];
      tet_block(kind, t, m, v, j, b);
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < 3; ++q) o[(3 * v + a) * 12 + 3 * j + q] = b[3 * a + q];
    }
}

// Element post-processing of a solution u (results.py:809-830, :769-781 for 3 DOF per node):
// out[e] = (eps_xx, eps_yy, eps_zz, gamma_xy, gamma_yz, gamma_zx,  sig_xx, sig_yy, sig_zz, tau_xy,
// tau_yz, tau_zx,  energy) with strain = B u_e, stress = D B u_e, energy = 1/2 u_e^T Ke u_e
// = V/2 strain . stress  (Ke = V B^T D B).
__global__ void __launch_bounds__(128) k_tet_post(int64_t n_elems, const double *__restrict__ coords,
                                                 const int32_t *__restrict__ conn, const int32_t *__restrict__ mat_id,
                                                 const double *__restrict__ mat, const double *__restrict__ u,
                                                 double *__restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int4 c = *reinterpret_cast<const int4 *>(conn + 4 * e);
  const int nodes[4] = {c.x, c.y, c.z, c.w};
  const TetGeom t = tet_geom(coords, c.x, c.y, c.z, c.w);
  double eps[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const double ux = u[3 * (int64_t)nodes[i]], uy = u[3 * (int64_t)nodes[i] + 1], uz = u[3 * (int64_t)nodes[i] + 2];
    const double a = t.g[i][0], b = t.g[i][1], cz = t.g[i][2];
    eps[0] += a * ux;            // elements.py:741-746: the six rows of B
    eps[1] += b * uy;
    eps[2] += cz * uz;
    eps[3] += b * ux + a * uy;
    eps[4] += cz * uy + b * uz;
    eps[5] += cz * ux + a * uz;
  }
  const int mid = mat_id ? mat_id[e] : 0;
  const double e_mod = mat[4 * mid + 0], nu = mat[4 * mid + 1];
  const double coeff = e_mod / ((1 + nu) * (1 - 2 * nu));
  const double da = coeff * (1 - nu), db = coeff * nu, ds = coeff * ((1 - 2 * nu) / 2);
  double sig[6];
  sig[0] = da * eps[0] + db * eps[1] + db * eps[2];
  sig[1] = db * eps[0] + da * eps[1] + db * eps[2];
  sig[2] = db * eps[0] + db * eps[1] + da * eps[2];
  sig[3] = ds * eps[3];
  sig[4] = ds * eps[4];
  sig[5] = ds * eps[5];
  double w = 0.0;
#pragma unroll
  for (int k = 0; k < 6; ++k) w += eps[k] * sig[k];
  double *o = out + 13 * e;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    o[k] = eps[k];
    o[6 + k] = sig[k];
  }
  o[12] = 0.5 * t.vol * w;
}

// One thread per owned node.  corner_elem[corner_ptr[i] .. corner_ptr[i+1]) = the elements incident
// to node i in ascending order; adj[adj_ptr[i] .. adj_ptr[i+1]) = its sorted neighbour nodes (incl.
// itself).  vals holds the node's three rows back to back: row r at 9 adj_ptr[i] + r * 3 deg, the
// block towards neighbour slot k in columns 3k .. 3k+2 (the layout csr3 of the host layer exports).
__global__ void __launch_bounds__(128) k_tet_assemble(int kind, int32_t n_owned, const int32_t *__restrict__ corner_ptr,
                                                     const int32_t *__restrict__ corner_elem,
                                                     const int32_t *__restrict__ adj_ptr,
                                                     const int32_t *__restrict__ adj, const double *__restrict__ coords,
                                                     const int32_t *__restrict__ conn,
                                                     const int32_t *__restrict__ mat_id,
                                                     const double *__restrict__ mat, double *__restrict__ vals) {
  const int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_owned) return;
  const int32_t a0 = adj_ptr[i], deg = adj_ptr[i + 1] - a0;
  double *rows = vals + 9 * (int64_t)a0;
  for (int q = 0; q < 9 * deg; ++q) rows[q] = 0.0;
  for (int32_t cidx = corner_ptr[i]; cidx < corner_ptr[i + 1]; ++cidx) {
    const int32_t e = corner_elem[cidx];
    // (an element listing the node twice appears twice in the corner list: k-th appearance = k-th vertex)
    const int dup = (cidx > corner_ptr[i] && corner_elem[cidx - 1] == e) ? 1 : 0;
    TetRow r;
    if (!tet_row(kind, coords, conn, mat_id, mat, e, i, dup, r)) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int lo = 0, hi = deg - 1;  // slot of the neighbour in the sorted list
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (adj[a0 + mid] < r.nbr[j]) lo = mid + 1; else hi = mid;
      }
#pragma unroll
      for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int q = 0; q < 3; ++q) rows[a * 3 * deg + 3 * lo + q] += r.blk[j][3 * a + q];
    }
  }
}

// Same traversal with the accumulation in shared memory: a CTA owns kTetTile consecutive nodes, i.e.
// one contiguous slice of vals; thread t accumulates its node's 9 deg values in column t of a
// [9 max_deg][kTetTile + 1] tile (conflict-free: consecutive threads, consecutive banks), then the
// warps stream the tile out row segment by row segment with coalesced stores -- every value reaches
// HBM exactly once instead of being read-modified-written through L1/L2 once per contribution.
// The additions happen in the same order as in k_tet_assemble: the two kernels are bit-identical.
constexpr int kTetTile = 64;
constexpr int kTetLD = kTetTile + 1;

__global__ void __launch_bounds__(kTetTile) k_tet_assemble_tile(
    int kind, int32_t n_owned, const int32_t *__restrict__ corner_ptr, const int32_t *__restrict__ corner_elem,
    const int32_t *__restrict__ adj_ptr, const int32_t *__restrict__ adj, const double *__restrict__ coords,
    const int32_t *__restrict__ conn, const int32_t *__restrict__ mat_id, const double *__restrict__ mat,
    double *__restrict__ vals) {
  extern __shared__ __align__(16) double acc[];  // [9 * max_deg][kTetLD]
  const int tid = threadIdx.x;
  const int32_t n0 = blockIdx.x * kTetTile;
  const int32_t i = n0 + tid;
  const int n_in_tile = min(kTetTile, n_owned - n0);
  int32_t a0 = 0, deg = 0;
  if (i < n_owned) {
    a0 = adj_ptr[i];
    deg = adj_ptr[i + 1] - a0;
    double *my = acc + tid;
    for (int q = 0; q < 9 * deg; ++q) my[q * kTetLD] = 0.0;
    for (int32_t cidx = corner_ptr[i]; cidx < corner_ptr[i + 1]; ++cidx) {
      const int32_t e = corner_elem[cidx];
      const int dup = (cidx > corner_ptr[i] && corner_elem[cidx - 1] == e) ? 1 : 0;
      TetRow r;
      if (!tet_row(kind, coords, conn, mat_id, mat, e, i, dup, r)) continue;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        int lo = 0, hi = deg - 1;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (adj[a0 + mid] < r.nbr[j]) lo = mid + 1; else hi = mid;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
          for (int q = 0; q < 3; ++q) my[(a * 3 * deg + 3 * lo + q) * kTetLD] += r.blk[j][3 * a + q];
      }
    }
  }
  __syncthreads();
  // write-out: warp w takes nodes w, w + 2, ... of the tile; lanes run over the node's 9 deg values
  const int lane = tid & 31, w = tid >> 5;
  for (int k = w; k < n_in_tile; k += kTetTile / 32) {
    const int32_t b0 = adj_ptr[n0 + k];
    const int len = 9 * (adj_ptr[n0 + k + 1] - b0);
    double *dst = vals + 9 * (int64_t)b0;
    for (int q = lane; q < len; q += 32) dst[q] = acc[q * kTetLD + k];
  }
}

// ---------------------------------------------------------------------------------------
// Default: one LANE per block of the global matrix.  16 lanes share a node; lane l owns the neighbour
// slots l, l + 16, ... and walks, for each, the plan's list of elements that hold both nodes (ascending
// element id, fe_tet_plan_create), accumulating the 3x3 block in registers and writing it once -- no
// shared-memory tile (the node-owner kernels above are capped at 6 warps per SM by theirs), no slot
// search, no serial 24-element loop.  Only the two gradients a block needs are evaluated: with the
// element relabelled (other_a, self, neighbour, other_b),  grad N_self = (b x c) / det and
// grad N_nbr = (c x a) / det  for the edge vectors a, b, c from other_a.  The diagonal block follows from the
// off-diagonal ones (rigid-translation null space, see the end of the kernel): fixed order, bit-reproducible.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ int pick4(const int4 &c, int i) { return i == 0 ? c.x : (i == 1 ? c.y : (i == 2 ? c.z : c.w)); }

#ifndef FE_TET_MINB
#define FE_TET_MINB 3  // resident CTAs per SM the register allocation targets (80 registers; measured best of 2/3/4)
#endif
template <bool MASS>
__global__ void __launch_bounds__(256, FE_TET_MINB) k_tet_assemble_slots(int32_t n_owned, const int32_t *__restrict__ adj_ptr,
                                                           const int32_t *__restrict__ adj,
                                                           const int32_t *__restrict__ contrib_ptr,
                                                           const int32_t *__restrict__ contrib,
                                                           const double *__restrict__ coords,
                                                           const int32_t *__restrict__ conn,
                                                           const int32_t *__restrict__ mat_id,
                                                           const double *__restrict__ mat, double *__restrict__ vals) {
  const int32_t node = (blockIdx.x * 256 + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  if (node >= n_owned) return;  // (whole 16-lane groups leave together)
  const int32_t a0 = __ldg(adj_ptr + node);
  const int deg = __ldg(adj_ptr + node + 1) - a0;
  double *rows = vals + 9 * (int64_t)a0;
  const int4 *conn4 = reinterpret_cast<const int4 *>(conn);
  // this lane's share of the sum of the node's off-diagonal blocks, parked in shared memory (one column
  // per thread) so that it does not occupy 18 registers across the element loop
  __shared__ double tot_s[9][256];
#pragma unroll
  for (int q = 0; q < 9; ++q) tot_s[q][threadIdx.x] = 0.0;
  int kself = -1;
  for (int k = lane; k < deg; k += 16) {
    if (__ldg(adj + a0 + k) == node) kself = k;
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    const int32_t q0 = __ldg(contrib_ptr + a0 + k), q1 = __ldg(contrib_ptr + a0 + k + 1);
    // software pipeline: the next entry's code and connectivity row travel while this one is evaluated
    int32_t code = 0;
    int4 c = make_int4(0, 0, 0, 0);
    if (q0 < q1) {
      code = __ldg(contrib + q0);
      c = __ldg(conn4 + (code >> 4));
    }
    for (int32_t q = q0; q < q1; ++q) {
      const int32_t e = code >> 4;
      const int vi = (code >> 2) & 3, vj = code & 3;
      const int4 cc = c;
      if (q + 1 < q1) {
        code = __ldg(contrib + q + 1);
        c = __ldg(conn4 + (code >> 4));
      }
      const unsigned rest = 0xFu ^ (1u << vi) ^ (1u << vj);
      const int oa = __ffs(rest) - 1, ob = 31 - __clz(rest);
      const double *p0 = coords + 3 * (int64_t)pick4(cc, oa), *pi = coords + 3 * (int64_t)pick4(cc, vi);
      const double *pj = coords + 3 * (int64_t)pick4(cc, vj), *pb = coords + 3 * (int64_t)pick4(cc, ob);
      const double x0 = __ldg(p0), y0 = __ldg(p0 + 1), z0 = __ldg(p0 + 2);
      const double ax = __ldg(pi) - x0, ay = __ldg(pi + 1) - y0, az = __ldg(pi + 2) - z0;
      const double bx = __ldg(pj) - x0, by = __ldg(pj + 1) - y0, bz = __ldg(pj + 2) - z0;
      const double cx = __ldg(pb) - x0, cy = __ldg(pb + 1) - y0, cz = __ldg(pb + 2) - z0;
      const double ux = by * cz - bz * cy, uy = bz * cx - bx * cz, uz = bx * cy - by * cx;  // b x c
      const double det = ax * ux + ay * uy + az * uz;                                        // 6 V, signed
      const TetMat m = tet_material(MASS ? FE_MASS_TET : FE_ELAST_TET, mat, mat_id ? __ldg(mat_id + e) : 0, fabs(det) / 6.0);
      if (MASS) {
        acc[0] += m.p0;
        acc[4] += m.p0;
        acc[8] += m.p0;
      } else {
        const double inv = 1.0 / det;
        const double gi[3] = {ux * inv, uy * inv, uz * inv};                                  // grad N_self
        const double gj[3] = {(cy * az - cz * ay) * inv, (cz * ax - cx * az) * inv, (cx * ay - cy * ax) * inv};  // (c x a) / det
        const double mdot = m.p1 * (gi[0] * gj[0] + gi[1] * gj[1] + gi[2] * gj[2]);
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const double lr = m.p0 * gi[r], mr = m.p1 * gj[r];
#pragma unroll
          for (int t = 0; t < 3; ++t) acc[3 * r + t] += lr * gj[t] + mr * gi[t] + (r == t ? mdot : 0.0);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        rows[r * 3 * deg + 3 * k + t] = acc[3 * r + t];
        tot_s[3 * r + t][threadIdx.x] += acc[3 * r + t];
      }
  }
  double tot[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) tot[q] = tot_s[q][threadIdx.x];
  // Diagonal block.  Stiffness: a rigid translation carries no force, so every block row of K sums to zero
  // and K_ii = -(sum of the off-diagonal blocks) -- the same number as the sum of the elements' self blocks
  // up to rounding, without evaluating them.  Consistent mass: an element gives 2 m to (i, i) and m to each
  // of its three (i, j), so M_ii = 2/3 of the off-diagonal sum.  The 16 partial sums are added by
  // xor-shuffles in a fixed order.
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
#pragma unroll
  for (int q = 0; q < 9; ++q) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tot[q] += __shfl_xor_sync(gmask, tot[q], o);
  }
  if (kself >= 0) {
    const double f = MASS ? (2.0 / 3.0) : -1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) rows[r * 3 * deg + 3 * kself + t] = f * tot[3 * r + t];
  }
}

// ---------------------------------------------------------------------------------------
// Two-pass form of the lane-per-block assembly (default).  Pass 1 evaluates every element ONCE and
// leaves four 32-byte records h_k = (sqrt(mu V) grad N_k, lambda / mu) in a scratch table (128 B per
// element).  Pass 2 walks the same per-block element lists and reads exactly the two records it needs,
// each with ONE 256-bit load (LDG.E.256: one sector per lane):
//     block(i, j) = (lambda / mu) h_i h_j^T + h_j h_i^T + (h_i . h_j) I        (= V B_i^T D B_j)
// ncu on the earlier forms of this walk: L1TEX 95 % busy on scattered 8-byte loads (7 sector accesses per
// visit with an unpadded table, 13 when the geometry is rebuilt from connectivity and coordinates);
// the number of divergent accesses per visit (l1tex__data_pipe_lsu_wavefronts ~ 1 per lane access per
// cycle and SM), not FP64 or DRAM, sets the time: 0.94 ms (round 1) -> 0.36 (rebuild) -> 0.30 ms (table)
// on 1.33 M elements.  Staging the rows in shared memory for coalesced stores was tried and bought nothing.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void st_global_f64x4(double *p, double a, double b, double c, double d) {
  asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}
struct D4 {
  double x, y, z, w;
};
__device__ __forceinline__ D4 ldg_nc_f64x4(const double *p) {
  D4 r;
  asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}

template <bool MASS>
__global__ void __launch_bounds__(128) k_tet_gradient_table(int64_t n_elems, const double *__restrict__ coords,
                                                           const int32_t *__restrict__ conn,
                                                           const int32_t *__restrict__ mat_id,
                                                           const double *__restrict__ mat, double *__restrict__ table) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_elems) return;
  const int4 c = __ldg(reinterpret_cast<const int4 *>(conn) + e);
  const TetGeom t = tet_geom(coords, c.x, c.y, c.z, c.w);
  const TetMat m = tet_material(MASS ? FE_MASS_TET : FE_ELAST_TET, mat, mat_id ? mat_id[e] : 0, t.vol);
  // elasticity: m = (lambda V, mu V); mass: m.p0 = rho V / 20 rides in the fourth component
  const double sc = MASS ? 0.0 : sqrt(m.p1);
  const double w = MASS ? m.p0 : (m.p1 != 0.0 ? m.p0 / m.p1 : 0.0);
#pragma unroll
  for (int k = 0; k < 4; ++k) st_global_f64x4(table + 16 * e + 4 * k, sc * t.g[k][0], sc * t.g[k][1], sc * t.g[k][2], w);
}

template <bool MASS>
__device__ __forceinline__ void tet_pair_add(const D4 &hi, const D4 &hj, double (&acc)[9]) {
  if (MASS) {
    acc[0] += hi.w;
    acc[4] += hi.w;
    acc[8] += hi.w;
  } else {
    const double dot = hi.x * hj.x + hi.y * hj.y + hi.z * hj.z;
    const double gi[3] = {hi.x, hi.y, hi.z}, gj[3] = {hj.x, hj.y, hj.z};
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const double lr = hi.w * gi[r];
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[3 * r + c] += lr * gj[c] + gj[r] * gi[c] + (r == c ? dot : 0.0);
    }
  }
}

template <bool MASS>
__global__ void __launch_bounds__(256, 4) k_tet_assemble_table(int32_t n_owned, const int32_t *__restrict__ adj_ptr,
                                                              const int32_t *__restrict__ adj,
                                                              const int32_t *__restrict__ contrib_ptr,
                                                              const int32_t *__restrict__ contrib,
                                                              const double *__restrict__ table, double *__restrict__ vals) {
  const int32_t node = (blockIdx.x * 256 + threadIdx.x) >> 4;
  const int lane = threadIdx.x & 15;
  if (node >= n_owned) return;
  const int32_t a0 = __ldg(adj_ptr + node);
  const int deg = __ldg(adj_ptr + node + 1) - a0;
  double *rows = vals + 9 * (int64_t)a0;
  __shared__ double tot_s[9][256];  // per-lane sum of its off-diagonal blocks (kept out of the registers)
#pragma unroll
  for (int q = 0; q < 9; ++q) tot_s[q][threadIdx.x] = 0.0;
  int kself = -1;
  for (int k = lane; k < deg; k += 16) {
    if (__ldg(adj + a0 + k) == node) kself = k;
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    const int32_t q0 = __ldg(contrib_ptr + a0 + k), q1 = __ldg(contrib_ptr + a0 + k + 1);
    // two visits in flight: both codes, then the four 256-bit record loads, then the arithmetic
    auto rec = [&](int32_t code, int v) { return ldg_nc_f64x4(table + 16 * (int64_t)(code >> 4) + 4 * v); };
    int32_t q = q0;
    for (; q + 1 < q1; q += 2) {
      const int32_t c0 = __ldg(contrib + q), c1 = __ldg(contrib + q + 1);
      const D4 i0 = rec(c0, (c0 >> 2) & 3), j0 = MASS ? i0 : rec(c0, c0 & 3);
      const D4 i1 = rec(c1, (c1 >> 2) & 3), j1 = MASS ? i1 : rec(c1, c1 & 3);
      tet_pair_add<MASS>(i0, j0, acc);
      tet_pair_add<MASS>(i1, j1, acc);
    }
    if (q < q1) {
      const int32_t c0 = __ldg(contrib + q);
      const D4 i0 = rec(c0, (c0 >> 2) & 3), j0 = MASS ? i0 : rec(c0, c0 & 3);
      tet_pair_add<MASS>(i0, j0, acc);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        rows[r * 3 * deg + 3 * k + t] = acc[3 * r + t];
        tot_s[3 * r + t][threadIdx.x] += acc[3 * r + t];
      }
  }
  // diagonal block from the off-diagonal ones: see k_tet_assemble_slots
  double tot[9];
#pragma unroll
  for (int q = 0; q < 9; ++q) tot[q] = tot_s[q][threadIdx.x];
  const unsigned gmask = 0xffffu << (threadIdx.x & 16);
#pragma unroll
  for (int q = 0; q < 9; ++q) {
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) tot[q] += __shfl_xor_sync(gmask, tot[q], o);
  }
  if (kself >= 0) {
    const double f = MASS ? (2.0 / 3.0) : -1.0;
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) rows[r * 3 * deg + 3 * kself + t] = f * tot[3 * r + t];
  }
}

// ---------------------------------------------------------------------------------------
// Staged form (default): one CTA per tile of kTetStageNodes consecutive nodes, everything through shared
// memory.  The table walk above is bound by the L1's rate for scattered sectors (ncu: ~1 sector per clock and
// SM, 109 M sectors per launch); here the only scattered global loads left are one coordinate triple per tile
// node, every other input is a contiguous slice prepared by the plan (plan.cu: k_tet_stage_tiles):
//   stage  : coordinates of the tile's nodes, contribution offsets and 16-bit codes -> shared memory
//   phase A: every element of the tile is evaluated once (one thread each) from the staged coordinates into
//            four 32-byte records (sqrt(mu V) grad N_k, lambda / mu) in shared memory
//   phase B: one lane per block walks its code list (LDS.U16) and reads the two records it needs with two
//            LDS.128 each; blocks land in the shared-memory image of the tile's slice of vals
//   diag   : nine lanes per node add the row's off-diagonal blocks in slot order (K_ii = -sum, M_ii = 2/3 sum)
//   store  : the image is copied out with fully coalesced stores.
// An element is evaluated once per tile it touches (~3x on a Kuhn mesh numbered lexicographically).
// ---------------------------------------------------------------------------------------
constexpr int kTetStageThreads = kTetStageNodes * 16;
constexpr int kTetGradStride = 13;  // doubles per element in shared memory: 4 x (sqrt(mu V) grad N_k) + lambda / mu
                                    // (odd stride: elements spread over all banks; 128 B records aliased 8-way)

struct TetStageSmem {
  size_t grad, out, xyz, cptr, aptr, codes, total;
};
__host__ __device__ inline TetStageSmem tet_stage_smem(int ne_max, int na_max, int nn_max, int nq_max) {
  TetStageSmem s;
  s.grad = 0;
  s.out = s.grad + ((size_t)ne_max * kTetGradStride * 8 + 15) / 16 * 16;
  s.xyz = s.out + (size_t)na_max * 72;
  s.cptr = s.xyz + (size_t)nn_max * 24;
  s.aptr = s.cptr + ((size_t)na_max + 2) / 2 * 8;
  s.codes = s.aptr + ((size_t)kTetStageNodes + 2) / 2 * 8;
  s.total = (s.codes + (size_t)nq_max * 2 + 15) / 16 * 16;
  return s;
}

template <bool MASS>
__global__ void __launch_bounds__(kTetStageThreads, 4) k_tet_assemble_staged(
    int32_t n_owned, const int32_t *__restrict__ adj_ptr, const int32_t *__restrict__ contrib_ptr,
    const uint16_t *__restrict__ contrib16, const int4 *__restrict__ tile_desc, const int32_t *__restrict__ tile_elist,
    const ushort4 *__restrict__ tile_erec, const int32_t *__restrict__ tile_nodes, const uint8_t *__restrict__ kself,
    const double *__restrict__ coords, const int32_t *__restrict__ mat_id, const double *__restrict__ mat,
    double *__restrict__ vals, int ne_max, int na_max, int nn_max, int nq_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TetStageSmem L = tet_stage_smem(ne_max, na_max, nn_max, nq_max);
  double *s_grad = reinterpret_cast<double *>(smem_raw + L.grad);
  double *s_out = reinterpret_cast<double *>(smem_raw + L.out);
  double *s_xyz = reinterpret_cast<double *>(smem_raw + L.xyz);
  int32_t *s_cptr = reinterpret_cast<int32_t *>(smem_raw + L.cptr);
  int32_t *s_aptr = reinterpret_cast<int32_t *>(smem_raw + L.aptr);
  uint16_t *s_codes = reinterpret_cast<uint16_t *>(smem_raw + L.codes);

  const int tid = threadIdx.x;
  const int32_t tile = blockIdx.x, n0 = tile * kTetStageNodes, n1 = min(n0 + kTetStageNodes, n_owned);
  const int nt = n1 - n0;
  // one descriptor per tile (plan.cu): every slice below is addressed from it -- one dependent load, not three
  const int4 d0 = __ldg(tile_desc + 2 * tile), d1 = __ldg(tile_desc + 2 * tile + 1);
  const int32_t a0 = d0.x, eb = d0.z, nb = d1.x, q0 = d1.z;
  const int na = d0.y, ne = d0.w, nn = d1.y, nq = d1.w;
  // ---- stage (all loads independent of each other except coordinates <- node list)
  ushort4 er0 = make_ushort4(0, 0, 0, 0), er1 = er0;  // this thread's elements of phase A
  int32_t eg0 = 0, eg1 = 0;
  if (tid < ne) {
    er0 = __ldg(tile_erec + eb + tid);
    if (mat_id) eg0 = __ldg(tile_elist + eb + tid);
  }
  if (tid + kTetStageThreads < ne) {
    er1 = __ldg(tile_erec + eb + tid + kTetStageThreads);
    if (mat_id) eg1 = __ldg(tile_elist + eb + tid + kTetStageThreads);
  }
  for (int k = tid; k < 3 * nn; k += kTetStageThreads) {
    const int nd = k / 3;
    s_xyz[k] = __ldg(coords + 3 * (int64_t)__ldg(tile_nodes + nb + nd) + (k - 3 * nd));
  }
  if (tid <= nt) s_aptr[tid] = __ldg(adj_ptr + n0 + tid) - a0;
  for (int k = tid; k <= na; k += kTetStageThreads) s_cptr[k] = __ldg(contrib_ptr + a0 + k) - q0;
  for (int k = tid; k < nq; k += kTetStageThreads) s_codes[k] = __ldg(contrib16 + q0 + k);
  const int mid0 = (mat_id && tid < ne) ? __ldg(mat_id + eg0) : 0;
  const int mid1 = (mat_id && tid + kTetStageThreads < ne) ? __ldg(mat_id + eg1) : 0;
  __syncthreads();
  // ---- phase A: the tile's elements, once each
  auto evaluate = [&](int le, const ushort4 r, int mid) {
    const double *p0 = s_xyz + 3 * r.x, *p1 = s_xyz + 3 * r.y, *p2 = s_xyz + 3 * r.z, *p3 = s_xyz + 3 * r.w;
    const TetGeom t = tet_geom_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
    const TetMat m = tet_material(MASS ? FE_MASS_TET : FE_ELAST_TET, mat, mid, t.vol);
    const double sc = MASS ? 0.0 : sqrt(m.p1);
    double *g = s_grad + kTetGradStride * le;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      g[3 * k] = sc * t.g[k][0];
      g[3 * k + 1] = sc * t.g[k][1];
      g[3 * k + 2] = sc * t.g[k][2];
    }
    g[12] = MASS ? m.p0 : (m.p1 != 0.0 ? m.p0 / m.p1 : 0.0);
  };
  if (tid < ne) evaluate(tid, er0, mid0);
  if (tid + kTetStageThreads < ne) evaluate(tid + kTetStageThreads, er1, mid1);
  for (int le = tid + 2 * kTetStageThreads; le < ne; le += kTetStageThreads)
    evaluate(le, __ldg(tile_erec + eb + le), mat_id ? __ldg(mat_id + __ldg(tile_elist + eb + le)) : 0);
  __syncthreads();
  // ---- phase B: one lane per block
  const int nl = tid >> 4, lane = tid & 15;
  int al = 0, deg = 0;
  if (nl < nt) {
    al = s_aptr[nl];
    deg = s_aptr[nl + 1] - al;
  }
  double *rows = s_out + 9 * al;
  for (int k = lane; k < deg; k += 16) {
    double acc[9];
#pragma unroll
    for (int q = 0; q < 9; ++q) acc[q] = 0.0;
    const int qb = s_cptr[al + k + 1];
    for (int q = s_cptr[al + k]; q < qb; ++q) {
      const uint32_t code = s_codes[q];
      const double *ge = s_grad + kTetGradStride * (code >> 4);
      const double *gi = ge + 3 * ((code >> 2) & 3), *gj = ge + 3 * (code & 3);
      const D4 hi = {gi[0], gi[1], gi[2], ge[12]};
      D4 hj = hi;
      if (!MASS) hj = D4{gj[0], gj[1], gj[2], 0.0};
      tet_pair_add<MASS>(hi, hj, acc);
    }
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
      for (int t = 0; t < 3; ++t) rows[r * 3 * deg + 3 * k + t] = acc[3 * r + t];
  }
  __syncwarp();
  // ---- diagonal block from the off-diagonal ones (see k_tet_assemble_slots), slot order
  if (lane < 9 && deg > 0) {
    const int ks = __ldg(kself + n0 + nl);
    const int r = lane / 3, t = lane - 3 * r;
    const double *row = rows + r * 3 * deg + t;
    double sum = 0.0;
    for (int k = 0; k < deg; ++k)
      if (k != ks) sum += row[3 * k];
    rows[r * 3 * deg + 3 * ks + t] = (MASS ? (2.0 / 3.0) : -1.0) * sum;
  }
  __syncthreads();
  // ---- store the image of vals[9 a0 .. 9 (a0 + na))
  double *dst = vals + 9 * (int64_t)a0;
  for (int k = tid; k < 9 * na; k += kTetStageThreads) dst[k] = s_out[k];
}

// ---------------------------------------------------------------------------------------
// Pipelined form of the staged kernel (default): persistent CTAs, the inputs of tile t+1 arrive by cp.async while
// tile t is computed.  ncu on the one-tile-per-CTA form above (r02 capture N): 42 % of the stall samples in the
// staging phase -- three dependent global round trips (descriptor -> slices and node list -> coordinates) with four
// CTAs per SM to overlap them.  Here per trip
//   top    : wait for the coordinates of tile t (requested during tile t-1), barrier; request the slices of tile t+1
//            (node list, element records, contribution offsets / codes, k_self) and the descriptor of tile t+2
//   phase A: the tile's elements, once each, from the staged coordinates
//   middle : wait for those slices (the barrier that closes phase A publishes them); request the coordinates of
//            tile t+1 through its node list
//   phase B, diagonal, copy-out as above.
// Two input buffers, three descriptors in flight; the arithmetic and the order of every sum are those of the form
// above: bit-identical values.
// ---------------------------------------------------------------------------------------
struct TetPipeSmem {
  size_t grad, out, in0, in_bytes, total;
  // offsets inside one input buffer
  size_t nodes, erec, elist, cptr, aptr, codes, kself, xyz;
};
__host__ __device__ inline TetPipeSmem tet_pipe_smem(int ne_max, int na_max, int nn_max, int nq_max) {
  auto up16 = [](size_t b) { return (b + 15) / 16 * 16; };
  TetPipeSmem s;
  s.nodes = 0;
  s.erec = s.nodes + up16((size_t)nn_max * 4);
  s.elist = s.erec + up16((size_t)ne_max * 8);
  s.cptr = s.elist + up16((size_t)ne_max * 4);
  s.aptr = s.cptr + up16(((size_t)na_max + 1) * 4);
  s.codes = s.aptr + up16(((size_t)kTetStageNodes + 1) * 4);
  s.kself = s.codes + up16(((size_t)nq_max + 3) * 2);
  s.xyz = s.kself + up16((size_t)kTetStageNodes + 4);
  s.in_bytes = s.xyz + up16((size_t)nn_max * 24);
  s.grad = 128;  // three descriptors (2 x int4 each) in front
  s.out = s.grad + up16((size_t)ne_max * kTetGradStride * 8);
  s.in0 = s.out + up16((size_t)na_max * 72);
  s.total = s.in0 + 2 * s.in_bytes;
  return s;
}

__device__ __forceinline__ void cp_async_b4(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_b8(void *dst, const void *src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_b16(void *dst, const void *src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <bool MASS>
__global__ void __launch_bounds__(kTetStageThreads, 3) k_tet_assemble_pipe(
    int32_t n_owned, int32_t n_tiles, const int32_t *__restrict__ adj_ptr, const int32_t *__restrict__ contrib_ptr,
    const uint16_t *__restrict__ contrib16, const int4 *__restrict__ tile_desc, const int32_t *__restrict__ tile_elist,
    const ushort4 *__restrict__ tile_erec, const int32_t *__restrict__ tile_nodes, const uint8_t *__restrict__ kself,
    const double *__restrict__ coords, const int32_t *__restrict__ mat_id, const double *__restrict__ mat,
    double *__restrict__ vals, int ne_max, int na_max, int nn_max, int nq_max) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const TetPipeSmem L = tet_pipe_smem(ne_max, na_max, nn_max, nq_max);
  int4 *s_desc = reinterpret_cast<int4 *>(smem_raw);  // [3][2]
  double *s_grad = reinterpret_cast<double *>(smem_raw + L.grad);
  double *s_out = reinterpret_cast<double *>(smem_raw + L.out);
  const int tid = threadIdx.x;
  const int stride = gridDim.x;

  auto in_buf = [&](int b) { return smem_raw + L.in0 + (size_t)b * L.in_bytes; };
  // slices of a tile whose descriptor sits in s_desc[slot]: contiguous ranges, 4- and 8-byte cp.async
  auto request_slices = [&](int tile, int slot, int b) {
    const int4 d0 = s_desc[2 * slot], d1 = s_desc[2 * slot + 1];
    const int32_t a0 = d0.x, eb = d0.z, nb = d1.x, q0 = d1.z;
    const int na = d0.y, ne = d0.w, nn = d1.y, nq = d1.w;
    unsigned char *in = in_buf(b);
    const int32_t n0 = tile * kTetStageNodes;
    const int nt = min(kTetStageNodes, n_owned - n0);
    for (int k = tid; k < nn; k += kTetStageThreads) cp_async_b4(in + L.nodes + 4 * k, tile_nodes + nb + k);
    for (int k = tid; k < ne; k += kTetStageThreads) cp_async_b8(in + L.erec + 8 * k, tile_erec + eb + k);
    if (mat_id)
      for (int k = tid; k < ne; k += kTetStageThreads) cp_async_b4(in + L.elist + 4 * k, tile_elist + eb + k);
    for (int k = tid; k <= na; k += kTetStageThreads) cp_async_b4(in + L.cptr + 4 * k, contrib_ptr + a0 + k);
    if (tid <= nt) cp_async_b4(in + L.aptr + 4 * tid, adj_ptr + n0 + tid);
    // 16-bit codes: copied in 4-byte words from the word that holds code q0 (index shifted by q0 & 1 at use)
    const int32_t qw = q0 >> 1;
    const int nw = ((q0 & 1) + nq + 1) >> 1;
    for (int k = tid; k < nw; k += kTetStageThreads)
      cp_async_b4(in + L.codes + 4 * k, reinterpret_cast<const uint32_t *>(contrib16) + qw + k);
    // k_self: bytes n0 .. n0 + nt (n0 is a multiple of 16: word aligned)
    if (tid < (nt + 3) / 4) cp_async_b4(in + L.kself + 4 * tid, reinterpret_cast<const uint32_t *>(kself + n0) + tid);
  };
  auto request_coords = [&](int slot, int b) {
    const int nn = s_desc[2 * slot + 1].y;
    unsigned char *in = in_buf(b);
    const int32_t *nodes = reinterpret_cast<const int32_t *>(in + L.nodes);
    for (int k = tid; k < 3 * nn; k += kTetStageThreads) {
      const int nd = k / 3;
      cp_async_b8(in + L.xyz + 8 * k, coords + 3 * (int64_t)nodes[nd] + (k - 3 * nd));
    }
  };
  auto request_desc = [&](int tile, int slot) {
    if (tile < n_tiles && tid < 2) cp_async_b16(s_desc + 2 * slot + tid, tile_desc + 2 * tile + tid);
  };

  // ---- prologue: descriptors of the first two tiles, slices and coordinates of the first
  int tile = blockIdx.x;
  if (tile >= n_tiles) return;
  request_desc(tile, 0);
  request_desc(tile + stride, 1);
  cp_async_commit_wait_all();
  __syncthreads();
  request_slices(tile, 0, 0);
  cp_async_commit_wait_all();
  __syncthreads();
  request_coords(0, 0);
  asm volatile("cp.async.commit_group;" ::: "memory");

  for (int it = 0; tile < n_tiles; tile += stride, ++it) {
    const int b = it & 1, slot = it % 3, slot1 = (it + 1) % 3, slot2 = (it + 2) % 3;
    const int next = tile + stride;
    // ---- top: this tile's coordinates (and, first trip, everything else) have landed
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const int4 d0 = s_desc[2 * slot], d1 = s_desc[2 * slot + 1];
    const int32_t a0 = d0.x, q0 = d1.z;
    const int na = d0.y, ne = d0.w;
    const int32_t n0 = tile * kTetStageNodes;
    const int nt = min(kTetStageNodes, n_owned - n0);
    if (next < n_tiles) request_slices(next, slot1, b ^ 1);
    request_desc(next + stride, slot2);
    asm volatile("cp.async.commit_group;" ::: "memory");
    const unsigned char *in = in_buf(b);
    const double *s_xyz = reinterpret_cast<const double *>(in + L.xyz);
    const ushort4 *s_erec = reinterpret_cast<const ushort4 *>(in + L.erec);
    const int32_t *s_elist = reinterpret_cast<const int32_t *>(in + L.elist);
    const int32_t *s_cptr = reinterpret_cast<const int32_t *>(in + L.cptr);
    const int32_t *s_aptr = reinterpret_cast<const int32_t *>(in + L.aptr);
    const uint16_t *s_codes = reinterpret_cast<const uint16_t *>(in + L.codes) + (q0 & 1);
    const uint8_t *s_kself = in + L.kself;
    // ---- phase A: the tile's elements, once each
    for (int le = tid; le < ne; le += kTetStageThreads) {
      const ushort4 r = s_erec[le];
      const double *p0 = s_xyz + 3 * r.x, *p1 = s_xyz + 3 * r.y, *p2 = s_xyz + 3 * r.z, *p3 = s_xyz + 3 * r.w;
      const TetGeom t = tet_geom_xyz(p0[0], p0[1], p0[2], p1[0], p1[1], p1[2], p2[0], p2[1], p2[2], p3[0], p3[1], p3[2]);
      const int mid = mat_id ? __ldg(mat_id + s_elist[le]) : 0;
      const TetMat m = tet_material(MASS ? FE_MASS_TET : FE_ELAST_TET, mat, mid, t.vol);
      const double sc = MASS ? 0.0 : sqrt(m.p1);
      double *g = s_grad + kTetGradStride * le;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        g[3 * k] = sc * t.g[k][0];
        g[3 * k + 1] = sc * t.g[k][1];
        g[3 * k + 2] = sc * t.g[k][2];
      }
      g[12] = MASS ? m.p0 : (m.p1 != 0.0 ? m.p0 / m.p1 : 0.0);
    }
    // ---- middle: the next tile's slices have landed (requested a whole phase ago); its coordinates go out
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (next < n_tiles) request_coords(slot1, b ^ 1);
    asm volatile("cp.async.commit_group;" ::: "memory");
    // ---- phase B: one lane per block
    const int nl = tid >> 4, lane = tid & 15;
    int al = 0, deg = 0;
    if (nl < nt) {
      al = s_aptr[nl] - a0;
      deg = s_aptr[nl + 1] - s_aptr[nl];
    }
    double *rows = s_out + 9 * al;
    for (int k = lane; k < deg; k += 16) {
      double acc[9];
#pragma unroll
      for (int q = 0; q < 9; ++q) acc[q] = 0.0;
      const int qb = s_cptr[al + k + 1] - q0;
      for (int q = s_cptr[al + k] - q0; q < qb; ++q) {
        const uint32_t code = s_codes[q];
        const double *ge = s_grad + kTetGradStride * (code >> 4);
        const double *gi = ge + 3 * ((code >> 2) & 3), *gj = ge + 3 * (code & 3);
        const D4 hi = {gi[0], gi[1], gi[2], ge[12]};
        D4 hj = hi;
        if (!MASS) hj = D4{gj[0], gj[1], gj[2], 0.0};
        tet_pair_add<MASS>(hi, hj, acc);
      }
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int t = 0; t < 3; ++t) rows[r * 3 * deg + 3 * k + t] = acc[3 * r + t];
    }
    __syncwarp();
    if (lane < 9 && deg > 0) {  // diagonal block from the off-diagonal ones, slot order
      const int ks = s_kself[nl];
      const int r = lane / 3, t = lane - 3 * r;
      const double *row = rows + r * 3 * deg + t;
      double sum = 0.0;
      for (int k = 0; k < deg; ++k)
        if (k != ks) sum += row[3 * k];
      rows[r * 3 * deg + 3 * ks + t] = (MASS ? (2.0 / 3.0) : -1.0) * sum;
    }
    __syncthreads();
    // ---- store the image of vals[9 a0 .. 9 (a0 + na))
    double *dst = vals + 9 * (int64_t)a0;
    for (int k = tid; k < 9 * na; k += kTetStageThreads) dst[k] = s_out[k];
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

}  // namespace fe

using namespace fe;

extern "C" {

int fe_tet_elem_matrices(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords,
                         const int32_t *conn, const int32_t *mat_id, const double *mat, int32_t n_mat, double *out) {
  FE_REQUIRE(ctx && coords && conn && mat && out, "fe_tet_elem_matrices: NULL argument");
  FE_REQUIRE(kind == FE_ELAST_TET || kind == FE_MASS_TET, "fe_tet_elem_matrices: kind %d is not a tetrahedral kind", kind);
  FE_REQUIRE(n_elems >= 0 && n_mat > 0, "fe_tet_elem_matrices: bad sizes");
  FE_REQUIRE(((uintptr_t)conn & 15) == 0, "fe_tet_elem_matrices: conn must be 16-byte aligned");
  if (n_elems == 0) return FE_OK;
  k_tet_elem_matrices<<<grid_for(n_elems, 128), 128, 0, as_stream(stream)>>>(kind, n_elems, coords, conn, mat_id, mat,
                                                                           out);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_tet_elem_post(fe_ctx *ctx, void *stream, int64_t n_elems, const double *coords, const int32_t *conn,
                     const int32_t *mat_id, const double *mat, int32_t n_mat, const double *u, double *out) {
  FE_REQUIRE(ctx && coords && conn && mat && u && out, "fe_tet_elem_post: NULL argument");
  FE_REQUIRE(n_elems >= 0 && n_mat > 0, "fe_tet_elem_post: bad sizes");
  FE_REQUIRE(((uintptr_t)conn & 15) == 0, "fe_tet_elem_post: conn must be 16-byte aligned");
  if (n_elems == 0) return FE_OK;
  k_tet_post<<<grid_for(n_elems, 128), 128, 0, as_stream(stream)>>>(n_elems, coords, conn, mat_id, mat, u, out);
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

int fe_tet_assemble(fe_ctx *ctx, void *stream, const fe_plan *p, int kind, const double *coords, const int32_t *conn,
                    const int32_t *mat_id, const double *mat, int32_t n_mat, double *vals, int32_t variant) {
  FE_REQUIRE(ctx && p && coords && conn && mat && (vals || p->nnz == 0), "fe_tet_assemble: NULL argument");
  FE_REQUIRE(p->npe == 4 && p->dim == 3, "fe_tet_assemble: the plan was not built by fe_tet_plan_create");
  FE_REQUIRE(kind == FE_ELAST_TET || kind == FE_MASS_TET, "fe_tet_assemble: kind %d is not a tetrahedral kind", kind);
  FE_REQUIRE(n_mat > 0, "fe_tet_assemble: bad sizes");
  FE_REQUIRE(((uintptr_t)conn & 15) == 0, "fe_tet_assemble: conn must be 16-byte aligned");
  FE_REQUIRE(variant >= 0 && variant <= 6, "fe_tet_assemble: unknown variant %d", variant);
  if (p->n_owned == 0 || p->nnz == 0) return FE_OK;
  cudaStream_t st = as_stream(stream);
  const int max_degree = p->max_degree;
  const size_t smem = (size_t)9 * (max_degree > 0 ? max_degree : 1) * kTetLD * sizeof(double);
  const bool fits = max_degree > 0 && smem <= 200 * 1024;
  if (variant >= 3 && p->tet_degenerate)
    return fail(FE_ERR_UNSUPPORTED, "fe_tet_assemble: the slot variant needs elements with four distinct nodes");
  if (variant == 2 && !fits)
    return fail(FE_ERR_UNSUPPORTED, "fe_tet_assemble: the tile variant needs %zu B of shared memory (valence %d)", smem, max_degree);
  const TetStageSmem sl = tet_stage_smem(p->tile_elems_max, p->tile_adj_max, p->tile_nodes_max, p->tile_contrib_max);
  const TetPipeSmem pl = tet_pipe_smem(p->tile_elems_max, p->tile_adj_max, p->tile_nodes_max, p->tile_contrib_max);
  const bool stage_fits = p->tet_stage_ok && sl.total <= 200 * 1024;
  const bool pipe_fits = p->tet_stage_ok && pl.total <= 200 * 1024;
  if ((variant == 5 && !stage_fits) || (variant == 6 && !pipe_fits))
    return fail(FE_ERR_UNSUPPORTED, "fe_tet_assemble: the staged variants need tiles of <= 4095 elements within %d B of shared memory",
                200 * 1024);
  if (variant == 0) variant = pipe_fits ? 6 : (stage_fits ? 5 : (!p->tet_degenerate ? 4 : (fits ? 2 : 1)));
  if (variant == 6) {
#define FE_TET_PIPE(MASS)                                                                                            \
  do {                                                                                                               \
    int nb = 1;                                                                                                      \
    FE_CUDA(cudaFuncSetAttribute(k_tet_assemble_pipe<MASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.total)); \
    FE_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_tet_assemble_pipe<MASS>, kTetStageThreads, pl.total)); \
    if (nb < 1) nb = 1;                                                                                              \
    const int grid = p->n_tiles < nb * ctx->num_sms ? p->n_tiles : nb * ctx->num_sms;                                \
    k_tet_assemble_pipe<MASS><<<grid, kTetStageThreads, pl.total, st>>>(                                              \
        p->n_owned, p->n_tiles, p->adj_ptr, p->contrib_ptr, p->contrib16, p->tile_desc, p->tile_elist, p->tile_erec, \
        p->tile_nodes, p->tet_kself, coords, mat_id, mat, vals, p->tile_elems_max, p->tile_adj_max,                  \
        p->tile_nodes_max, p->tile_contrib_max);                                                                     \
  } while (0)
    if (kind == FE_MASS_TET)
      FE_TET_PIPE(true);
    else
      FE_TET_PIPE(false);
#undef FE_TET_PIPE
  } else
  if (variant == 5) {
#define FE_TET_STAGED(MASS)                                                                                          \
  do {                                                                                                               \
    FE_CUDA(cudaFuncSetAttribute(k_tet_assemble_staged<MASS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sl.total)); \
    k_tet_assemble_staged<MASS><<<p->n_tiles, kTetStageThreads, sl.total, st>>>(                                      \
        p->n_owned, p->adj_ptr, p->contrib_ptr, p->contrib16, p->tile_desc, p->tile_elist, p->tile_erec,             \
        p->tile_nodes, p->tet_kself, coords, mat_id, mat, vals, p->tile_elems_max, p->tile_adj_max, p->tile_nodes_max, \
        p->tile_contrib_max);                                                                                        \
  } while (0)
    if (kind == FE_MASS_TET)
      FE_TET_STAGED(true);
    else
      FE_TET_STAGED(false);
#undef FE_TET_STAGED
  } else if (variant == 4) {
    int rc = ctx->scratch_g.reserve((size_t)p->n_elems * 16 * sizeof(double));
    if (rc) return rc;
    double *table = (double *)ctx->scratch_g.ptr;
    const int g1 = grid_for(p->n_elems, 128), g2 = grid_for((int64_t)p->n_owned * 16, 256);
    if (kind == FE_MASS_TET) {
      k_tet_gradient_table<true><<<g1, 128, 0, st>>>(p->n_elems, coords, conn, mat_id, mat, table);
      FE_LAUNCH_CHECK(ctx);
      k_tet_assemble_table<true><<<g2, 256, 0, st>>>(p->n_owned, p->adj_ptr, p->adj, p->contrib_ptr, p->contrib, table, vals);
    } else {
      k_tet_gradient_table<false><<<g1, 128, 0, st>>>(p->n_elems, coords, conn, mat_id, mat, table);
      FE_LAUNCH_CHECK(ctx);
      k_tet_assemble_table<false><<<g2, 256, 0, st>>>(p->n_owned, p->adj_ptr, p->adj, p->contrib_ptr, p->contrib, table, vals);
    }
  } else if (variant == 3) {
    const int grid = grid_for((int64_t)p->n_owned * 16, 256);
    if (kind == FE_MASS_TET)
      k_tet_assemble_slots<true><<<grid, 256, 0, st>>>(p->n_owned, p->adj_ptr, p->adj, p->contrib_ptr, p->contrib, coords, conn,
                                                      mat_id, mat, vals);
    else
      k_tet_assemble_slots<false><<<grid, 256, 0, st>>>(p->n_owned, p->adj_ptr, p->adj, p->contrib_ptr, p->contrib, coords, conn,
                                                       mat_id, mat, vals);
  } else if (variant == 2) {
    FE_CUDA(cudaFuncSetAttribute(k_tet_assemble_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_tet_assemble_tile<<<grid_for(p->n_owned, kTetTile), kTetTile, smem, st>>>(
        kind, p->n_owned, p->corner_ptr, p->corner_elem, p->adj_ptr, p->adj, coords, conn, mat_id, mat, vals);
  } else {
    k_tet_assemble<<<grid_for(p->n_owned, 128), 128, 0, st>>>(kind, p->n_owned, p->corner_ptr, p->corner_elem, p->adj_ptr,
                                                             p->adj, coords, conn, mat_id, mat, vals);
  }
  FE_LAUNCH_CHECK(ctx);
  return FE_OK;
}

}  // extern "C"
