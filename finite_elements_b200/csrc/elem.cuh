// Per-element FP64 arithmetic for linear triangles, shared by the element-dump kernel
// (fe_elem_matrices) and the row-owner assembly kernels (fe_assemble) so that both
// produce bit-identical entries.
//
// Reference formulas (paths under /root/reference/finite_elements/):
//   beta/gamma/detJ, B = (1/detJ) [...]              elements.py:395-416
//   D plane strain / plane stress                    elements.py:418-453
//   Ke = thickness * area * B^T D B                  elements.py:466-511
//   magnetic Ke_ij = (1/mu)(b_i b_j + c_i c_j) area  elements.py:93-118
//   mass (rho area t / 12) [[2,1,1],..] (x) I2       elements.py:513-536
//   area = |u x v| / 2 (volmdlr TriangularElement2D) -- third-party, restated
#pragma once
#include "common.cuh"

namespace fe {

// Per-material constants precomputed once per call (k_material_table).
//   elasticity: (c t, a t, b t, t) with D = [[c,a,0],[a,c,0],[0,0,b]] and t = thickness
//   magnetic  : (1/mu, 0, 0, 0)
//   mass      : (rho * thickness / 12, 0, 0, 0)
struct MatRow {
  double p0, p1, p2, p3;
};

// core.cu: fills `where` with n_mat MatRow entries for `kind` (one tiny launch).
int build_material_table(fe_ctx *ctx, cudaStream_t st, int kind, const double *mat, int n_mat, MatRow **tab_out,
                         Scratch *where);

struct TriGeom {
  double beta[3];   // y1-y2, y2-y0, y0-y1            (elements.py:403)
  double gamma[3];  // x2-x1, x0-x2, x1-x0            (elements.py:404)
  double det;       // (x0-x2)(y1-y2)-(y0-y2)(x1-x2)  (elements.py:406-408), signed
  double cross;     // (x1-x0)(y2-y0)-(y1-y0)(x2-x0)  volmdlr area / form-function divisor
};

__device__ __forceinline__ TriGeom tri_geom(double2 p0, double2 p1, double2 p2) {
  TriGeom g;
  g.beta[0] = p1.y - p2.y;
  g.beta[1] = p2.y - p0.y;
  g.beta[2] = p0.y - p1.y;
  g.gamma[0] = p2.x - p1.x;
  g.gamma[1] = p0.x - p2.x;
  g.gamma[2] = p1.x - p0.x;
  // (x0-x2) = gamma1, (y1-y2) = beta0, (y0-y2) = -beta1, (x1-x2) = -gamma0
  g.det = g.gamma[1] * g.beta[0] - g.beta[1] * g.gamma[0];
  // (x1-x0) = gamma2, (y2-y0) = beta1, (y1-y0) = -beta2, (x2-x0) = -gamma1
  g.cross = g.gamma[2] * g.beta[1] - g.beta[2] * g.gamma[1];
  return g;
}

// Row-block v (local vertex v) of the 6x6 elasticity Ke: out[j] = 2x2 block (v, j) as
// (k00, k01, k10, k11).
struct Blk2 {
  double k00, k01, k10, k11;
};

__device__ __forceinline__ void elast_row_blocks(const TriGeom &g, const MatRow &m, int v, Blk2 out[3]) {
  // Ke(v, j) = t A B_v^T D B_j with B_i = (1/det) [[beta_i, 0], [0, gamma_i], [gamma_i, beta_i]].
  // The two 1/det factors and t*A are folded into the row-v multipliers so that the column
  // side uses the raw beta_j / gamma_j (saves the six per-vertex scalings).
  const double inv = 1.0 / g.det;
  const double s = (0.5 * fabs(g.cross)) * inv * inv;  // area / det^2   (thickness is inside m)
  // select row v without dynamic register indexing
  const double bv = (v == 0) ? g.beta[0] : ((v == 1) ? g.beta[1] : g.beta[2]);
  const double gv = (v == 0) ? g.gamma[0] : ((v == 1) ? g.gamma[1] : g.gamma[2]);
  const double tb = s * bv, tg = s * gv;
  const double cb = m.p0 * tb, cg = m.p0 * tg;  // c t
  const double ab = m.p1 * tb, ag = m.p1 * tg;  // a t
  const double sb = m.p2 * tb, sg = m.p2 * tg;  // b t (shear)
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    out[j].k00 = cb * g.beta[j] + sg * g.gamma[j];
    out[j].k01 = ab * g.gamma[j] + sg * g.beta[j];
    out[j].k10 = ag * g.beta[j] + sb * g.gamma[j];
    out[j].k11 = cg * g.gamma[j] + sb * g.beta[j];
  }
}

__device__ __forceinline__ void mass_row_blocks(const TriGeom &g, const MatRow &m, int v, Blk2 out[3]) {
  const double area = 0.5 * fabs(g.cross);
  const double s = m.p0 * area;  // rho * t / 12 * area
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const double d = (j == v) ? 2.0 * s : s;
    out[j].k00 = d;
    out[j].k01 = 0.0;
    out[j].k10 = 0.0;
    out[j].k11 = d;
  }
}

// Row v of the 3x3 magnetic Ke.
__device__ __forceinline__ void mag_row(const TriGeom &g, const MatRow &m, int v, double out[3]) {
  const double inv = 1.0 / g.cross;
  const double area = 0.5 * fabs(g.cross);
  const double s = m.p0 * area;  // (1/mu) * area
  double b[3], c[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    b[j] = g.beta[j] * inv;
    c[j] = g.gamma[j] * inv;
  }
  const double bv = (v == 0) ? b[0] : ((v == 1) ? b[1] : b[2]);
  const double cv = (v == 0) ? c[0] : ((v == 1) ? c[1] : c[2]);
  const double sbv = s * bv, scv = s * cv;
#pragma unroll
  for (int j = 0; j < 3; ++j) out[j] = sbv * b[j] + scv * c[j];
}

}  // namespace fe
