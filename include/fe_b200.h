/*
 * fe_b200.h -- C ABI of libfe_b200.so: the B200 (sm_100a) implementation of the
 * data-parallel hot path of Dessia-tech/finite_elements v0.2.0.
 *
 * The reference is pure Python and has no FFI of its own: the
 * seam is the Python class surface of finite_elements/analysis.py.  Each entry
 * point below names the reference code (file:line under
 * /root/reference/finite_elements/) whose work it replaces; the Python host
 * layer (finite_elements_b200/analysis.py) keeps the reference's class names and
 * calls these through ctypes.  INTEGRATION.md shows the binding a maintainer of
 * the reference would add.
 *
 * Conventions
 *  - Every array pointer is a CUDA DEVICE pointer owned by the caller (torch
 *    tensors' data_ptr()).  The library never frees or keeps caller buffers
 *    beyond the call.  The library owns only fe_ctx (scratch, NCCL comm) and
 *    fe_plan (the per-mesh symbolic data built once by fe_plan_create).
 *  - `stream` is a cudaStream_t passed as void* (torch.cuda.current_stream()
 *    .cuda_stream).  Calls are asynchronous on that stream unless stated.
 *  - Return value: 0 = FE_OK, negative = error class; text via fe_last_error()
 *    (thread-local).  No global mutable state; one ctx per (thread, device).
 *  - Indices are int32 (all sizes on the path fit; fe_plan_create fails with
 *    FE_ERR_UNSUPPORTED when nnz would overflow int32).  Values are FP64.
 *  - DOF numbering: dof = node*dim + d  (core.py:89-108).
 *  - conn is int32[E][3] in the reference's local order points[0..2];
 *    coords is double[N][2]; mat_id int32[E] (may be NULL = all 0);
 *    mat is double[G][4]: elasticity rows (E_modulus, poisson, thickness,
 *    mass_density), magnetic rows (mu_total, -, -, -).
 */
#ifndef FE_B200_H
#define FE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FE_B200_VERSION 100 /* 0.1.0 */

typedef struct fe_ctx fe_ctx;
typedef struct fe_plan fe_plan;

enum fe_status {
  FE_OK = 0,
  FE_ERR_ARG = -1,           /* bad argument (maps to ValueError)                       */
  FE_ERR_CUDA = -2,          /* CUDA runtime error (RuntimeError)                       */
  FE_ERR_NCCL = -3,          /* NCCL error (RuntimeError)                               */
  FE_ERR_NOT_CONVERGED = -4, /* PCG hit maxit (results still written)                   */
  FE_ERR_BREAKDOWN = -5,     /* PCG: p.Ap <= 0 or non-finite (matrix not SPD / singular;
                                 the reference raises NotImplementedError from
                                 MatrixRankWarning, analysis.py:824-826)                */
  FE_ERR_UNSUPPORTED = -6    /* size/valence outside the implemented range              */
};

/* elements.py:466-511 (plane stress / plane strain via :251-273), :93-118, :513-536 */
enum fe_kind {
  FE_ELAST_PSTRESS = 0, /* ElasticityTriangularElement2D.elementary_matrix(False, True) */
  FE_ELAST_PSTRAIN = 1, /* ... (True, False)                                            */
  FE_MAGNETIC = 2,      /* MagneticElement2D.elementary_matrix()                        */
  FE_MASS = 3,          /* ElasticityTriangularElement2D.elementary_mass_matrix()       */
  /* linear tetrahedra, 3 DOF per node (elements.py:663-876); fe_tet_* entry points only */
  FE_ELAST_TET = 4,     /* ElasticityTetrahedralElement3D.elementary_matrix(.., ..)     */
  FE_MASS_TET = 5       /* ElasticityTetrahedralElement3D.elementary_mass_matrix()      */
};

int fe_version(void);
const char *fe_last_error(void);

int fe_ctx_create(int device, fe_ctx **out);
int fe_ctx_destroy(fe_ctx *ctx);

/* ---- per-element kernels (parity / debug / host-side element API) -------------------
 * replaces elements.py:395-416 (_b_matrix), :418-453 (D), :466-511 / :93-118 / :513-536.
 * out: double[E][(3*dim)^2] row-major, local DOF order [u0,v0,u1,v1,u2,v2] (or [n0,n1,n2]). */
int fe_elem_matrices(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords,
                     const int32_t *conn, const int32_t *mat_id, const double *mat, int32_t n_mat,
                     double *out);

/* replaces elements.py:18-53 / :156-191 (element_to_node_factors) for the elements listed
 * in elem_sel (int32[n_sel], NULL = elements 0..n_sel-1).  out: double[n_sel][3].
 * out_area (may be NULL): double[n_sel] = volmdlr TriangularElement2D.area (loads.py:31-34). */
int fe_source_factors(fe_ctx *ctx, void *stream, int64_t n_sel, const int32_t *elem_sel,
                      const double *coords, const int32_t *conn, double *out, double *out_area);

/* Element post-processing of a solution vector u (double[n_nodes*dim], device):
 *   elasticity kinds: out double[E][7] = (eps_xx, eps_yy, gamma_xy, sig_xx, sig_yy, tau_xy, energy)
 *     replaces results.py:809-830 (strain = B u_e, stress = D B u_e) and results.py:769-781 /
 *     elements.py:275-292 (energy = 1/2 u_e^T Ke u_e);
 *   FE_MAGNETIC: out double[E][2] = (B_x, B_y) = (sum c_i A_i, -sum b_i A_i), results.py:121-152. */
int fe_elem_post(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords,
                 const int32_t *conn, const int32_t *mat_id, const double *mat, int32_t n_mat,
                 const double *u, double *out);

/* ---- symbolic phase, once per mesh ----------------------------------------------------
 * replaces analysis.py:714-735 (get_row_col_indices) for all elements plus the
 * COO->CSR sort/unique scipy does at analysis.py:661 (pattern part).
 * Builds, for rows of nodes [0, n_owned_nodes): the node->corner incidence (counting sort
 * keyed by node, corners ordered by element id), the sorted unique node adjacency, and the
 * corner->CSR-slot map used by fe_assemble.  n_owned_nodes == n_nodes on one GPU; on a
 * rank of a partition, nodes are numbered owned-first and ghosts after, and `conn` lists
 * every element incident to an owned node.  SYNCHRONISES the stream.
 * Limits (FE_ERR_UNSUPPORTED beyond them): a node may have at most 255 neighbours incl. itself
 * (slot positions are stored in 8 bits); E < 2^29; n_nodes*dim and nnz < 2^31.  The fan-ordered
 * assembly variant additionally needs <= 32 elements per node, material ids < 2^19 and node stars
 * that are simple fans -- meshes outside that use variant 2 / 1 automatically (same results). */
int fe_plan_create(fe_ctx *ctx, void *stream, int32_t n_nodes, int32_t n_owned_nodes,
                   int64_t n_elems, int32_t dim, const int32_t *conn, const int32_t *mat_id,
                   fe_plan **out);
int fe_plan_destroy(fe_plan *plan);
int64_t fe_plan_nnz(const fe_plan *plan);        /* nnz of the (n_owned*dim) x (n_nodes*dim) CSR */
int32_t fe_plan_n_rows(const fe_plan *plan);     /* n_owned_nodes * dim                         */
int32_t fe_plan_max_degree(const fe_plan *plan); /* max node valence incl. self                  */
int64_t fe_plan_bytes(const fe_plan *plan);      /* device bytes held by the plan                */
/* Bytes per fan record fe_assemble's default variant will read: 4 (compact records: banded numbering,
 * |neighbour - node| < 2^17 on the owned block and < 2^18 ghost columns, at most two materials with ids
 * < 4096 around any node), 8 otherwise, 0 when the mesh has no fan ordering (variants 2 / 1). */
int32_t fe_plan_fan_record_bytes(const fe_plan *plan);
/* Canonical CSR pattern (sorted columns, explicit zeros kept): bit-exact with
 * scipy.sparse.csr_matrix((data,(row,col))) of analysis.py:661 on the K block.
 * rowptr int32[n_rows+1], colidx int32[nnz]. */
int fe_plan_csr(const fe_plan *plan, void *stream, int32_t *rowptr, int32_t *colidx);

/* ---- numeric assembly ----------------------------------------------------------------
 * replaces analysis.py:324-339 (k_matrix_data) / :357-365 (m_matrix_data) and the
 * duplicate summation of analysis.py:661 / :387-405.  Deterministic: every CSR slot is
 * summed by one thread in a fixed (element id) order, no atomics.  vals double[nnz] is
 * fully overwritten.  variant: 0 = default (fastest available), 1 = generic row-owner
 * kernel (accumulates in global memory), 2 = shared-memory staged tiles, 3 = fan-ordered
 * traversal + staged tiles (needs node stars that are simple fans, i.e. no edge shared by
 * more than two elements; FE_ERR_UNSUPPORTED otherwise -- variant 0 then picks 2), 4 = variant 3
 * forced onto the 8-byte fan records (bit-identical to 3; for tests and A/B timing). */
int fe_assemble(fe_ctx *ctx, void *stream, const fe_plan *plan, int kind, const double *coords,
                const double *mat, int32_t n_mat, double *vals, int variant);

/* ---- boundary conditions --------------------------------------------------------------
 * The reference appends Lagrange rows (analysis.py:241-279, :509-543).  On the solution
 * block that system is equivalent to symmetric elimination, done here in
 * place on the caller's CSR: rhs -= K[:,c] g, row/col c zeroed, K[c,c] = 1, rhs[c] = g.
 * bc_dof int32[n_bc] (unique; local column numbering), bc_val double[n_bc].
 * n_rows x n_cols CSR (n_cols >= n_rows; columns >= n_rows are ghost columns), columns sorted within a
 * row and the pattern structurally symmetric on the owned block -- what fe_plan_csr produces and every
 * finite-element matrix has; only the rows next to a condition are visited (O(n_bc * valence)).
 * FE_B200_BC_SWEEP=1 (environment) selects a sweep over all rows that needs neither property. */
int fe_dirichlet_apply(fe_ctx *ctx, void *stream, int32_t n_rows, int32_t n_cols,
                       const int32_t *rowptr, const int32_t *colidx, double *vals, double *rhs,
                       int32_t n_bc, const int32_t *bc_dof, const double *bc_val);

/* rhs[dof[i]] += val[i] (analysis.py:698-702); dof unique => deterministic. */
int fe_scatter_add(fe_ctx *ctx, void *stream, int32_t n, const int32_t *dof, const double *val,
                   double *rhs);

/* ---- solve ---------------------------------------------------------------------------
 * block_dim (all solve entry points): 1 = arbitrary CSR; 2 = the CSR was produced by an
 * fe_plan with dim == 2 (rows 2i and 2i+1 share one column list made of (2m, 2m+1) pairs;
 * fe_dirichlet_apply keeps that structure), which lets the SpMV read ONE column index per
 * 2x2 block and use 128-bit loads.  Passing 2 for a matrix without that structure is an error
 * the library cannot detect.  3 = the CSR was produced by fe_tet_plan_create (rows 3i .. 3i+2 share one
 * column list of (3m, 3m+1, 3m+2) triples): the PCG's SpMV reads one column index per 3x3 block; this
 * structure IS verified when the solver builds its node-level pattern (a CSR without it takes the scalar path).
 *
 * y = A x, CSR, sub-warp-per-row (block_dim 1) or 8-lanes-per-node (block_dim 2). */
int fe_spmv(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr,
            const int32_t *colidx, const double *vals, const double *x, double *y,
            int32_t block_dim);

/* ---- linear tetrahedra --------------------------------------------
 * coords double[N][3], conn int32[E][4] (16-byte aligned), mat rows as above (thickness unused).
 * out: double[E][144], row-major 12x12 in the reference's DOF order [u0,v0,w0,u1,...]
 * (elements.py:809-828 / :830-857). */
int fe_tet_elem_matrices(fe_ctx *ctx, void *stream, int kind, int64_t n_elems, const double *coords,
                         const int32_t *conn, const int32_t *mat_id, const double *mat,
                         int32_t n_mat, double *out);
/* Element post-processing of a solution u (results.py:809-830, :769-781): out double[E][13] =
 * (eps_xx, eps_yy, eps_zz, gamma_xy, gamma_yz, gamma_zx, the six stresses in the same order, energy). */
int fe_tet_elem_post(fe_ctx *ctx, void *stream, int64_t n_elems, const double *coords,
                     const int32_t *conn, const int32_t *mat_id, const double *mat, int32_t n_mat,
                     const double *u, double *out);
/* Symbolic phase of a tetrahedral mesh, once per mesh, entirely on the device (replaces
 * get_row_col_indices analysis.py:714-735 for 12 x 12 element matrices and the pattern half of csr_matrix,
 * :661): node->element lists (counting sort keyed by node, ascending element id), sorted node adjacency,
 * and -- per off-diagonal block (node, neighbour) -- the list of elements that hold both nodes.  The plan
 * works with fe_plan_nnz / _n_rows / _max_degree / _bytes / _csr / _destroy like a triangle plan (dim = 3:
 * node i's rows 3i, 3i+1, 3i+2 lie back to back from 9 adj_ptr[i], each 3 deg_i long, the block of
 * neighbour slot k at columns 3k..3k+2 -- the scipy-canonical CSR of the same triplets).  n_owned_nodes <
 * n_nodes selects the multi-GPU layout (owned nodes first).  conn int32[E][4], 16-byte aligned.  Same
 * limits as fe_plan_create (<= 255 neighbours per node), E < 2^27.  SYNCHRONISES the stream. */
int fe_tet_plan_create(fe_ctx *ctx, void *stream, int32_t n_nodes, int32_t n_owned_nodes,
                       int64_t n_elems, const int32_t *conn, fe_plan **out);
/* Global matrix values of the tetrahedral mesh; replaces the k_matrix_data / m_matrix_data loops
 * (analysis.py:324-339, :357-365) + csr_matrix (:661) for 3 DOF per node.  Deterministic, no atomics.
 * variant 0 = automatic; 6 = staged tiles, pipelined (default): per tile of 16 consecutive nodes a CTA evaluates
 * the elements its rows touch once into shared memory and one lane per 3x3 block walks the plan's per-block
 * element lists there, while the next tile's inputs arrive by cp.async (needs elements with four distinct
 * nodes, <= 4095 elements per tile and the tile within 200 KB of shared memory); 5 = the same, one tile per
 * CTA without the prefetch (bit-identical); 4 = the same walk over a global gradient table written by a first pass (128 B per element
 * of ctx scratch; the fallback of 6 / 5); 3 = the same walk rebuilding the
 * geometry at every visit (no scratch); 1 = one thread per node accumulating in global memory, 2 = the
 * same through a shared-memory tile (1 and 2 are bit-identical to each other and add a block's elements
 * in ascending order; 3, 4, 5 and 6 agree with them to rounding). */
int fe_tet_assemble(fe_ctx *ctx, void *stream, const fe_plan *plan, int kind, const double *coords,
                    const int32_t *conn, const int32_t *mat_id, const double *mat, int32_t n_mat,
                    double *vals, int32_t variant);

/* ---- modal analysis building blocks (analysis.py:741-796) -------------------------------
 * The reference gives K and M to scipy.sparse.linalg.eigsh (:779-782), whose Lanczos loop
 * multiplies one vector at a time; the LOBPCG driver of the host layer works on blocks of m
 * vectors.  y_a = A x and (when vals_b / y_b are non-NULL) y_b = B x in one pass over the
 * pattern both matrices share.  x: double[n_cols][m] row-major, y_*: double[n_rows][m].
 * block_dim as in the solve entry points below: 2 lets a group of lanes own both rows of a node and fetch every
 * row of x once for the two of them (bit-identical results); 1 (or 3) = row by row. */
int fe_spmm_pair(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr,
                 const int32_t *colidx, const double *vals_a, const double *vals_b,
                 const double *x, double *y_a, double *y_b, int32_t m, int32_t block_dim);
/* One step of the Chebyshev iteration that preconditions the block eigensolver, fused into the
 * block product (all of it is row-local once y = A d_in is known):
 *   z += d_in;  r -= A d_in;  d_out = c1 d_in + c2 diag(dinv) r          (blocks are [n][m])
 * d_out must not alias d_in. */
int fe_cheb_step(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr,
                 const int32_t *colidx, const double *vals, const double *dinv, const double *d_in,
                 double *d_out, double *r, double *z, double c1, double c2, int32_t m, int32_t block_dim);
/* diag[i] = A[i][i] (0 if the entry is not stored): the Jacobi preconditioner of the block solver */
int fe_csr_diagonal(fe_ctx *ctx, void *stream, int32_t n_rows, const int32_t *rowptr,
                    const int32_t *colidx, const double *vals, double *diag);

/* Jacobi-preconditioned CG; replaces scipy spsolve at analysis.py:820-822 on the
 * eliminated SPD system.  x: initial guess in, solution out.  work: double[fe_pcg_work_len(n)].
 * Stops when ||r||_2 <= rtol * ||b||_2, where r is re-computed as b - A x once the recurrence
 * signals convergence (restart from x if the recurrence had drifted).  If restarts stop
 * reducing the true residual (attainable FP64 accuracy reached) the call returns FE_OK only when
 * relres <= max(100 rtol, 1e-8); otherwise FE_ERR_NOT_CONVERGED (x, iters, relres still written).
 * SYNCHRONISES the stream; writes iters / relres (the true relative residual). */
int64_t fe_pcg_work_len(int32_t n_rows, int32_t n_cols);
int fe_pcg(fe_ctx *ctx, void *stream, int32_t n, const int32_t *rowptr, const int32_t *colidx,
           const double *vals, const double *b, double *x, double *work, int32_t block_dim,
           double rtol, int32_t maxit, int32_t *iters, double *relres);
/* Runs exactly `iters` PCG iterations with no convergence test (throughput measurement). */
int fe_pcg_fixed(fe_ctx *ctx, void *stream, int32_t n, const int32_t *rowptr,
                 const int32_t *colidx, const double *vals, const double *b, double *x,
                 double *work, int32_t block_dim, int32_t iters);
/* Optional: the caller vouches that the CSR pattern stored at (rowptr, colidx) does not change
 * while it keeps passing the same non-zero `token` (the host layer uses one token per mesh).  The
 * solver then builds its node-level block pattern once instead of once per solve -- the analogue
 * of the reference caching `_boundary_conditions` / `_positions` on the analysis object
 * (analysis.py:169-171).  (NULL, NULL, 0) clears the hint. */
int fe_pcg_cache_pattern(fe_ctx *ctx, const int32_t *rowptr, const int32_t *colidx, int64_t token);

/* ---- multi-GPU (one process per GPU) --------------------------------------
 * nccl_unique_id: 128 bytes from ncclGetUniqueId on rank 0 (fe_dist_unique_id), broadcast
 * by the host (torch.distributed). */
int fe_dist_unique_id(void *out128);
int fe_dist_init(fe_ctx *ctx, const void *nccl_unique_id, int32_t rank, int32_t nranks);
/* Halo description (device int32 arrays unless noted):
 *  n_nbr neighbours; nbr_rank (HOST int32[n_nbr]);
 *  send_ptr (HOST int32[n_nbr+1]) into send_idx (device; owned local DOFs to pack);
 *  recv_ptr (HOST int32[n_nbr+1]): ghost DOFs received from neighbour k occupy local columns
 *  n_rows + recv_ptr[k] .. n_rows + recv_ptr[k+1].
 * work: double[fe_pcg_work_len(n_rows, n_cols)]. */
int fe_dist_pcg(fe_ctx *ctx, void *stream, int32_t n_rows, int32_t n_cols, const int32_t *rowptr,
                const int32_t *colidx, const double *vals, const double *b, double *x,
                double *work, int32_t n_nbr, const int32_t *nbr_rank, const int32_t *send_ptr,
                const int32_t *send_idx, const int32_t *recv_ptr, const int32_t *peer_dst_off,
                int32_t block_dim, double rtol, int32_t maxit, int32_t fixed_iters, int32_t *iters,
                double *relres);

/* Peer-memory transport for fe_dist_pcg (one node, NVLink / NVSwitch): every rank exports one
 * communication block (all-reduce slots + flags, ghost values) with CUDA IPC, the host gathers the
 * 64-byte handles of all ranks (rank order) and every rank imports them.  Afterwards fe_dist_pcg
 * -- given peer_dst_off (HOST int32[n_nbr]: where this rank's interface values start inside
 * neighbour k's ghost block, i.e. that neighbour's recv_ptr entry for this rank) -- runs without
 * any NCCL call in the iteration: halo values are stored straight into the neighbours' memory and
 * the dot products are all-reduced through peer-written slots inside the PCG kernels themselves.
 * n_ghost_dofs = this rank's ghost count (n_cols - n_rows).  Both calls are collective. */
int fe_dist_p2p_export(fe_ctx *ctx, int32_t n_ghost_dofs, void *handle64_out);
int fe_dist_p2p_import(fe_ctx *ctx, const void *handles /* nranks * 64 bytes */);

/* Number of kernel launches issued by this ctx since creation (bench.py's gpu_launches). */
int64_t fe_ctx_launch_count(const fe_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FE_B200_H */
