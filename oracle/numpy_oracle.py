"""TEST INFRASTRUCTURE -- the CPU oracle.  Not product code.

A vectorised numpy/scipy RESTATEMENT of the reference's hot path
(Dessia-tech/finite_elements v0.2.0).  Only tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs may import this module; the
product (finite_elements_b200/) never does.

Pinning: the reference ships no asserting tests for this path, so
the oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF: tests/golden/*.npz
are produced by oracle/make_golden.py, which runs the reference's unmodified
modules (oracle/ref_loader.py) in the build container.  tests/test_oracle.py
checks every function below against those fixtures.

Every function cites the reference file:line it follows (paths relative to
/root/reference/finite_elements/).

Array conventions (shared with the product's flat problem description):
  coords f64[N,2]; conn i32[E,3] (reference local order points[0..2]);
  mat_id i32[E]; mat f64[G,4]:
     elasticity rows = (E_modulus, poisson, thickness, mass_density)
     magnetic   rows = (mu_total, 0, 0, 0)
  DOF numbering: dof = node*dim + d   (core.py:89-108, d zero-based here,
  `dimension` in load/BC records is 1-based as in the reference API).
"""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

KIND_ELAST_PSTRESS = 0
KIND_ELAST_PSTRAIN = 1
KIND_MAGNETIC = 2
KIND_MASS = 3
KIND_ELAST_TET = 4   # ElasticityTetrahedralElement3D.elementary_matrix (3 DOF per node, conn (E,4))
KIND_MASS_TET = 5    # ElasticityTetrahedralElement3D.elementary_mass_matrix


def _xy(coords, conn):
    p = np.asarray(coords, dtype=np.float64)[np.asarray(conn)]
    return p[:, :, 0], p[:, :, 1]


# ---------------------------------------------------------------- a-1
def b_matrix(coords, conn):
    """elements.py:395-416.  Returns (B (E,3,6), detJ (E,))."""
    x, y = _xy(coords, conn)
    beta = np.stack([y[:, 1] - y[:, 2], y[:, 2] - y[:, 0], y[:, 0] - y[:, 1]], axis=1)   # :403
    gamma = np.stack([x[:, 2] - x[:, 1], x[:, 0] - x[:, 2], x[:, 1] - x[:, 0]], axis=1)  # :404
    det = (x[:, 0] - x[:, 2]) * (y[:, 1] - y[:, 2]) - (y[:, 0] - y[:, 2]) * (x[:, 1] - x[:, 2])  # :406-408
    ne = len(det)
    data = np.zeros((ne, 3, 6))
    data[:, 0, 0::2] = beta          # :410
    data[:, 1, 1::2] = gamma         # :411
    data[:, 2, 0::2] = gamma         # :412
    data[:, 2, 1::2] = beta
    return (1.0 / det)[:, None, None] * data, det   # :414


# ---------------------------------------------------------------- a-2
def d_matrix(e_mod, nu, plane_strain, plane_stress):
    """elements.py:251-273 (flag checks), :418-434 (strain), :436-453 (stress)."""
    if plane_strain and plane_stress:
        raise ValueError('just one of plane_strain or plane_stress can be True')
    if not plane_strain and not plane_stress:
        raise ValueError('one of plane_strain or plane_stress must be True')
    e_mod = np.asarray(e_mod, dtype=np.float64)
    nu = np.asarray(nu, dtype=np.float64)
    if plane_strain:
        a = (e_mod * nu) / ((1 + nu) * (1 - 2 * nu))
    else:
        a = (e_mod * nu) / (1 - nu ** 2)
    b = e_mod / (2 * (1 + nu))
    c = a + 2 * b
    d = np.zeros(e_mod.shape + (3, 3))
    d[..., 0, 0] = c
    d[..., 0, 1] = a
    d[..., 1, 0] = a
    d[..., 1, 1] = c
    d[..., 2, 2] = b
    return d


def tri_area(coords, conn):
    """volmdlr TriangularElement2D.area = |u x v| / 2, u = p1-p0, v = p2-p0."""
    x, y = _xy(coords, conn)
    return 0.5 * np.abs((x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (y[:, 1] - y[:, 0]) * (x[:, 2] - x[:, 0]))


# ---------------------------------------------------------------- a-3
def ke_elasticity(coords, conn, mat_id, mat, plane_strain, plane_stress):
    """elements.py:466-511: thickness * area * (B^T D B), row-major 6x6."""
    bm, _ = b_matrix(coords, conn)
    m = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id)]
    d = d_matrix(m[:, 0], m[:, 1], plane_strain, plane_stress)
    area = tri_area(coords, conn)
    btd = np.matmul(bm.transpose(0, 2, 1), d)
    return (m[:, 2] * area)[:, None, None] * np.matmul(btd, bm)   # :508-509


# ---------------------------------------------------------------- a-4
def form_functions(coords, conn):
    """volmdlr TriangularElement2D.form_functions restated in closed form:
    N_i = a_i + b_i x + c_i y with N_i(p_j) = delta_ij.  Returns a, b, c (E,3)."""
    x, y = _xy(coords, conn)
    d = (x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0])
    i1, i2 = [1, 2, 0], [2, 0, 1]
    a = (x[:, i1] * y[:, i2] - x[:, i2] * y[:, i1]) / d[:, None]
    b = (y[:, i1] - y[:, i2]) / d[:, None]
    c = (x[:, i2] - x[:, i1]) / d[:, None]
    return a, b, c


def ke_magnetic(coords, conn, mat_id, mat):
    """elements.py:93-118: (1/mu) (b_i b_j + c_i c_j) * area, row-major 3x3."""
    _, b, c = form_functions(coords, conn)
    mu = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id), 0]
    area = tri_area(coords, conn)
    bb = b[:, :, None] * b[:, None, :] + c[:, :, None] * c[:, None, :]
    return (1.0 / mu)[:, None, None] * bb * area[:, None, None]


# ---------------------------------------------------------------- a-5
_MASS_PATTERN = np.kron(np.array([[2, 1, 1], [1, 2, 1], [1, 1, 2]], dtype=np.float64), np.eye(2))


def me_mass(coords, conn, mat_id, mat):
    """elements.py:513-536: (rho * area * t / 12) * ([[2,1,1],[1,2,1],[1,1,2]] (x) I2)."""
    m = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id)]
    area = tri_area(coords, conn)
    return ((m[:, 3] * area * m[:, 2]) / 12)[:, None, None] * _MASS_PATTERN[None]


# ---------------------------------------------------------------- §8f rank 4: P1 tetrahedra
def tet_geometry(coords, conn):
    """volmdlr TetrahedralElement restated (third-party, absent from /root/reference): with
    A = rows [1, x_j, y_j, z_j], volume = |det A| / 6 and the form functions as the reference uses
    them (elements.py:726-739, :749): tuples (alpha_i, a_i, b_i, c_i) with
    N_i = (alpha_i + a_i x + b_i y + c_i z) / (6 V), i.e. det(A) * column i of A^-1.
    Returns (volume (E,), form (E,4,4) with form[:, i] = (alpha_i, a_i, b_i, c_i))."""
    p = np.asarray(coords, dtype=np.float64)[np.asarray(conn)]          # (E,4,3)
    a = np.concatenate([np.ones(p.shape[:2] + (1,)), p], axis=2)       # (E,4,4)
    det = np.linalg.det(a)
    inv = np.linalg.inv(a)                                              # column i = coefficients of N_i
    form = (inv * np.abs(det)[:, None, None]).transpose(0, 2, 1)       # 6 V * coefficients
    return np.abs(det) / 6.0, form


def b_matrix_tet(coords, conn):
    """elements.py:719-751: B (E,6,12) = 1/(6 V) * [[a_i,0,0],[0,b_i,0],[0,0,c_i],[b_i,a_i,0],
    [0,c_i,b_i],[c_i,0,a_i]] per node i."""
    vol, form = tet_geometry(coords, conn)
    a, b, c = form[:, :, 1], form[:, :, 2], form[:, :, 3]              # (E,4)
    data = np.zeros((len(vol), 6, 12))
    data[:, 0, 0::3] = a
    data[:, 1, 1::3] = b
    data[:, 2, 2::3] = c
    data[:, 3, 0::3] = b
    data[:, 3, 1::3] = a
    data[:, 4, 1::3] = c
    data[:, 4, 2::3] = b
    data[:, 5, 0::3] = c
    data[:, 5, 2::3] = a
    return (1.0 / (6.0 * vol))[:, None, None] * data, vol               # :749


def d_matrix_tet(e_mod, nu):
    """elements.py:773-797 (the plane flags are ignored in 3D, :753-771)."""
    e_mod = np.asarray(e_mod, dtype=np.float64)
    nu = np.asarray(nu, dtype=np.float64)
    d = np.zeros(e_mod.shape + (6, 6))
    for i in range(3):
        for j in range(3):
            d[..., i, j] = np.where(i == j, 1 - nu, nu)
        d[..., 3 + i, 3 + i] = (1 - 2 * nu) / 2
    return (e_mod / ((1 + nu) * (1 - 2 * nu)))[..., None, None] * d     # :794-795


def ke_tet(coords, conn, mat_id, mat):
    """elements.py:809-828: volume * (B^T D B), row-major 12x12, DOF order [u0,v0,w0,u1,...]."""
    bm, vol = b_matrix_tet(coords, conn)
    m = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id)]
    d = d_matrix_tet(m[:, 0], m[:, 1])
    return vol[:, None, None] * np.matmul(np.matmul(bm.transpose(0, 2, 1), d), bm)


_MASS_PATTERN_TET = np.kron(np.ones((4, 4)) + np.eye(4), np.eye(3))


def me_tet(coords, conn, mat_id, mat):
    """elements.py:830-857: (rho * volume / 20) * ((1 + delta_ij) (x) I3)."""
    vol, _ = tet_geometry(coords, conn)
    m = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id)]
    return ((m[:, 3] * vol) / 20)[:, None, None] * _MASS_PATTERN_TET[None]


def post_tet(coords, conn, mat_id, mat, u):
    """results.py:809-830 / :769-781 for tetrahedra: strain = B u_e (E,6), stress = D B u_e (E,6),
    energy = 1/2 u_e^T Ke u_e (E,)."""
    bm, _ = b_matrix_tet(coords, conn)
    m = np.asarray(mat, dtype=np.float64)[np.asarray(mat_id)]
    d = d_matrix_tet(m[:, 0], m[:, 1])
    ue = np.asarray(u, dtype=np.float64)[element_dofs(conn, 3)]
    strain = np.einsum('eij,ej->ei', bm, ue)
    stress = np.einsum('eij,ej->ei', d, strain)
    ke = ke_tet(coords, conn, mat_id, mat)
    return strain, stress, 0.5 * np.einsum('ei,eij,ej->e', ue, ke, ue)


# ---------------------------------------------------------------- a-6
def element_to_node_factors(coords, conn):
    """elements.py:18-53 / :156-191: |det| * N_i(midpoint of points[1], points[2]).
    Numerically (~0, A, A) -- reproduced as written, not 'fixed'."""
    x, y = _xy(coords, conn)
    det = np.abs((x[:, 1] - x[:, 0]) * (y[:, 2] - y[:, 0]) - (x[:, 2] - x[:, 0]) * (y[:, 1] - y[:, 0]))  # :33
    a, b, c = form_functions(coords, conn)
    x2, y2, x3, y3 = x[:, 1, None], y[:, 1, None], x[:, 2, None], y[:, 2, None]
    return det[:, None] * (a + 0.5 * b * x2 + 0.5 * c * y2 + 0.5 * b * x3 + 0.5 * c * y3)   # :46-51


# ---------------------------------------------------------------- a-10 / a-11
def element_dofs(conn, dim):
    """analysis.py:714-735 with core.py:89-108: dofs[p*dim+d] = node_p*dim + d."""
    conn = np.asarray(conn, dtype=np.int64)
    return (conn[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(len(conn), conn.shape[1] * dim)


def triplet_indices(conn, dim):
    """analysis.py:722-733: row = repeat(dofs, 3*dim), col = tile(dofs, 3*dim)."""
    dofs = element_dofs(conn, dim)
    nd = dofs.shape[1]          # 3 dim for triangles, 4 dim for tetrahedra
    rows = np.repeat(dofs, nd, axis=1)
    cols = np.tile(dofs, (1, nd))
    return rows, cols


def element_matrices(kind, coords, conn, mat_id, mat):
    if kind == KIND_ELAST_PSTRESS:
        return ke_elasticity(coords, conn, mat_id, mat, False, True)
    if kind == KIND_ELAST_PSTRAIN:
        return ke_elasticity(coords, conn, mat_id, mat, True, False)
    if kind == KIND_MAGNETIC:
        return ke_magnetic(coords, conn, mat_id, mat)
    if kind == KIND_MASS:
        return me_mass(coords, conn, mat_id, mat)
    if kind == KIND_ELAST_TET:
        return ke_tet(coords, conn, mat_id, mat)
    if kind == KIND_MASS_TET:
        return me_tet(coords, conn, mat_id, mat)
    raise NotImplementedError(kind)


def kind_dim(kind):
    if kind in (KIND_ELAST_TET, KIND_MASS_TET):
        return 3
    return 1 if kind == KIND_MAGNETIC else 2


# ---------------------------------------------------------------- a-12 / a-13 / a-14
def assemble_k(kind, coords, conn, mat_id, mat, n_nodes=None):
    """analysis.py:324-339 + :661 restricted to the K block: COO of all
    (3 dim)^2 E triplets -> scipy canonical CSR (sorted, duplicates summed,
    explicit zeros kept), shape ndof x ndof."""
    dim = kind_dim(kind)
    n_nodes = len(coords) if n_nodes is None else n_nodes
    ke = element_matrices(kind, coords, conn, mat_id, mat)
    rows, cols = triplet_indices(conn, dim)
    ndof = n_nodes * dim
    k = sp.csr_matrix((ke.reshape(-1), (rows.reshape(-1), cols.reshape(-1))), shape=(ndof, ndof))
    k.sum_duplicates()
    k.sort_indices()
    return k


# ---------------------------------------------------------------- a-7
def dedup_last_wins(dofs, values):
    """analysis.py:28-47 / :71-90: `d[key] = + value` => assignment; the LAST
    record for a (node, dimension) key wins, and dict order = first-seen order.
    Returns (unique dofs in first-seen order, winning values)."""
    out = {}
    for k, v in zip(np.asarray(dofs).tolist(), np.asarray(values, dtype=np.float64).tolist()):
        out[k] = v
    return np.array(list(out.keys()), dtype=np.int64), np.array(list(out.values()), dtype=np.float64)


# ---------------------------------------------------------------- a-8
def loads_to_dof_records(coords, conn, dim, node_loads=(), elements_loads=(), edge_loads=()):
    """analysis.py:457-489 (order: node loads, element->node :407-429, edge->node
    :431-447); loads.py:24-35 (value_per_element = value * A_j / sum A).
    node_loads: (node, value, dimension); elements_loads: (elem_idx_list, value, dimension);
    edge_loads: (node_a, node_b, value, dimension)."""
    dofs, vals = [], []
    for n, v, d in node_loads:
        dofs.append(n * dim + (d - 1))
        vals.append(v)
    if len(elements_loads):
        area = tri_area(coords, conn)
        fac = element_to_node_factors(coords, conn)
    for idx, v, d in elements_loads:
        idx = np.asarray(idx, dtype=np.int64)
        total = sum(area[j] for j in idx)
        for j in idx:
            vpe = v * area[j] / total
            for p in range(3):
                dofs.append(int(conn[j][p]) * dim + (d - 1))
                vals.append(vpe * fac[j, p])
    for a, b, v, d in edge_loads:
        for n in (a, b):
            dofs.append(n * dim + (d - 1))
            vals.append(v * 0.5)
    return dedup_last_wins(dofs, vals)


def magnet_load_records(coords, conn, magnet_loads):
    """analysis.py:556-577 + loads.py:129-147 (MagnetLoad.contour_linear_elements).
    magnet_loads: (elem_idx_list, non_contour_node_list, m_x, m_y).  Contour = edges of the magnet's
    elements that occur exactly once and do not have BOTH ends in non_contour_nodes, in first-seen
    order (element order, then edges (p0,p1), (p1,p2), (p2,p0): volmdlr's TriangularElement2D
    .linear_elements -- third-party, restated).  Per contour edge both end nodes receive
    M . t * length / 2 with t = (-n_y, n_x), n the unit normal pointing to the triangle's third
    vertex (volmdlr `interior_normal`, restated); the rows are NODE indices (analysis.py:565-566),
    added to f without de-duplication (analysis.py:701-702).  Returns (rows, values)."""
    xy = np.asarray(coords, dtype=np.float64)
    rows, vals = [], []
    for idx, non_contour, mx, my in magnet_loads:
        count, first = {}, {}
        for j in idx:
            n = [int(v) for v in conn[j]]
            for i in range(3):
                a, b, c = n[i], n[(i + 1) % 3], n[(i + 2) % 3]
                key = frozenset((a, b))
                count[key] = count.get(key, 0) + 1
                if key not in first:
                    first[key] = (a, b, c)
        skip = set(int(v) for v in non_contour)
        for key, (a, b, c) in first.items():
            if count[key] != 1 or (a in skip and b in skip):
                continue
            t = xy[b] - xy[a]
            nrm = np.array([-t[1], t[0]])
            if nrm @ (xy[c] - xy[a]) < 0:
                nrm = -nrm
            nrm = nrm / np.hypot(nrm[0], nrm[1])
            share = (mx * (-nrm[1]) + my * nrm[0]) * np.hypot(t[0], t[1]) / 2
            rows += [a, b]
            vals += [share, share]
    return np.array(rows, dtype=np.int64), np.array(vals, dtype=np.float64)


# ---------------------------------------------------------------- a-9
def bcs_to_dof_records(coords, conn, dim, node_bcs=(), element_bcs=(), edge_bcs=()):
    """analysis.py:241-265 (order: node BCs, element->node :201-220, edge->node :222-239)."""
    dofs, vals = [], []
    for n, v, d in node_bcs:
        dofs.append(n * dim + (d - 1))
        vals.append(v)
    if len(element_bcs):
        fac = element_to_node_factors(coords, conn)
    for j, v, d in element_bcs:
        for p in range(3):
            dofs.append(int(conn[j][p]) * dim + (d - 1))
            vals.append(v * fac[j, p])
    for a, b, v, d in edge_bcs:
        for n in (a, b):
            dofs.append(n * dim + (d - 1))
            vals.append(v * 0.5)
    return dedup_last_wins(dofs, vals)


def augmented_matrix(k, bc_dofs):
    """analysis.py:617-663 + :272-277: K plus unit Lagrange rows/cols,
    canonical CSR of shape (ndof + n_bc)^2."""
    ndof = k.shape[0]
    nbc = len(bc_dofs)
    kc = k.tocoo()
    r = np.concatenate([kc.row, ndof + np.arange(nbc), bc_dofs])
    c = np.concatenate([kc.col, bc_dofs, ndof + np.arange(nbc)])
    v = np.concatenate([kc.data, np.ones(nbc), np.ones(nbc)])
    m = sp.csr_matrix((v, (r, c)), shape=(ndof + nbc, ndof + nbc))
    m.sum_duplicates()
    m.sort_indices()
    return m


# ---------------------------------------------------------------- a-15
def source_vector(ndof, load_dofs, load_vals, bc_vals):
    """analysis.py:665-708: zeros(ndof + n_bc, 1); f[row] += value."""
    f = np.zeros((ndof + len(bc_vals), 1))
    np.add.at(f[:, 0], np.asarray(load_dofs, dtype=np.int64), load_vals)
    f[ndof:, 0] += bc_vals
    return f


# ---------------------------------------------------------------- a-16
def solve_augmented(k_aug, f):
    """analysis.py:798-830: spsolve(K, f, permc_spec='NATURAL', use_umfpack=True)."""
    return spla.spsolve(k_aug, f, permc_spec='NATURAL', use_umfpack=True)


def eliminate_dirichlet(k, f, bc_dofs, bc_vals):
    """Symmetric in-place elimination on the K block (pattern unchanged):
    b -= K[:, c] g; zero row+col c; unit diagonal; b[c] = g.  Equivalent to the
    reference's Lagrange system on the displacement block."""
    k = k.tocsr().copy()
    b = np.array(f, dtype=np.float64).reshape(-1).copy()
    g = np.zeros(k.shape[0])
    g[bc_dofs] = bc_vals
    b -= k @ g
    is_bc = np.zeros(k.shape[0], dtype=bool)
    is_bc[bc_dofs] = True
    rows = np.repeat(np.arange(k.shape[0]), np.diff(k.indptr))
    kill = is_bc[rows] | is_bc[k.indices]
    diag = rows == k.indices
    k.data[kill] = 0.0
    k.data[diag & is_bc[rows]] = 1.0
    b[bc_dofs] = bc_vals
    return k, b


def solve_reduced_direct(k, f, bc_dofs, bc_vals, permc_spec='NATURAL'):
    ke, b = eliminate_dirichlet(k, f, bc_dofs, bc_vals)
    return spla.spsolve(ke.tocsc(), b, permc_spec=permc_spec)


def multipliers(k, u, f_loads, bc_dofs):
    """Rows bc_dofs of the augmented system: K u + lambda = f  =>  lambda = f - K u."""
    r = f_loads - k @ u
    return r[bc_dofs]


def jacobi_pcg(k, b, rtol=1e-8, maxit=100000, x0=None):
    """Plain Jacobi-preconditioned CG on an SPD CSR matrix (CPU baseline for the
    PCG DOF-iterations/s metric).  Returns x, iters, relres."""
    dinv = 1.0 / k.diagonal()
    x = np.zeros_like(b) if x0 is None else x0.copy()
    r = b - k @ x
    z = dinv * r
    p = z.copy()
    rz = float(r @ z)
    bnorm = float(np.sqrt(b @ b))
    if bnorm == 0.0:
        return x, 0, 0.0
    it = 0
    rel = float(np.sqrt(r @ r)) / bnorm
    while it < maxit and rel > rtol:
        q = k @ p
        alpha = rz / float(p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        rz_new = float(r @ z)
        rel = float(np.sqrt(r @ r)) / bnorm
        p = z + (rz_new / rz) * p
        rz = rz_new
        it += 1
    return x, it, rel


# ---------------------------------------------------------------- synthetic meshes
def modal_eigenvalues(k_mat, m_mat, k, order, free_dofs=None):
    """k 'largest' / 'smallest' eigenvalues (ascending) and M-orthonormal eigenvectors of
    K x = lambda M x by DENSE scipy.linalg.eigh -- the answer analysis.py:779-782
    (eigsh(A=K, M=M, which='LM', k=k)) converges to, and the mathematically intended result of the
    'smallest' branch (:788-794: 1 / eigs(K^-1 M), unusable as shipped because the unconstrained K
    is singular).  free_dofs restricts the pencil to those rows/columns (constrained modes).
    Small problems only: O(n^3)."""
    import scipy.linalg as sla
    kd = k_mat.toarray() if sp.issparse(k_mat) else np.asarray(k_mat)
    md = m_mat.toarray() if sp.issparse(m_mat) else np.asarray(m_mat)
    n = kd.shape[0]
    idx = np.arange(n) if free_dofs is None else np.asarray(free_dofs)
    w, v = sla.eigh(kd[np.ix_(idx, idx)], md[np.ix_(idx, idx)])
    if order == 'largest':
        sel = slice(len(w) - k, len(w))
    elif order == 'smallest':
        sel = slice(0, k)
    else:
        raise ValueError("Order parameter should be either 'largest' or 'smallest'")  # analysis.py:796
    vec = np.zeros((n, k))
    vec[idx] = v[:, sel]
    return w[sel], vec


def structured_mesh(nx, ny, h=None, jitter=0.0, seed=0):
    """Nodes row-major id = j*(nx+1)+i at (i h, j h), h = 1/ny; each cell ->
    T0 = [(i,j),(i+1,j),(i,j+1)], T1 = [(i+1,j+1),(i+1,j),(i,j+1)]
    (orientation of scripts/Elasticity/beam2d_example_2.py:39-40)."""
    h = 1.0 / ny if h is None else h
    ii, jj = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1))
    coords = np.stack([ii.reshape(-1) * h, jj.reshape(-1) * h], axis=1).astype(np.float64)
    if jitter:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter * h, jitter * h, size=coords.shape)
        interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny)).reshape(-1)
        coords[interior] += d[interior]
    ci, cj = np.meshgrid(np.arange(nx), np.arange(ny))
    ci, cj = ci.reshape(-1), cj.reshape(-1)
    n00 = cj * (nx + 1) + ci
    n10, n01, n11 = n00 + 1, n00 + nx + 1, n00 + nx + 2
    conn = np.empty((2 * nx * ny, 3), dtype=np.int32)
    conn[0::2] = np.stack([n00, n10, n01], axis=1)
    conn[1::2] = np.stack([n11, n10, n01], axis=1)
    return coords, conn


def structured_tet_mesh(nx, ny, nz, h=1.0, jitter=0.0, seed=0):
    """Box [0, nx h] x [0, ny h] x [0, nz h]: nodes id = (k (ny+1) + j) (nx+1) + i, every cell cut
    into 6 tetrahedra around the main diagonal (Kuhn triangulation; conforming across cells).
    Interior nodes optionally displaced by U(-jitter h, jitter h)^3."""
    ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (kk * (ny + 1) + jj) * (nx + 1) + ii
    coords = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    coords[nid.reshape(-1)] = np.stack([ii, jj, kk], axis=-1).reshape(-1, 3) * h
    if jitter:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter * h, jitter * h, size=coords.shape)
        interior = ((ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz))
        idx = nid[interior]
        coords[idx] += d[idx]
    c = nid[:-1, :-1, :-1].reshape(-1)
    dx, dy, dz = 1, nx + 1, (nx + 1) * (ny + 1)
    paths = [(dx, dy, dz), (dx, dz, dy), (dy, dx, dz), (dy, dz, dx), (dz, dx, dy), (dz, dy, dx)]
    tets = []
    for a, b, d3 in paths:
        tets.append(np.stack([c, c + a, c + a + b, c + a + b + d3], axis=1))
    conn = np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
    return coords, conn
