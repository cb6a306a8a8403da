"""GPU parity: libfe_b200.so (through ctypes) against fixtures minted from the reference's
own code (tests/golden) and against the numpy oracle on seeded inputs.

Bars (BASELINE.json north_star): CSR pattern bit-exact; element and global matrix entries
within 1e-12 (relative to the row maximum, so structurally-present zeros are covered);
solutions within 1e-8 relative to the reference's spsolve.
"""
import numpy as np
import pytest

from tests.fixtures import Fixture, names, assert_close_rowscaled, assert_csr_values_close

pytestmark = pytest.mark.gpu

ALL = names()


def _dm(fx, **kw):
    from finite_elements_b200.device import DeviceMesh
    return DeviceMesh(fx.coords, fx.conn, fx.mat_id, dim=fx.dim, **kw)


def _kind(fx):
    return fx.oracle_kind  # numeric values are shared with include/fe_b200.h


@pytest.mark.parametrize("name", ALL)
def test_element_matrices_vs_reference(name):
    from finite_elements_b200.device import KIND_MASS
    fx = Fixture(name)
    dm = _dm(fx)
    ke = dm.element_matrices(_kind(fx), fx.mat).cpu().numpy()
    assert_close_rowscaled(ke, fx.ref("ke"), 1e-12)
    if fx.kind == "elasticity":
        me = dm.element_matrices(KIND_MASS, fx.mat).cpu().numpy()
        assert_close_rowscaled(me, fx.ref("me"), 1e-12)


@pytest.mark.parametrize("name", ALL)
def test_source_factors_vs_reference(name):
    fx = Fixture(name)
    dm = _dm(fx)
    fac, area = dm.source_factors()
    fac, area = fac.cpu().numpy(), area.cpu().numpy()
    ref = fx.ref("factors")
    tol = 1e-12 * (2 * area) * max(1.0, np.abs(fx.coords).max()) ** 2 / np.minimum(1.0, 2 * area)
    assert np.all(np.abs(fac - ref) <= tol[:, None] + 1e-12 * np.abs(ref))
    sel = np.arange(len(fx.conn))[::3].astype(np.int32)
    fac2, _ = dm.source_factors(sel)
    assert np.array_equal(fac2.cpu().numpy(), fac[sel])


@pytest.mark.parametrize("name", ALL)
def test_pattern_bit_exact_and_values(name):
    from finite_elements_b200.device import KIND_MASS
    fx = Fixture(name)
    dm = _dm(fx)
    ref_k = fx.csr("k")
    rowptr, colidx = dm.csr_pattern()
    assert dm.nnz == ref_k.nnz
    assert np.array_equal(rowptr.cpu().numpy(), ref_k.indptr)
    assert np.array_equal(colidx.cpu().numpy(), ref_k.indices)
    v1 = dm.assemble(_kind(fx), fx.mat, variant=1)
    v2 = dm.assemble(_kind(fx), fx.mat, variant=2)
    v3 = dm.assemble(_kind(fx), fx.mat, variant=3)   # fan-ordered (all fixtures are manifold)
    v0 = dm.assemble(_kind(fx), fx.mat, variant=0)
    assert np.array_equal(v1.cpu().numpy(), v2.cpu().numpy()), "element-order variants must agree bit for bit"
    assert np.array_equal(v0.cpu().numpy(), v3.cpu().numpy()), "default = fan variant on a manifold mesh"
    v4 = dm.assemble(_kind(fx), fx.mat, variant=4)   # the same walk on the 8-byte records
    assert dm.fan_record_bytes in (4, 8)
    assert np.array_equal(v3.cpu().numpy(), v4.cpu().numpy()), "4-byte and 8-byte fan records must agree bit for bit"
    assert_csr_values_close(dm.to_scipy(v3), dm.to_scipy(v2), 1e-14)  # vertex relabelling: last-ulp only
    assert_csr_values_close(dm.to_scipy(v2), ref_k, 1e-12)
    assert_csr_values_close(dm.to_scipy(v3), ref_k, 1e-12)
    if fx.kind == "elasticity":
        m = dm.assemble(KIND_MASS, fx.mat)
        assert_csr_values_close(dm.to_scipy(m), fx.csr("m"), 1e-12)


@pytest.mark.parametrize("name", ["struct24x16_jit_pstress", "gmsh_beam_0.1", "semantics_mag"])
def test_assembly_is_ordered_sum_of_element_matrices(name):
    """Determinism contract: every slot = sum of its element contributions in ascending
    element order, bit for bit (the CPU replays the same additions on the GPU's Ke dump)."""
    fx = Fixture(name)
    dm = _dm(fx)
    ke = dm.element_matrices(_kind(fx), fx.mat).cpu().numpy()
    vals = dm.assemble(_kind(fx), fx.mat, variant=2).cpu().numpy()   # element-order kernel
    again = dm.assemble(_kind(fx), fx.mat, variant=2).cpu().numpy()
    assert np.array_equal(vals, again), "two runs must be bit-identical"
    for v in (0, 1, 3):
        a = dm.assemble(_kind(fx), fx.mat, variant=v).cpu().numpy()
        assert np.array_equal(a, dm.assemble(_kind(fx), fx.mat, variant=v).cpu().numpy()), f"variant {v} not reproducible"
    k = dm.to_scipy(dm.assemble(_kind(fx), fx.mat, variant=2))
    from oracle import numpy_oracle as no
    rows, cols = no.triplet_indices(fx.conn, fx.dim)
    # slot index of every triplet in the canonical CSR
    slot = np.empty(rows.size, dtype=np.int64)
    r, c = rows.reshape(-1), cols.reshape(-1)
    for t in range(rows.size):
        s, e = k.indptr[r[t]], k.indptr[r[t] + 1]
        slot[t] = s + np.searchsorted(k.indices[s:e], c[t])
    acc = np.zeros(k.nnz)
    first = np.ones(k.nnz, dtype=bool)
    flat = ke.reshape(-1)
    for t in range(rows.size):  # ascending element order, sequential fp64 adds
        if first[slot[t]]:
            acc[slot[t]] = flat[t]
            first[slot[t]] = False
        else:
            acc[slot[t]] += flat[t]
    assert np.array_equal(acc, vals)


@pytest.mark.parametrize("name", ALL)
def test_solution_vs_reference_spsolve(name):
    from finite_elements_b200.device import solve_dirichlet_system
    from oracle import numpy_oracle as no
    fx = Fixture(name)
    dm = _dm(fx)
    ld, lv = no.loads_to_dof_records(fx.coords, fx.conn, fx.dim, fx.rec("node_loads"), fx.rec("elements_loads"),
                                     fx.rec("edge_loads"))
    if fx.records.get("magnet_loads"):   # analysis.py:556-577: added on top of the merged loads (f[row] += value)
        mr, mv = no.magnet_load_records(fx.coords, fx.conn, fx.rec("magnet_loads"))
        f = np.zeros(fx.ndof)
        np.add.at(f, ld, lv)
        np.add.at(f, mr, mv)
        ld = np.nonzero(f)[0]
        lv = f[ld]
    bc_dofs, bc_vals = fx.ref("bc_dofs"), fx.ref("bc_vals")
    u, lam, iters, relres = solve_dirichlet_system(dm, _kind(fx), fx.mat, ld, lv, bc_dofs, bc_vals, rtol=1e-13)
    ref_x = fx.ref("x")
    un = np.linalg.norm(ref_x[:fx.ndof])
    err = np.linalg.norm(u.cpu().numpy() - ref_x[:fx.ndof]) / un
    assert err <= 1e-8, f"solution error {err:.2e} after {iters} iterations (relres {relres:.1e})"
    ln = np.linalg.norm(ref_x[fx.ndof:])
    assert np.linalg.norm(lam.cpu().numpy() - ref_x[fx.ndof:]) <= 1e-6 * ln


@pytest.mark.parametrize("name", ["gmsh_beam_0.3", "struct24x16_jit_mag"])
def test_spmv_and_dirichlet_vs_scipy(name):
    import torch
    from oracle import numpy_oracle as no
    fx = Fixture(name)
    dm = _dm(fx)
    vals = dm.assemble(_kind(fx), fx.mat)
    k = dm.to_scipy(vals)
    rng = np.random.default_rng(1)
    x = rng.standard_normal(fx.ndof)
    y = dm.spmv(vals, torch.as_tensor(x).cuda()).cpu().numpy()
    yr = k @ x
    assert np.max(np.abs(y - yr)) <= 1e-13 * np.max(np.abs(yr))
    f = rng.standard_normal(fx.ndof)
    bc_dofs, bc_vals = fx.ref("bc_dofs"), rng.standard_normal(len(fx.ref("bc_dofs")))
    rhs = torch.as_tensor(f.copy()).cuda()
    dm.dirichlet(vals, rhs, bc_dofs, bc_vals)
    ke, b = no.eliminate_dirichlet(k, f, bc_dofs, bc_vals)
    assert_csr_values_close(dm.to_scipy(vals), ke, 1e-15)
    assert np.max(np.abs(rhs.cpu().numpy() - b)) <= 1e-12 * np.max(np.abs(b))


# ------------------------------------------------------------------ seeded mid-size vs oracle
@pytest.mark.parametrize("kind_name,jitter", [("stress", 0.0), ("strain", 0.2), ("mag", 0.2), ("mass", 0.2)])
def test_mid_size_vs_oracle(kind_name, jitter):
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh
    coords, conn = no.structured_mesh(96, 64, jitter=jitter, seed=3)
    rng = np.random.default_rng(5)
    mat_id = rng.integers(0, 3, size=len(conn)).astype(np.int32)
    if kind_name == "mag":
        kind, mat = no.KIND_MAGNETIC, np.array([[4e-7 * np.pi * s, 0, 0, 0] for s in (1e5, 1, 5e4)])
    else:
        kind = {"stress": no.KIND_ELAST_PSTRESS, "strain": no.KIND_ELAST_PSTRAIN, "mass": no.KIND_MASS}[kind_name]
        mat = np.array([[210e9, 0.25, 1.0, 7860], [70e9, 0.33, 0.5, 2700], [1e9, 0.45, 2.0, 1200]])
    dm = DeviceMesh(coords, conn, mat_id, dim=no.kind_dim(kind))
    k_ref = no.assemble_k(kind, coords, conn, mat_id, mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(kind, mat)), k_ref, 1e-12)
    ke = dm.element_matrices(kind, mat).cpu().numpy()
    assert_close_rowscaled(ke, no.element_matrices(kind, coords, conn, mat_id, mat).reshape(len(conn), -1), 1e-12)


def test_unstructured_shuffled_numbering_vs_oracle():
    """Random node permutation + random element order + flipped orientations: nothing in
    the plan may depend on the structured numbering."""
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh
    coords, conn = no.structured_mesh(40, 30, jitter=0.25, seed=11)
    rng = np.random.default_rng(12)
    perm = rng.permutation(len(coords))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    coords2 = coords[perm]
    conn2 = inv[conn][rng.permutation(len(conn))].astype(np.int32)
    flip = rng.random(len(conn2)) < 0.5
    conn2[flip] = conn2[flip][:, [0, 2, 1]]
    mat = np.array([[210e9, 0.25, 1.0, 7860]])
    dm = DeviceMesh(coords2, conn2, None, dim=2)
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords2, conn2, np.zeros(len(conn2), np.int32), mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat)), k_ref, 1e-12)


# ------------------------------------------------------------------ edge cases
def test_empty_single_isolated_and_fan():
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh
    mat = np.array([[1.0, 0.3, 1.0, 1.0]])
    dm = DeviceMesh(np.zeros((3, 2)), np.zeros((0, 3), np.int32), None, dim=2)
    assert dm.nnz == 0 and np.array_equal(dm.csr_pattern()[0].cpu().numpy(), np.zeros(7, np.int32))
    # one element + an isolated node (empty rows, as scipy gives)
    coords = np.array([[0, 0], [1, 0], [0, 1], [5, 5]], float)
    conn = np.array([[0, 1, 2]], np.int32)
    dm = DeviceMesh(coords, conn, None, dim=2)
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, np.zeros(1, np.int32), mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat)), k_ref, 1e-12)
    # fan: hub of valence 60 -> beyond the shared-memory tile variant, default must still work
    nf = 60
    ang = np.linspace(0, 2 * np.pi, nf, endpoint=False)
    coords = np.vstack([[0.0, 0.0], np.stack([np.cos(ang), np.sin(ang)], axis=1)])
    conn = np.array([[0, 1 + i, 1 + (i + 1) % nf] for i in range(nf)], np.int32)
    dm = DeviceMesh(coords, conn, None, dim=2)
    assert dm.max_degree == nf + 1
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, np.zeros(nf, np.int32), mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat)), k_ref, 1e-12)
    with pytest.raises(NotImplementedError):
        dm.assemble(no.KIND_ELAST_PSTRESS, mat, variant=2)


def test_non_manifold_and_bowtie_fall_back_from_fan_variant():
    """Edge shared by three elements: not a simple fan -> variant 3 refuses, default still exact.
    Bow-tie (two fans meeting at one node) and open boundary fans stay on the fan path."""
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh
    mat = np.array([[3.0, 0.3, 1.0, 1.0]])
    coords = np.array([[0, 0], [1, 0], [0.5, 1], [0.5, -1], [0.4, 0.6]], float)
    conn = np.array([[0, 1, 2], [1, 0, 3], [0, 1, 4]], np.int32)       # edge (0,1) in three triangles
    dm = DeviceMesh(coords, conn, None, dim=2)
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, np.zeros(3, np.int32), mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat)), k_ref, 1e-12)
    with pytest.raises(NotImplementedError):
        dm.assemble(no.KIND_ELAST_PSTRESS, mat, variant=3)
    coords = np.array([[0, 0], [1, 0], [1, 1], [-1, 0], [-1, -1], [0.2, 1.0]], float)
    conn = np.array([[0, 1, 2], [0, 3, 4], [2, 5, 0]], np.int32)       # node 0: a 2-corner fan + a 1-corner fan
    dm = DeviceMesh(coords, conn, None, dim=2)
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, np.zeros(3, np.int32), mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat, variant=3)), k_ref, 1e-12)
    dm1 = DeviceMesh(coords, conn, None, dim=1)
    k_ref = no.assemble_k(no.KIND_MAGNETIC, coords, conn, np.zeros(3, np.int32), np.array([[2.0, 0, 0, 0]]))
    assert_csr_values_close(dm1.to_scipy(dm1.assemble(no.KIND_MAGNETIC, np.array([[2.0, 0, 0, 0]]), variant=3)),
                            k_ref, 1e-12)


def test_compact_fan_records_and_their_fallbacks():
    """4-byte fan records (plan.cu: k_fan_compact) need a banded numbering and at most two materials around a
    node; outside that the plan keeps the 8-byte records.  Same values either way."""
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh
    mat = np.array([[1.0, 0.3, 1.0, 1.0], [2.5, 0.25, 1.0, 0.5], [0.7, 0.1, 1.0, 2.0]])
    coords, conn = no.structured_mesh(40, 24, jitter=0.2, seed=3)
    two = (np.arange(len(conn)) >= len(conn) // 2).astype(np.int32)      # two bands: <= 2 materials per star
    three = (np.arange(len(conn)) % 3).astype(np.int32)                   # three materials around most nodes
    for kind, dim in ((no.KIND_ELAST_PSTRESS, 2), (no.KIND_MAGNETIC, 1)):
        m = mat if dim == 2 else np.array([[1.0, 0, 0, 0], [50.0, 0, 0, 0], [3.0, 0, 0, 0]])
        for mat_id, rb in ((two, 4), (three, 8)):
            dm = DeviceMesh(coords, conn, mat_id, dim=dim)
            assert dm.fan_record_bytes == rb
            v3 = dm.assemble(kind, m, variant=3)
            assert np.array_equal(v3.cpu().numpy(), dm.assemble(kind, m, variant=4).cpu().numpy())
            assert_csr_values_close(dm.to_scipy(v3), no.assemble_k(kind, coords, conn, mat_id, m), 1e-12)
    # a numbering that is not banded: neighbours further than 2^18 apart
    coords, conn = no.structured_mesh(720, 500, jitter=0.1, seed=5)   # 361 221 nodes: differences beyond 2^18
    perm = np.random.default_rng(7).permutation(len(coords))
    inv = np.empty_like(perm)
    inv[perm] = np.arange(len(perm))
    coords_p, conn_p = coords[perm], inv[conn].astype(np.int32)
    mid = np.zeros(len(conn), np.int32)
    dm = DeviceMesh(coords_p, conn_p, mid, dim=2)
    assert dm.fan_record_bytes == 8
    assert DeviceMesh(coords, conn, mid, dim=2).fan_record_bytes == 4
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords_p, conn_p, mid, mat)
    assert_csr_values_close(dm.to_scipy(dm.assemble(no.KIND_ELAST_PSTRESS, mat)), k_ref, 1e-12)


def test_error_mapping():
    from finite_elements_b200.device import DeviceMesh, KIND_MAGNETIC, KIND_ELAST_PSTRESS
    import torch
    coords = np.array([[0, 0], [1, 0], [0, 1]], float)
    with pytest.raises(ValueError):
        DeviceMesh(coords, np.array([[0, 1, 7]], np.int32), None, dim=2)  # node out of range
    dm = DeviceMesh(coords, np.array([[0, 1, 2]], np.int32), None, dim=2)
    with pytest.raises(ValueError):
        dm.assemble(KIND_MAGNETIC, np.array([[1.0, 0, 0, 0]]))  # dim mismatch
    # no Dirichlet condition + negative modulus -> not SPD -> the reference's NotImplementedError
    vals = dm.assemble(KIND_ELAST_PSTRESS, np.array([[-1.0, 0.3, 1.0, 1.0]]))
    with pytest.raises(NotImplementedError):
        dm.pcg(vals, torch.ones(6, dtype=torch.float64, device="cuda"))


# ------------------------------------------------------------------ size-independent properties at scale
def _device_structured(nx, ny):
    import torch
    from finite_elements_b200.mesh import structured_mesh_torch
    return structured_mesh_torch(nx, ny, device=torch.device("cuda", 0))


@pytest.mark.parametrize("nx,ny", [(1024, 512), (4096, 2048)])
def test_properties_at_baseline_sizes(nx, ny):
    """S1M / S16M: nnz formula, sorted rows, symmetry, rigid-body null space,
    mass sum, run-to-run determinism, PCG residual."""
    import torch
    from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS, KIND_MASS
    coords, conn = _device_structured(nx, ny)
    dm = DeviceMesh(coords, conn, None, dim=2)
    n_nodes = (nx + 1) * (ny + 1)
    n_el = 2 * nx * ny
    edges = (3 * n_el + 2 * (nx + ny)) // 2
    assert dm.nnz == 4 * (n_nodes + 2 * edges)
    rowptr, colidx = dm.csr_pattern()
    lens = (rowptr[1:] - rowptr[:-1]).long()
    rows = torch.repeat_interleave(torch.arange(dm.n_rows, device="cuda"), lens)
    same_row = rows[1:] == rows[:-1]
    assert bool(torch.all(colidx[1:][same_row] > colidx[:-1][same_row])), "columns must be strictly increasing"
    del same_row
    mat = np.array([[210e9, 0.25, 1.0, 7860.0]])
    vals = dm.assemble(KIND_ELAST_PSTRESS, mat)
    assert torch.equal(vals, dm.assemble(KIND_ELAST_PSTRESS, mat))
    scale = float(vals.abs().max())
    # rigid modes: K [1,0,1,0..] = K [0,1,0,1..] = K (rotation) = 0
    xy = coords.reshape(-1)
    for mode in (torch.tensor([1.0, 0.0]), torch.tensor([0.0, 1.0])):
        x = mode.double().cuda().repeat(n_nodes)
        assert float(dm.spmv(vals, x).abs().max()) <= 1e-12 * scale
    rot = torch.stack([-coords[:, 1], coords[:, 0]], dim=1).reshape(-1).contiguous()
    assert float(dm.spmv(vals, rot).abs().max()) <= 1e-11 * scale * float(xy.abs().max())
    # symmetry through x^T K y == y^T K x
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(dm.n_rows, dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand(dm.n_rows, dtype=torch.float64, device="cuda", generator=g)
    a, b = float(torch.dot(x, dm.spmv(vals, y))), float(torch.dot(y, dm.spmv(vals, x)))
    assert abs(a - b) <= 1e-12 * abs(a)
    # total mass per direction = rho * area * t
    m = dm.assemble(KIND_MASS, mat)
    ones_x = torch.tensor([1.0, 0.0]).double().cuda().repeat(n_nodes)
    total = float(torch.dot(ones_x, dm.spmv(m, ones_x)))
    assert abs(total - 7860.0 * (nx / ny) * 1.0) <= 1e-10 * total
    del m
    # clamp the left edge, load the right edge, solve to 1e-8, check the TRUE residual
    h = 1.0 / ny
    left = torch.arange(ny + 1, device="cuda") * (nx + 1)
    bc = torch.stack([2 * left, 2 * left + 1], dim=1).reshape(-1).int()
    f = torch.zeros(dm.n_rows, dtype=torch.float64, device="cuda")
    f[2 * (left + nx) + 1] = -1000.0 * h
    rhs = f.clone()
    dm.dirichlet(vals, rhs, bc, torch.zeros(bc.numel(), dtype=torch.float64, device="cuda"))
    u, iters, relres = dm.pcg(vals, rhs, rtol=1e-8)
    true = float(torch.linalg.norm(rhs - dm.spmv(vals, u)) / torch.linalg.norm(rhs))
    assert relres <= 1e-8 and true <= 1.0001e-8, (iters, relres, true)   # relres is the TRUE residual
    assert float(u[bc.long()].abs().max()) == 0.0


@pytest.mark.parametrize("jitter", [0.0, 0.2])
def test_s1m_values_and_solution_vs_oracle_direct_solve(jitter):
    """BASELINE configs[2] (S1M, 1024 x 512 cells, plane stress) against the oracle -- SURVEY §8c
    "Oracle at scale": the assembled K (pattern bit-exact, values 1e-12 of the row maximum) and the
    displacements against a DIRECT solve of the Dirichlet-reduced SPD system (the reference's contract
    is spsolve, analysis.py:820-822; the reduced solve equals its Lagrange solve on the solution block,
    tests/test_oracle.py), tolerance 1e-8 relative as north_star states.  PCG runs at rtol 1e-12
    because a 1e-8 residual does not bound the solution error at this condition number.
    jitter 0.2: general geometry (no exact zeros), same sizes."""
    import scipy.sparse.linalg as spla
    import torch
    from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS
    from oracle import numpy_oracle as no
    nx, ny = 1024, 512
    coords, conn = no.structured_mesh(nx, ny, jitter=jitter, seed=0)
    mat = np.array([[210e9, 0.25, 1.0, 7860.0]])
    mat_id = np.zeros(len(conn), np.int32)
    dm = DeviceMesh(coords, conn, None, dim=2)
    vals = dm.assemble(KIND_ELAST_PSTRESS, mat)
    k_ref = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, mat_id, mat)
    assert_csr_values_close(dm.to_scipy(vals), k_ref, 1e-12)
    if jitter:
        return          # one direct solve is enough (23 s of SuperLU); the jittered case pins the values
    left = np.arange(ny + 1) * (nx + 1)
    bc = np.stack([2 * left, 2 * left + 1], axis=1).reshape(-1)
    f = np.zeros(dm.n_rows)
    f[2 * (left + nx) + 1] = -1000.0 / ny
    ke, b = no.eliminate_dirichlet(k_ref, f, bc, np.zeros(len(bc)))
    lu = spla.splu(ke.tocsc(), permc_spec='MMD_AT_PLUS_A')      # SPD: a fill-reducing symmetric ordering is safe
    x_ref = lu.solve(b)
    x_ref += lu.solve(b - ke @ x_ref)                            # one step of iterative refinement
    rhs = torch.as_tensor(f).cuda()
    dm.dirichlet(vals, rhs, bc, np.zeros(len(bc)))
    assert np.allclose(rhs.cpu().numpy(), b, rtol=0, atol=1e-12 * np.abs(b).max())
    u, iters, relres = dm.pcg(vals, rhs, rtol=1e-12, maxit=200000, raise_on_maxit=False)
    err = np.linalg.norm(u.cpu().numpy() - x_ref) / np.linalg.norm(x_ref)
    # (the true residual stalls near 3e-9 at this condition number; the SOLUTION error is what north_star bounds)
    assert relres <= 1e-8 and err <= 1e-8, (iters, relres, err)


@pytest.mark.parametrize("name", ["struct24x16_jit_pstress", "struct24x16_jit_mag", "gmsh_beam_0.1"])
def test_restricted_dirichlet_equals_full_sweep(name, monkeypatch):
    """fe_dirichlet_apply visits only the rows next to a condition (k_bc_rows); the sweep over every row it
    replaces (FE_B200_BC_SWEEP=1) must give the same matrix and right-hand side bit for bit -- random
    condition sets with adjacent condition nodes, non-zero values, up to a third of all DOFs."""
    import torch
    fx = Fixture(name)
    dm = _dm(fx)
    k0 = dm.assemble(_kind(fx), fx.mat)
    rng = np.random.default_rng(5)
    for frac in (0.002, 0.05, 0.33):
        n_bc = max(1, int(frac * fx.ndof))
        dofs = rng.choice(fx.ndof, size=n_bc, replace=False).astype(np.int32)
        g = rng.standard_normal(n_bc)
        f = torch.as_tensor(rng.standard_normal(fx.ndof)).cuda()
        out = []
        for sweep in (False, True):
            if sweep:
                monkeypatch.setenv("FE_B200_BC_SWEEP", "1")
            else:
                monkeypatch.delenv("FE_B200_BC_SWEEP", raising=False)
            vals, rhs = k0.clone(), f.clone()
            dm.dirichlet(vals, rhs, dofs, g)
            out.append((vals, rhs))
        assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1]), frac
        assert torch.equal(out[0][1][torch.as_tensor(dofs).long().cuda()], torch.as_tensor(g).cuda())


@pytest.mark.parametrize("grid_cap", [0, 37, 3])
def test_persistent_pcg_kernel_matches_three_kernel_path(grid_cap, monkeypatch):
    """pcg_persist.cuh (single-reduction CG in one cooperative kernel; the default from 4 ranks up) forced on
    one GPU against the three-kernel path on a mesh with several SpMV tiles per CTA (grid_cap limits the
    grid so that CTAs run different numbers of tiles and chunks -- the configuration that exposes a missing
    grid barrier): same iteration count within 1 %, same solution, same fixed-iteration iterates."""
    import torch
    from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS
    from finite_elements_b200.mesh import structured_mesh_torch
    nx, ny = 512, 256
    coords, conn = structured_mesh_torch(nx, ny, torch.device("cuda", 0))
    dm = DeviceMesh(coords, conn, None, dim=2)
    left = torch.arange(ny + 1, device="cuda") * (nx + 1)
    bc = torch.stack([2 * left, 2 * left + 1], dim=1).reshape(-1).int()
    f = torch.zeros(dm.n_rows, dtype=torch.float64, device="cuda")
    f[2 * (left + nx) + 1] = -1000.0 / ny
    vals = dm.assemble(KIND_ELAST_PSTRESS, np.array([[210e9, 0.25, 1.0, 7860.0]]))
    rhs = f.clone()
    dm.dirichlet(vals, rhs, bc, torch.zeros(bc.numel(), dtype=torch.float64, device="cuda"))
    monkeypatch.setenv("FE_B200_PERSIST", "0")
    u0, it0, rel0 = dm.pcg(vals, rhs, rtol=1e-9)
    x0 = dm.pcg_fixed(vals, rhs, torch.zeros_like(rhs), 25).clone()
    monkeypatch.setenv("FE_B200_PERSIST", "1")
    if grid_cap:
        monkeypatch.setenv("FE_B200_PERSIST_GRID", str(grid_cap))
    launches = dm.ctx.launches
    u1, it1, rel1 = dm.pcg(vals, rhs, rtol=1e-9)
    assert dm.ctx.launches - launches <= 12, "the persistent path must run (a handful of launches per solve)"
    x1 = dm.pcg_fixed(vals, rhs, torch.zeros_like(rhs), 25)
    assert rel0 <= 1e-9 and rel1 <= 1e-9 and abs(it1 - it0) <= max(3, it0 // 100), (it0, it1, rel0, rel1)
    assert float(torch.linalg.norm(u1 - u0) / torch.linalg.norm(u0)) <= 1e-7     # two solves at relres 1e-9, kappa ~1e6
    assert float(torch.linalg.norm(x1 - x0) / torch.linalg.norm(x0)) <= 1e-12    # same 25 iterates in exact arithmetic


@pytest.mark.gpu
def test_magnetic_properties_at_s1m():
    """BASELINE configs[1] scaled up (SURVEY §8d magnetic config, 1024 x 512 cells, three mu bands):
    nnz formula, K 1 = 0 (a constant potential carries no field), symmetry, determinism, and the
    Dirichlet-reduced system solved by PCG to a true residual of 1e-8."""
    import torch
    from finite_elements_b200.device import DeviceMesh, KIND_MAGNETIC
    nx, ny = 1024, 512
    coords, conn = _device_structured(nx, ny)
    n_el, n_nodes = 2 * nx * ny, (nx + 1) * (ny + 1)
    mu0 = 4e-7 * np.pi
    mat = np.array([[mu0 * 1e5, 0, 0, 0], [mu0, 0, 0, 0], [mu0 * 5e4, 0, 0, 0]])
    mat_id = (((torch.arange(n_el, device="cuda") // 2) % nx) * 3 // nx).to(torch.int32)
    dm = DeviceMesh(coords, conn, mat_id, dim=1)
    edges = (3 * n_el + 2 * (nx + ny)) // 2
    assert dm.nnz == n_nodes + 2 * edges and dm.n_rows == n_nodes
    vals = dm.assemble(KIND_MAGNETIC, mat)
    assert torch.equal(vals, dm.assemble(KIND_MAGNETIC, mat))
    scale = float(vals.abs().max())
    ones = torch.ones(n_nodes, dtype=torch.float64, device="cuda")
    assert float(dm.spmv(vals, ones).abs().max()) <= 1e-12 * scale
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(n_nodes, dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand(n_nodes, dtype=torch.float64, device="cuda", generator=g)
    a, b = float(torch.dot(x, dm.spmv(vals, y))), float(torch.dot(y, dm.spmv(vals, x)))
    assert abs(a - b) <= 1e-12 * abs(a)
    # the default (fan) variant and the row-owner gather in global memory agree to rounding
    v1 = dm.assemble(KIND_MAGNETIC, mat, variant=1)
    assert float((vals - v1).abs().max()) <= 1e-13 * scale
    right = (torch.arange(ny + 1, device="cuda") * (nx + 1) + nx).int()
    f = torch.zeros(n_nodes, dtype=torch.float64, device="cuda")
    f[conn[:2].reshape(-1).long()] = 2.5e9
    rhs = f.clone()
    dm.dirichlet(vals, rhs, right, torch.zeros(right.numel(), dtype=torch.float64, device="cuda"))
    # mu contrast 1e5 x mesh 1e6: ||A|| ||x|| / ||b|| puts the attainable FP64 residual near 1e-8
    # (fe_pcg stops there and says so); 1e-6 is comfortably reachable and must be met exactly
    u, iters, relres = dm.pcg(vals, rhs, rtol=1e-6, maxit=400000)
    true = float(torch.linalg.norm(rhs - dm.spmv(vals, u)) / torch.linalg.norm(rhs))
    assert relres <= 1e-6 and abs(true - relres) <= 1e-3 * relres, (iters, relres, true)
    assert float(u[right.long()].abs().max()) == 0.0


@pytest.mark.parametrize("dofs", [1, 2, 3])
def test_streamed_spmv_kernels_vs_scipy_cg(dofs):
    """The PCG's SpMV is a different kernel per DOF count (k_spmv_stream1 / k_spmv_stream / k_spmv_stream3: TMA-streamed
    tiles, one column index per entry / 2x2 block / 3x3 block).  Three iterations of fe_pcg_fixed from x = 0 must
    reproduce the same three Jacobi-PCG iterations done with scipy on the matrix copied to the host."""
    import torch
    from oracle import numpy_oracle as no
    from finite_elements_b200.device import DeviceMesh, DeviceMesh3D, KIND_MAGNETIC, KIND_ELAST_PSTRESS, KIND_ELAST_TET
    rng = np.random.default_rng(11)
    if dofs == 3:
        coords, conn = no.structured_tet_mesh(13, 9, 7, h=0.3, jitter=0.15, seed=2)
        dm = DeviceMesh3D(coords, conn, None)
        vals = dm.assemble(KIND_ELAST_TET, np.array([[2.0e3, 0.3, 1.0, 1.0]]))
        fixed = np.nonzero(coords[:, 0] == 0)[0]
    else:
        coords, conn = no.structured_mesh(97, 61, jitter=0.2, seed=4)
        dm = DeviceMesh(coords, conn, None, dim=dofs)
        mat = np.array([[1.0e3, 0.3, 1.0, 1.0]]) if dofs == 2 else np.array([[2.0, 0, 0, 0]])
        vals = dm.assemble(KIND_ELAST_PSTRESS if dofs == 2 else KIND_MAGNETIC, mat)
        fixed = np.nonzero(coords[:, 0] == 0)[0]
    n = dm.n_rows
    bc = (dofs * fixed[:, None] + np.arange(dofs)[None, :]).reshape(-1)
    b = torch.as_tensor(rng.standard_normal(n)).cuda()
    dm.dirichlet(vals, b, bc, np.zeros(len(bc)))
    a = dm.to_scipy(vals)
    bh = b.cpu().numpy()
    dinv = 1.0 / a.diagonal()
    x, r = np.zeros(n), bh.copy()
    z = dinv * r
    p, rz = z.copy(), r @ z
    for _ in range(3):                                   # textbook PCG, the recurrence solve.cu implements
        q = a @ p
        alpha = rz / (p @ q)
        x += alpha * p
        r -= alpha * q
        z = dinv * r
        rz_new = r @ z
        p = z + (rz_new / rz) * p
        rz = rz_new
    xd = torch.zeros(n, dtype=torch.float64, device="cuda")
    dm.pcg_fixed(vals, b, xd, 3)
    err = np.abs(xd.cpu().numpy() - x).max() / np.abs(x).max()
    assert err <= 1e-12, err
