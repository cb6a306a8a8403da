"""Modal analysis  K x = lambda M x  (reference analysis.py:741-796).

CPU: the oracle's dense restatement against the reference's own eigsh output (golden
`eig_largest`), and the host-side LOBPCG logic (finite_elements_b200/modal.py, device-agnostic
torch code) driven by scipy products of the golden K / M.
GPU: `FiniteElementAnalysis.modal_analysis` through the C ABI (fe_assemble for K and M,
fe_spmm_pair / fe_csr_diagonal inside LOBPCG) against the same goldens and the oracle.

Tolerances: eigenvalues within 1e-8 relative (of the largest of the k values); eigenvector
residuals ||K x - lambda M x|| <= 1e-7 ||K x||."""
import numpy as np
import pytest
import torch

from oracle import numpy_oracle as no
from tests.fixtures import Fixture, names, names3d, build_object_analysis

MODAL = [n for n in names() + names3d() if "ref_eig_largest" in Fixture(n).z.files]
LOBPCG_SIZED = [n for n in MODAL if Fixture(n).ndof >= 300]


def test_goldens_exist():
    assert len(MODAL) >= 5 and len(LOBPCG_SIZED) >= 3


@pytest.mark.parametrize("name", MODAL)
def test_oracle_largest_vs_reference_eigsh(name):
    fx = Fixture(name)
    ref = fx.ref("eig_largest")
    lam, vec = no.modal_eigenvalues(fx.csr("k"), fx.csr("m"), len(ref), "largest")
    assert np.max(np.abs(lam - ref)) <= 1e-10 * ref.max()
    k, m = fx.csr("k"), fx.csr("m")
    assert np.allclose(vec.T @ (m @ vec), np.eye(len(ref)), atol=1e-10)
    assert np.max(np.abs(k @ vec - (m @ vec) * lam[None, :])) <= 1e-9 * np.max(np.abs(k @ vec))
    with pytest.raises(ValueError):
        no.modal_eigenvalues(k, m, 2, "middle")


def _scipy_pair(k, m):
    def apply_pair(v):
        a = v.numpy()
        return torch.from_numpy(k @ a), torch.from_numpy(m @ a)
    return apply_pair


@pytest.mark.parametrize("name", LOBPCG_SIZED)
@pytest.mark.parametrize("order", ["largest", "smallest"])
def test_lobpcg_host_logic_vs_dense(name, order):
    from finite_elements_b200.modal import lobpcg
    fx = Fixture(name)
    k, m = fx.csr("k").tocsr(), fx.csr("m").tocsr()
    n, nev = k.shape[0], 8
    largest = order == "largest"
    dinv = torch.from_numpy(1.0 / (m.diagonal() if largest else k.diagonal()))
    lam, vec, info = lobpcg(_scipy_pair(k, m), n, nev, "cpu", largest=largest,
                            precond=lambda r: dinv[:, None] * r, tol=1e-9, maxit=3000,
                            anorm=float(k.diagonal().max()))
    assert info.converged, info
    ref, _ = no.modal_eigenvalues(k, m, nev, order)
    lam = lam.numpy()
    assert np.all(np.diff(lam) >= 0)
    assert np.max(np.abs(lam - ref)) <= 1e-8 * np.abs(ref).max()
    if largest:
        assert np.max(np.abs(lam - fx.ref("eig_largest"))) <= 1e-8 * ref.max()
    v = vec.numpy()
    assert np.allclose(v.T @ (m @ v), np.eye(nev), atol=1e-8)


def test_lobpcg_constrained_and_chebyshev():
    from finite_elements_b200.modal import lobpcg, chebyshev_preconditioner, gershgorin_lmax
    fx = Fixture("struct24x16_jit_pstress")
    k, m = fx.csr("k").tocsr(), fx.csr("m").tocsr()
    n = k.shape[0]
    bc = np.unique(fx.ref("bc_dofs"))
    free = np.setdiff1d(np.arange(n), bc)
    mask = torch.ones(n, dtype=torch.float64)
    mask[torch.from_numpy(bc)] = 0.0
    dinv = torch.from_numpy(1.0 / k.diagonal())
    apply_k = lambda v: torch.from_numpy(k @ v.numpy())  # noqa: E731
    lmax = gershgorin_lmax(torch.from_numpy(np.asarray(abs(k).sum(axis=1)).ravel()), dinv)
    import scipy.sparse.linalg as spla
    import scipy.sparse as sp
    true_lmax = spla.eigsh(sp.diags(1.0 / np.sqrt(k.diagonal())) @ k @ sp.diags(1.0 / np.sqrt(k.diagonal())), k=1,
                           which='LA', return_eigenvectors=False)[0]
    assert true_lmax <= lmax <= 3.0 * true_lmax    # a guaranteed, reasonably tight bound
    prec = chebyshev_preconditioner(apply_k, dinv, lmax, 6, 30.0)
    lam, vec, info = lobpcg(_scipy_pair(k, m), n, 6, "cpu", precond=prec, mask=mask, tol=1e-9, maxit=2000,
                            anorm=float(k.diagonal().max()))
    assert info.converged
    ref, _ = no.modal_eigenvalues(k, m, 6, "smallest", free_dofs=free)
    assert ref[0] > 1e3    # clamped edge: no rigid-body modes
    assert np.max(np.abs(lam.numpy() - ref)) <= 1e-8 * ref.max()
    assert float(vec[torch.from_numpy(bc)].abs().max()) == 0.0
    # fewer outer iterations than plain Jacobi on the same problem
    lam_j, _, info_j = lobpcg(_scipy_pair(k, m), n, 6, "cpu", precond=lambda r: dinv[:, None] * r, mask=mask,
                              tol=1e-9, maxit=3000, anorm=float(k.diagonal().max()))
    assert info_j.converged and info.iterations < info_j.iterations
    with pytest.raises(ValueError):
        lobpcg(_scipy_pair(k, m), 40, 8, "cpu")


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", MODAL)
def test_gpu_modal_largest_vs_reference(name):
    fx = Fixture(name)
    an, mesh, _ = build_object_analysis(fx)
    ref = fx.ref("eig_largest")
    vals, vecs = an.modal_analysis("largest", len(ref))
    assert vals.shape == (len(ref),) and vecs.shape == (len(ref), fx.ndof)     # eigvecs.T like :784
    assert np.max(np.abs(vals - ref)) <= 1e-8 * ref.max()
    k, m = fx.csr("k"), fx.csr("m")
    kv, mv = k @ vecs.T, m @ vecs.T
    assert np.max(np.linalg.norm(kv - mv * vals[None, :], axis=0) / np.linalg.norm(kv, axis=0)) <= 1e-7
    assert np.allclose(vecs @ mv, np.eye(len(ref)), atol=1e-8)
    assert an.last_modal_info.get("dense", False) == (fx.ndof < 6 * (len(ref) + max(2, len(ref) // 4)))


@pytest.mark.gpu
@pytest.mark.parametrize("name", LOBPCG_SIZED)
def test_gpu_modal_smallest_and_constrained_vs_oracle(name):
    fx = Fixture(name)
    an, mesh, _ = build_object_analysis(fx)
    k, m = fx.csr("k"), fx.csr("m")
    vals, vecs = an.modal_analysis("smallest", 8)
    ref, _ = no.modal_eigenvalues(k, m, 8, "smallest")
    assert not an.last_modal_info.get("dense", False) and an.last_modal_info.converged
    assert np.max(np.abs(vals - ref)) <= 1e-8 * np.abs(ref).max()   # three rigid-body modes ~ 0 first
    assert np.sum(np.abs(vals) <= 1e-6 * np.abs(ref).max()) == 3
    bc = np.unique(fx.ref("bc_dofs"))
    free = np.setdiff1d(np.arange(fx.ndof), bc)
    vals_c, vecs_c = an.modal_analysis("smallest", 6, constrained=True, cheb_degree=6)
    ref_c, _ = no.modal_eigenvalues(k, m, 6, "smallest", free_dofs=free)
    assert np.max(np.abs(vals_c - ref_c)) <= 1e-8 * ref_c.max()
    assert np.abs(vecs_c[:, bc]).max() == 0.0
    with pytest.raises(ValueError):
        an.modal_analysis("middle", 3)


@pytest.mark.gpu
def test_gpu_spmm_pair_vs_scipy():
    from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS, KIND_MASS
    fx = Fixture("gmsh_beam_0.1")
    dm = DeviceMesh(fx.coords, fx.conn, fx.mat_id, dim=2)
    kv, mv = dm.assemble(KIND_ELAST_PSTRESS, fx.mat), dm.assemble(KIND_MASS, fx.mat)
    k, m = dm.to_scipy(kv), dm.to_scipy(mv)
    rng = np.random.default_rng(3)
    for width in (1, 3, 8, 12, 16, 33, 70):
        x = rng.standard_normal((dm.n_cols, width))
        xd = torch.as_tensor(x).cuda()
        ya, yb = dm.spmm_pair(kv, mv, xd)
        ra, rb = k @ x, m @ x
        assert np.max(np.abs(ya.cpu().numpy() - ra)) <= 1e-13 * np.abs(ra).max()
        assert np.max(np.abs(yb.cpu().numpy() - rb)) <= 1e-13 * np.abs(rb).max()
        ya2, none = dm.spmm_pair(kv, None, xd)
        assert none is None and torch.equal(ya2, ya)        # deterministic, same arithmetic
    assert np.array_equal(dm.csr_diagonal(kv).cpu().numpy(), k.diagonal())
    with pytest.raises(ValueError):
        dm.spmm_pair(kv, mv, torch.zeros(dm.n_cols, 4, device="cuda")[:, ::2])


@pytest.mark.gpu
@pytest.mark.parametrize("width", [1, 6, 10, 12])
def test_gpu_fused_chebyshev_step_vs_unfused(width):
    """fe_cheb_step == the three torch statements it replaces; the whole preconditioner agrees with
    the unfused one built on fe_spmm_pair."""
    from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS
    from finite_elements_b200.modal import chebyshev_preconditioner, gershgorin_lmax
    fx = Fixture("gmsh_beam_0.1")
    dm = DeviceMesh(fx.coords, fx.conn, fx.mat_id, dim=2)
    kv = dm.assemble(KIND_ELAST_PSTRESS, fx.mat)
    k = dm.to_scipy(kv)
    n = dm.n_rows
    dinv = 1.0 / dm.csr_diagonal(kv)
    rng = np.random.default_rng(5)
    d, r, z = (torch.as_tensor(rng.standard_normal((n, width))).cuda() for _ in range(3))
    d_out = torch.empty_like(d)
    r0, z0 = r.clone(), z.clone()
    dm.cheb_step(kv, dinv, d, d_out, r, z, 0.37, 1.9)
    r_ref = r0.cpu().numpy() - k @ d.cpu().numpy()
    assert np.max(np.abs(r.cpu().numpy() - r_ref)) <= 1e-13 * np.abs(r_ref).max()
    assert torch.equal(z, z0 + d)
    d_ref = 0.37 * d.cpu().numpy() + 1.9 * dinv.cpu().numpy()[:, None] * r_ref
    assert np.max(np.abs(d_out.cpu().numpy() - d_ref)) <= 1e-13 * np.abs(d_ref).max()
    with pytest.raises(ValueError):
        dm.cheb_step(kv, dinv, d, d, r, z, 0.1, 0.2)
    apply_k = lambda v: dm.spmm_pair(kv, None, v.contiguous())[0]  # noqa: E731
    ones = torch.ones(n, 1, dtype=torch.float64, device="cuda")
    lmax = gershgorin_lmax(dm.spmm_pair(kv.abs(), None, ones)[0][:, 0], dinv)
    fused = lambda a, b, c, e, c1, c2: dm.cheb_step(kv, dinv, a, b, c, e, c1, c2)  # noqa: E731
    res = torch.as_tensor(rng.standard_normal((n, width))).cuda()
    t_plain = chebyshev_preconditioner(apply_k, dinv, lmax, 7, 30.0)(res)
    t_fused = chebyshev_preconditioner(apply_k, dinv, lmax, 7, 30.0, fused_step=fused)(res)
    assert float((t_plain - t_fused).abs().max()) <= 1e-12 * float(t_plain.abs().max())


@pytest.mark.gpu
def test_gpu_modal_mid_size_properties():
    """256 x 128 cells (66 k DOF): no dense oracle; the pairs must satisfy the pencil (independent
    fe_spmv), be M-orthonormal, start with three rigid-body modes, and agree with scipy's
    shift-invert ARPACK on the oracle's matrices."""
    import scipy.sparse.linalg as spla
    import finite_elements_b200 as fe
    from finite_elements_b200.device import KIND_ELAST_PSTRESS, KIND_MASS
    coords, conn = no.structured_mesh(256, 128, jitter=0.2, seed=1)
    mat = np.array([[210e9, 0.25, 1.0, 7860.0]])
    mesh = fe.mesh.ArrayMesh(coords, conn, 'elasticity', mat, [0, len(conn)])
    an = fe.analysis.FiniteElementAnalysis(mesh, [], [], [], [], [], [], [], [], plane_strain=False,
                                           plane_stress=True)
    vals, vecs = an.modal_analysis("smallest", 10)
    info = an.last_modal_info
    assert info.converged and info.iterations < 1500
    dm = an._dm()
    kv, mv = dm.assemble(KIND_ELAST_PSTRESS, mat), dm.assemble(KIND_MASS, mat)
    for i in (0, 3, 9):
        x = torch.as_tensor(vecs[i]).cuda()
        kx, mx = dm.spmv(kv, x), dm.spmv(mv, x)
        res = float(torch.linalg.norm(kx - vals[i] * mx))
        assert res <= 1e-7 * float(torch.linalg.norm(kx)) + 1e-12 * float(kv.abs().max()) * float(torch.linalg.norm(x))
    mid = np.zeros(len(conn), dtype=np.int32)
    k = no.assemble_k(no.KIND_ELAST_PSTRESS, coords, conn, mid, mat)
    m = no.assemble_k(no.KIND_MASS, coords, conn, mid, mat)
    assert np.allclose(vecs @ (m @ vecs.T), np.eye(10), atol=1e-8)
    ref = np.sort(spla.eigsh(k, k=10, M=m, sigma=-1.0e4, which='LM', return_eigenvectors=False))
    assert np.sum(np.abs(vals) < 1e-3 * ref[3]) == 3
    assert np.max(np.abs(vals[3:] - ref[3:]) / ref[3:]) <= 1e-7
    top, _ = an.modal_analysis("largest", 4)
    ref_top = np.sort(spla.eigsh(k, k=4, M=m, which='LM', return_eigenvectors=False))
    assert np.max(np.abs(top - ref_top)) <= 1e-8 * ref_top.max()
