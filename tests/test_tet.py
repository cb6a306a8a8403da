"""Linear tetrahedra, 3 DOF per node (reference elements.py:663-876).

CPU: the host-side symbolic phase (device.tet_symbolic / tet_csr, plain torch ops) must give the
scipy-canonical CSR of the reference's triplets bit-exactly, and a Python emulation of
k_tet_assemble -- same corner order, same binary search, same row layout, same closed-form 3x3
blocks as csrc/tet.cu -- must reproduce the oracle's matrices; the mesh look-alikes must agree
with the oracle's restatement of volmdlr's TetrahedralElement.
GPU: fe_tet_elem_matrices / fe_tet_assemble and the drop-in API against goldens minted from the
reference's own ElasticityTetrahedralElement3D (oracle/make_golden.py --3d)."""
import numpy as np
import pytest
import torch

from oracle import numpy_oracle as no
from tests.fixtures import (Fixture, names3d, assert_close_rowscaled, assert_csr_values_close,
                            build_object_analysis)

TETS = names3d()
MAT2 = np.array([[210e9, 0.25, 1.0, 7860.0], [70e9, 0.33, 1.0, 2700.0]])


def test_goldens_exist():
    assert len(TETS) >= 4


def _geom(p):
    a, b, c = p[1] - p[0], p[2] - p[0], p[3] - p[0]
    c1, c2, c3 = np.cross(b, c), np.cross(c, a), np.cross(a, b)
    det = a @ c1
    g = np.zeros((4, 3))
    g[1], g[2], g[3] = c1 / det, c2 / det, c3 / det
    g[0] = -(g[1] + g[2] + g[3])
    return g, abs(det) / 6


def _block(kind, g, vol, m, v, j):
    """csrc/tet.cu: tet_material + tet_block."""
    e_mod, nu, rho = m[0], m[1], m[3]
    if kind == no.KIND_ELAST_TET:
        coeff = e_mod / ((1 + nu) * (1 - 2 * nu))
        p0, p1 = coeff * nu * vol, coeff * ((1 - 2 * nu) / 2) * vol
        return p0 * np.outer(g[v], g[j]) + p1 * np.outer(g[j], g[v]) + p1 * (g[v] @ g[j]) * np.eye(3)
    p0 = rho * vol / 20
    return (2 * p0 if v == j else p0) * np.eye(3)


@pytest.mark.parametrize("dims", [(2, 2, 2), (4, 2, 1)])
def test_symbolic_phase_and_kernel_emulation_vs_oracle(dims):
    from finite_elements_b200.device import tet_symbolic, tet_csr
    coords, conn = no.structured_tet_mesh(*dims, h=0.5, jitter=0.2, seed=1)
    n = len(coords)
    mid = (np.arange(len(conn)) % 2).astype(np.int32)
    cp, ce, ap, adj, deg = tet_symbolic(torch.as_tensor(conn).long(), n)
    rowptr, colidx = tet_csr(ap, adj, deg)
    cp, ce, ap, adj = cp.numpy(), ce.numpy(), ap.numpy(), adj.numpy()
    for i in range(n):   # per-node element lists: ascending, exactly the incident elements
        inc = np.nonzero((conn == i).any(axis=1))[0]
        assert np.array_equal(ce[cp[i]:cp[i + 1]], inc)
    for kind in (no.KIND_ELAST_TET, no.KIND_MASS_TET):
        ref = no.assemble_k(kind, coords, conn, mid, MAT2)
        assert np.array_equal(rowptr.numpy(), ref.indptr) and np.array_equal(colidx.numpy(), ref.indices)
        ke_ref = no.element_matrices(kind, coords, conn, mid, MAT2)
        vals = np.zeros(len(ref.data))
        for i in range(n):
            a0, d = ap[i], ap[i + 1] - ap[i]
            rows = vals[9 * a0:9 * a0 + 9 * d]
            for e in ce[cp[i]:cp[i + 1]]:
                nodes = conn[e]
                g, vol = _geom(coords[nodes])
                for v in range(4):
                    if nodes[v] != i:
                        continue
                    for j in range(4):
                        lo = int(np.searchsorted(adj[a0:a0 + d], nodes[j]))
                        blk = _block(kind, g, vol, MAT2[mid[e]], v, j)
                        assert np.abs(blk - ke_ref[e][3 * v:3 * v + 3, 3 * j:3 * j + 3]).max() \
                            <= 1e-13 * np.abs(ke_ref[e]).max()
                        for a in range(3):
                            rows[a * 3 * d + 3 * lo:a * 3 * d + 3 * lo + 3] += blk[a]
        assert np.abs(vals - ref.data).max() <= 1e-13 * np.abs(ref.data).max()


@pytest.mark.parametrize("name", TETS)
def test_oracle_post_processing_vs_reference(name):
    fx = Fixture(name)
    strain, stress, energy = no.post_tet(fx.coords, fx.conn, fx.mat_id, fx.mat, fx.ref("x")[:fx.ndof])
    assert_close_rowscaled(strain, fx.ref("strain"), 1e-12)
    assert_close_rowscaled(stress, fx.ref("stress"), 1e-12)
    assert np.max(np.abs(energy - fx.ref("energy"))) <= 1e-11 * np.max(np.abs(fx.ref("energy")))
    # the kernel's route to the energy: V/2 strain . stress
    vol, _ = no.tet_geometry(fx.coords, fx.conn)
    assert np.max(np.abs(0.5 * vol * np.sum(strain * stress, axis=1) - energy)) <= 1e-11 * np.max(np.abs(energy))


def test_mesh_lookalikes_vs_oracle():
    import finite_elements_b200 as fe
    coords, conn = fe.mesh.structured_tet_mesh(2, 1, 1, h=0.7, jitter=0.0)
    c2, t2 = no.structured_tet_mesh(2, 1, 1, h=0.7)
    assert np.array_equal(coords, c2) and np.array_equal(conn, t2)
    vol, form = no.tet_geometry(coords, conn)
    assert abs(vol.sum() - 2 * 0.7 ** 3) <= 1e-12           # the 6 tetrahedra tile each cell
    tet = fe.mesh.TetrahedralElement([fe.mesh.Node3D(*coords[i]) for i in conn[3]])
    assert abs(tet.volume - vol[3]) <= 1e-15
    assert np.allclose(np.array(tet.form_functions), form[3], rtol=1e-12, atol=1e-12)
    m = fe.mesh.ArrayMesh(coords, conn, 'elasticity3d', MAT2[:1])
    assert m.dimension == 3 and m.node_to_index[fe.mesh.Node3D(*coords[5])] == 5
    assert abs(m.element(3).area - vol[3]) <= 1e-15          # `.area` slot carries the volume
    flat = fe.mesh.flatten_mesh(m)
    assert flat['space_dim'] == 3 and flat['conn'].shape[1] == 4 and not flat['magnetic']


def test_host_element_records_and_flattening_vs_oracle():
    """Object meshes of ElasticityTetrahedralElement3D flatten to the fixture's arrays; the host-side
    B / D helpers (post-processing only) agree with the oracle's restatement of elements.py:719-797."""
    import finite_elements_b200 as fe
    fx = Fixture("tet_cube2_jit")
    m = fe.mesh
    nodes = [m.Node3D(*map(float, p)) for p in fx.coords]
    bounds = fx.meta["group_bounds"]
    groups = []
    for g in range(len(bounds) - 1):
        p = fx.mat[g]
        groups.append(m.ElementsGroup([fe.elements.ElasticityTetrahedralElement3D(
            m.TetrahedralElement([nodes[i] for i in fx.conn[e]]), p[0], p[1], p[3])
            for e in range(bounds[g], bounds[g + 1])], ''))
    mesh = m.Mesh(groups)
    mesh.nodes = nodes
    mesh.node_to_index = {nodes[i]: i for i in range(len(nodes))}
    flat = m.flatten_mesh(mesh)
    assert flat['space_dim'] == 3 and not flat['magnetic']
    assert np.array_equal(flat['coords'], fx.coords) and np.array_equal(flat['conn'], fx.conn)
    assert np.array_equal(flat['mat'][flat['mat_id']][:, [0, 1, 3]], fx.mat[fx.mat_id][:, [0, 1, 3]])
    bm, _ = no.b_matrix_tet(fx.coords, fx.conn)
    el = groups[1].elements[3]
    e = bounds[1] + 3
    assert el.dimension == 3
    assert np.max(np.abs(np.abs(el.b_matrix) - np.abs(bm[e]))) <= 1e-12 * np.abs(bm[e]).max()
    d_ref = no.d_matrix_tet(fx.mat[fx.mat_id[e], 0], fx.mat[fx.mat_id[e], 1])
    assert np.allclose(el.d_matrix(False, True), d_ref, rtol=1e-15)
    with pytest.raises(ValueError):
        el.d_matrix(True, True)
    an = fe.analysis.FiniteElementAnalysis(mesh, [], [], [], [], [], [], [], [], False, True)
    assert an.dimension == 3 and an.positions[(4, 3)] == 14
    rows, cols = an.get_row_col_indices(el)
    dofs = no.element_dofs(fx.conn[e:e + 1], 3)[0]
    assert rows == list(np.repeat(dofs, 12)) and cols == list(np.tile(dofs, 12))


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", TETS)
def test_gpu_tet_element_and_global_matrices_vs_reference(name):
    from finite_elements_b200.device import DeviceMesh3D, KIND_ELAST_TET, KIND_MASS_TET
    fx = Fixture(name)
    dm = DeviceMesh3D(fx.coords, fx.conn, fx.mat_id)
    ke = dm.element_matrices(KIND_ELAST_TET, fx.mat).cpu().numpy()
    me = dm.element_matrices(KIND_MASS_TET, fx.mat).cpu().numpy()
    assert_close_rowscaled(ke, fx.ref("ke"), 1e-12)
    assert_close_rowscaled(me, fx.ref("me"), 1e-13)
    kv = dm.assemble(KIND_ELAST_TET, fx.mat)
    k = dm.to_scipy(kv)
    assert_csr_values_close(k, fx.csr("k"), 1e-12)            # pattern bit-exact, values 1e-12
    assert_csr_values_close(dm.to_scipy(dm.assemble(KIND_MASS_TET, fx.mat)), fx.csr("m"), 1e-12)
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, fx.mat))   # deterministic
    # the node-owner kernels: tile == global accumulation, bitwise; the default (one lane per block, other
    # summation order and element labelling) agrees with them to rounding
    k1 = dm.assemble(KIND_ELAST_TET, fx.mat, variant=1)
    assert torch.equal(k1, dm.assemble(KIND_ELAST_TET, fx.mat, variant=2))
    # default = staged tiles, pipelined (6) = one tile per CTA (5); 4 = the same walk over a global gradient table: same
    # off-diagonal blocks bit for bit (same records, same order), the diagonal block is summed in another order
    k4 = dm.assemble(KIND_ELAST_TET, fx.mat, variant=4)
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, fx.mat, variant=5))
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, fx.mat, variant=6))
    assert torch.equal(dm.assemble(KIND_MASS_TET, fx.mat, variant=5), dm.assemble(KIND_MASS_TET, fx.mat, variant=6))
    assert float((kv - k4).abs().max()) <= 1e-14 * float(k4.abs().max())
    rp, ci = dm.csr_pattern()
    offdiag = (ci.long() // 3) != torch.repeat_interleave(torch.arange(dm.n_rows, device=ci.device) // 3,
                                                          (rp[1:] - rp[:-1]).long())
    assert torch.equal(kv[offdiag], k4[offdiag])
    k3 = dm.assemble(KIND_ELAST_TET, fx.mat, variant=3)
    assert float((k3 - k1).abs().max()) <= 1e-13 * float(k1.abs().max())
    assert float((kv - k1).abs().max()) <= 1e-13 * float(k1.abs().max())
    m1 = dm.assemble(KIND_MASS_TET, fx.mat, variant=1)
    assert float((dm.assemble(KIND_MASS_TET, fx.mat) - m1).abs().max()) <= 1e-14 * float(m1.abs().max())
    # the device-side symbolic phase (fe_tet_plan_create) against plain torch sorts of the same connectivity
    from finite_elements_b200.device import tet_symbolic, tet_csr
    cp, ce, ap, adj, deg = tet_symbolic(torch.as_tensor(fx.conn).long().cuda(), len(fx.coords))
    rp_t, ci_t = tet_csr(ap, adj, deg)
    rp_d, ci_d = dm.csr_pattern()
    assert torch.equal(rp_t, rp_d) and torch.equal(ci_t, ci_d)
    # the assembled matrix is the ordered sum of the dumped element matrices
    dofs = no.element_dofs(fx.conn, 3)
    import scipy.sparse as sp
    acc = sp.csr_matrix((ke.reshape(-1), (np.repeat(dofs, 12, axis=1).reshape(-1), np.tile(dofs, (1, 12)).reshape(-1))),
                        shape=k.shape)
    assert abs(acc - k).max() <= 1e-13 * abs(k).max()


@pytest.mark.gpu
@pytest.mark.parametrize("name", TETS)
def test_gpu_tet_api_solution_vs_reference_spsolve(name):
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx)
    assert an.dimension == 3
    m = an.create_matrix()
    assert_csr_values_close(m.tocsr(), fx.csr("kaug"), 1e-12)
    f = an.create_source_matrix()
    assert np.allclose(f, fx.ref("f"), rtol=1e-12, atol=1e-12 * np.abs(fx.ref("f")).max())
    x = np.array(an.solve().result_vector)
    ref = fx.ref("x")
    assert x.shape == ref.shape
    assert np.linalg.norm(x[:fx.ndof] - ref[:fx.ndof]) <= 1e-8 * np.linalg.norm(ref[:fx.ndof])
    lam_scale = np.abs(ref[fx.ndof:]).max()
    assert np.abs(x[fx.ndof:] - ref[fx.ndof:]).max() <= 1e-6 * lam_scale
    ke = elems[1].elementary_matrix(an.plane_strain, an.plane_stress)
    assert_close_rowscaled(np.asarray(ke)[None], fx.ref("ke")[1:2], 1e-12)
    assert_close_rowscaled(np.asarray(elems[1].elementary_mass_matrix())[None], fx.ref("me")[1:2], 1e-13)
    with pytest.raises(ValueError):
        type(an)(mesh, [], [], [], [], [], [], [], [], plane_strain=True, plane_stress=True).create_matrix()


@pytest.mark.gpu
@pytest.mark.parametrize("name", TETS)
def test_gpu_tet_results_vs_reference(name):
    import finite_elements_b200 as fe
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx)
    ref_x = fx.ref("x")
    res = fe.results.ElasticityResults3D(mesh, list(ref_x), an.plane_strain, an.plane_stress)
    assert res.strain_array.shape == (len(fx.conn), 6)
    assert_close_rowscaled(res.strain_array, fx.ref("strain"), 1e-11)
    assert_close_rowscaled(res.stress_array, fx.ref("stress"), 1e-11)
    energy = fx.ref("energy")
    assert np.max(np.abs(res.energy_array - energy)) <= 1e-10 * np.max(np.abs(energy))
    assert abs(res.energy - energy.sum()) <= 1e-10 * abs(energy.sum())
    d = res.displacement_vectors_per_node[mesh.nodes[1]]
    assert (d.x, d.y, d.z) == tuple(fx.ref("disp_node1"))
    assert res.displacements_per_element[elems[0]] == [ref_x[3 * n + k] for n in fx.conn[0] for k in (0, 1, 2)]
    assert np.allclose(res.strain[elems[1]], fx.ref("strain")[1], rtol=1e-9, atol=1e-11 * abs(fx.ref("strain")).max())
    assert res.shear_stress_zx()[elems[2]] == res.stress_array[2, 5]
    assert res.axial_strain_z()[elems[2]] == res.strain_array[2, 2] and len(res.displacement_per_node_z()) == len(fx.coords)


@pytest.mark.gpu
def test_gpu_tet_modal_largest_vs_reference():
    fx = Fixture("tet_beam6x2x2_jit")
    an, _, _ = build_object_analysis(fx)
    ref = fx.ref("eig_largest")
    vals, vecs = an.modal_analysis("largest", len(ref))
    assert np.max(np.abs(vals - ref)) <= 1e-8 * ref.max()
    assert vecs.shape == (len(ref), fx.ndof)


@pytest.mark.gpu
def test_gpu_tet_mid_size_properties():
    """24 x 12 x 12 cells (20 736 tetrahedra, 4225 nodes): oracle parity at a size the dense paths
    cannot fake, the 6 rigid-body modes, symmetry, total mass, PCG to a true residual of 1e-10."""
    import finite_elements_b200 as fe
    from finite_elements_b200.device import DeviceMesh3D, KIND_ELAST_TET, KIND_MASS_TET
    nx, ny, nz, h = 24, 12, 12, 0.25
    coords, conn = fe.mesh.structured_tet_mesh(nx, ny, nz, h=h, jitter=0.2, seed=2)
    mid = (np.arange(len(conn)) % 2).astype(np.int32)
    dm = DeviceMesh3D(coords, conn, mid)
    kv = dm.assemble(KIND_ELAST_TET, MAT2)
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, MAT2))
    k2 = dm.assemble(KIND_ELAST_TET, MAT2, variant=2)
    assert torch.equal(k2, dm.assemble(KIND_ELAST_TET, MAT2, variant=1))
    assert float((kv - k2).abs().max()) <= 1e-13 * float(k2.abs().max())
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, MAT2, variant=6))        # the default is the pipelined staged variant
    assert torch.equal(kv, dm.assemble(KIND_ELAST_TET, MAT2, variant=5))
    assert float((kv - dm.assemble(KIND_ELAST_TET, MAT2, variant=4)).abs().max()) <= 1e-14 * float(k2.abs().max())
    k = dm.to_scipy(kv)
    assert_csr_values_close(k, no.assemble_k(no.KIND_ELAST_TET, coords, conn, mid, MAT2), 1e-12)
    mv = dm.assemble(KIND_MASS_TET, MAT2)
    assert_csr_values_close(dm.to_scipy(mv), no.assemble_k(no.KIND_MASS_TET, coords, conn, mid, MAT2), 1e-12)
    scale = float(kv.abs().max())
    n = len(coords)
    c = torch.as_tensor(coords).cuda()
    modes = []
    for d in range(3):
        t = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
        t[:, d] = 1.0
        modes.append(t.reshape(-1))
    for a, b in ((0, 1), (1, 2), (2, 0)):   # infinitesimal rotations
        r = torch.zeros(n, 3, dtype=torch.float64, device="cuda")
        r[:, a], r[:, b] = -c[:, b], c[:, a]
        modes.append(r.reshape(-1).contiguous())
    for v in modes:
        assert float(dm.spmv(kv, v).abs().max()) <= 1e-11 * scale * max(1.0, float(c.abs().max()))
    ones_x = modes[0]
    vol, _ = no.tet_geometry(coords, conn)
    total = float(torch.dot(ones_x, dm.spmv(mv, ones_x)))
    assert abs(total - float((MAT2[mid, 3] * vol).sum())) <= 1e-10 * total
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.rand(3 * n, dtype=torch.float64, device="cuda", generator=g)
    y = torch.rand(3 * n, dtype=torch.float64, device="cuda", generator=g)
    assert abs(float(torch.dot(x, dm.spmv(kv, y))) - float(torch.dot(y, dm.spmv(kv, x)))) \
        <= 1e-12 * abs(float(torch.dot(x, dm.spmv(kv, y))))
    left = np.nonzero(coords[:, 0] == 0)[0]
    bc = (3 * left[:, None] + np.arange(3)[None, :]).reshape(-1)
    f = torch.zeros(3 * n, dtype=torch.float64, device="cuda")
    tip = np.nonzero(coords[:, 0] == nx * h)[0]
    f[torch.as_tensor(3 * tip + 2).cuda()] = -1000.0
    rhs = f.clone()
    dm.dirichlet(kv, rhs, bc, np.zeros(len(bc)))
    u, iters, relres = dm.pcg(kv, rhs, rtol=1e-10)
    true = float(torch.linalg.norm(rhs - dm.spmv(kv, u)) / torch.linalg.norm(rhs))
    assert relres <= 1e-10 and true <= 1.001e-10, (iters, relres, true)
    # against the oracle's direct solve of the same reduced system
    kk = no.assemble_k(no.KIND_ELAST_TET, coords, conn, mid, MAT2)
    fr = np.zeros(3 * n)
    fr[3 * tip + 2] = -1000.0
    ref = no.solve_reduced_direct(kk, fr, bc, np.zeros(len(bc)), permc_spec='COLAMD')
    assert np.linalg.norm(u.cpu().numpy() - ref) <= 1e-8 * np.linalg.norm(ref)
