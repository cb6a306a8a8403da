"""CPU: the numpy oracle (oracle/numpy_oracle.py) against fixtures minted from the
reference's own unmodified code (tests/golden, oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import numpy_oracle as no
from tests.fixtures import Fixture, names, names3d, assert_close_rowscaled, assert_csr_values_close

ALL = names() + names3d()


@pytest.mark.parametrize("name", ALL)
def test_element_matrices(name):
    fx = Fixture(name)
    ke = no.element_matrices(fx.oracle_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
    assert_close_rowscaled(ke.reshape(len(fx.conn), -1), fx.ref("ke"), 1e-13)
    if fx.kind in ("elasticity", "elasticity3d"):
        mass_kind = no.KIND_MASS if fx.kind == "elasticity" else no.KIND_MASS_TET
        me = no.element_matrices(mass_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
        assert_close_rowscaled(me.reshape(len(fx.conn), -1), fx.ref("me"), 1e-14)


@pytest.mark.parametrize("name", ALL)
def test_source_factors(name):
    fx = Fixture(name)
    if fx.kind == "elasticity3d":
        pytest.skip("the reference defines no element_to_node_factors for tetrahedra")
    fac = no.element_to_node_factors(fx.coords, fx.conn)
    ref = fx.ref("factors")
    area = no.tri_area(fx.coords, fx.conn)
    # first factor is ~0 by cancellation (SURVEY a-6): absolute tolerance scaled by |det| * max|coord|
    tol = 1e-13 * (2 * area) * max(1.0, np.abs(fx.coords).max()) ** 2 / np.minimum(1.0, 2 * area)
    assert np.all(np.abs(fac - ref) <= tol[:, None] + 1e-13 * np.abs(ref))
    assert np.allclose(fac[:, 1:], area[:, None], rtol=1e-9)


@pytest.mark.parametrize("name", ALL)
def test_k_block_pattern_and_values(name):
    fx = Fixture(name)
    k = no.assemble_k(fx.oracle_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
    assert_csr_values_close(k, fx.csr("k"), 1e-13)
    if fx.kind in ("elasticity", "elasticity3d"):
        mass_kind = no.KIND_MASS if fx.kind == "elasticity" else no.KIND_MASS_TET
        m = no.assemble_k(mass_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
        assert_csr_values_close(m, fx.csr("m"), 1e-13)


@pytest.mark.parametrize("name", ALL)
def test_augmented_system_rhs_and_solution(name):
    fx = Fixture(name)
    k = no.assemble_k(fx.oracle_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
    bc_dofs, bc_vals = no.bcs_to_dof_records(fx.coords, fx.conn, fx.dim, fx.rec("node_bcs"),
                                             fx.rec("element_bcs"), fx.rec("edge_bcs"))
    assert np.array_equal(bc_dofs, fx.ref("bc_dofs"))
    # element BCs carry the ~0 first factor (cancellation noise, SURVEY a-6) -> scaled atol
    assert np.allclose(bc_vals, fx.ref("bc_vals"), rtol=1e-13, atol=1e-12 * max(np.abs(bc_vals).max(), 1e-300))
    kaug = no.augmented_matrix(k, bc_dofs)
    assert_csr_values_close(kaug, fx.csr("kaug"), 1e-13)
    ld, lv = no.loads_to_dof_records(fx.coords, fx.conn, fx.dim, fx.rec("node_loads"),
                                     fx.rec("elements_loads"), fx.rec("edge_loads"))
    f = no.source_vector(fx.ndof, ld, lv, bc_vals)
    if fx.records.get("magnet_loads"):   # analysis.py:556-577: added on top, no de-duplication
        mr, mv = no.magnet_load_records(fx.coords, fx.conn, fx.rec("magnet_loads"))
        assert np.array_equal(mr, fx.ref("magnet_rows"))
        assert np.allclose(mv, fx.ref("magnet_data"), rtol=1e-13, atol=1e-13 * np.abs(mv).max())
        np.add.at(f[:, 0], mr, mv)
    ref_f = fx.ref("f")
    assert f.shape == ref_f.shape
    assert np.allclose(f, ref_f, rtol=1e-12, atol=1e-12 * np.abs(ref_f).max())
    x = no.solve_augmented(kaug, f)
    ref_x = fx.ref("x")
    un = np.linalg.norm(ref_x[:fx.ndof])
    assert np.linalg.norm(x[:fx.ndof] - ref_x[:fx.ndof]) <= 1e-9 * un
    # the Dirichlet-eliminated SPD solve equals the reference's Lagrange solve on the solution block
    u = no.solve_reduced_direct(k, f[:fx.ndof, 0], bc_dofs, bc_vals)
    assert np.linalg.norm(u - ref_x[:fx.ndof]) <= 1e-8 * un
    lam = no.multipliers(k, u, f[:fx.ndof, 0], bc_dofs)
    ln = np.linalg.norm(ref_x[fx.ndof:])
    assert np.linalg.norm(lam - ref_x[fx.ndof:]) <= 1e-7 * max(ln, 1e-300)


def test_known_answers_from_reference_scripts():
    """Values quoted in SURVEY.md §4 (probe of beam2d_example_1.py / _3.py / finite_element_beam.py)."""
    fx = Fixture("plate2_pstress")
    assert np.allclose(fx.ref("ke")[0][:6], [9833333.33333333, -5e6, -4.5e6, 2e6, -5333333.33333333, 3e6])
    assert list(fx.ref("kaug_indptr")) == [0, 6, 13, 21, 29, 38, 47, 54, 61, 62, 63, 64, 65, 66]
    assert np.allclose(fx.ref("x")[[0, 2, 3]], [1.907739e-05, 8.730330e-06, -7.415391e-05], rtol=1e-6)
    tips = {"0.8": -0.0842551, "0.5": -0.0926325, "0.3": -0.119740, "0.18": -0.127914, "0.1": -0.132526}
    for lc, uy in tips.items():
        fx = Fixture("gmsh_beam_" + lc)
        tip = int(np.where((fx.coords[:, 0] == 10) & (fx.coords[:, 1] == 1))[0][0])
        assert abs(fx.ref("x")[2 * tip + 1] - uy) < 2e-6
    fx = Fixture("magbar18")
    f = fx.ref("f")[:, 0]
    # last-wins: nodes (1,0) and (0,1) get 2.5e9 each, total 5e9, not 1e10 (SURVEY a-7)
    assert np.isclose(f[:fx.ndof].sum(), 5e9) and np.isclose(f[1], 2.5e9) and np.isclose(f[10], 2.5e9)


def test_dedup_is_last_wins_not_sum():
    d, v = no.dedup_last_wins([4, 7, 4], [5.0, 1.0, 7.0])
    assert list(d) == [4, 7] and list(v) == [7.0, 1.0]


def test_plane_flag_errors():
    with pytest.raises(ValueError):
        no.d_matrix(1.0, 0.3, True, True)
    with pytest.raises(ValueError):
        no.d_matrix(1.0, 0.3, False, False)


def test_jacobi_pcg_matches_direct():
    fx = Fixture("struct24x16_jit_pstress")
    k = no.assemble_k(fx.oracle_kind, fx.coords, fx.conn, fx.mat_id, fx.mat)
    ke, b = no.eliminate_dirichlet(k, fx.ref("f")[:fx.ndof, 0], fx.ref("bc_dofs"), fx.ref("bc_vals"))
    x, it, rel = no.jacobi_pcg(ke, b, rtol=1e-13)
    ref = fx.ref("x")[:fx.ndof]
    assert rel <= 1e-13 and np.linalg.norm(x - ref) <= 1e-8 * np.linalg.norm(ref)
