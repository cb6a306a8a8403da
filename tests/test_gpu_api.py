"""GPU: the drop-in class surface (finite_elements_b200.analysis / elements / results) driven the
way the reference's scripts drive the reference, against fixtures minted from the reference."""
import numpy as np
import pytest

from tests.fixtures import Fixture, names, build_object_analysis, assert_close_rowscaled, assert_csr_values_close

pytestmark = pytest.mark.gpu

SMALL = [n for n in names() if n not in ("gmsh_beam_0.1", "gmsh_beam_0.18")]  # object meshes: keep it quick


@pytest.mark.parametrize("name", SMALL)
def test_create_matrix_source_and_solve(name):
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx, with_element_records=True)
    m = an.create_matrix()
    assert_csr_values_close(m, fx.csr("kaug"), 1e-12)          # pattern bit-exact incl. Lagrange rows
    f = an.create_source_matrix()
    ref_f = fx.ref("f")
    assert f.shape == ref_f.shape
    assert np.allclose(f, ref_f, rtol=1e-12, atol=1e-12 * np.abs(ref_f).max())
    result = an.solve()
    x, ref_x = np.array(result.result_vector), fx.ref("x")
    assert isinstance(result.result_vector, list) and len(x) == len(ref_x) and result.mesh is mesh
    un = np.linalg.norm(ref_x[:fx.ndof])
    assert np.linalg.norm(x[:fx.ndof] - ref_x[:fx.ndof]) <= 1e-8 * un
    assert np.linalg.norm(x[fx.ndof:] - ref_x[fx.ndof:]) <= 1e-6 * np.linalg.norm(ref_x[fx.ndof:])
    assert result.dimension == fx.dim and an.last_solve_info["iterations"] > 0


@pytest.mark.parametrize("name", ["plate2_pstress", "beam18_pstrain", "semantics_elast", "magbar18"])
def test_k_m_views_and_element_methods(name):
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx)
    ks = an.k_matrix('sparse')
    assert ks.format == 'csc'
    assert_csr_values_close(ks.tocsr(), fx.csr("k"), 1e-12)
    assert np.allclose(an.k_matrix('dense'), fx.csr("k").toarray(), rtol=0, atol=1e-12 * abs(fx.csr("k")).max())
    data, rows, cols = an.k_matrix_data()
    nd = 3 * fx.dim
    assert len(data) == len(rows) == len(cols) == len(elems) * nd * nd
    assert_close_rowscaled(np.array(data).reshape(len(elems), -1), fx.ref("ke"), 1e-12)
    r5, c5 = an.get_row_col_indices(elems[-1])
    assert rows[-nd * nd:] == r5 and cols[-nd * nd:] == c5
    if fx.kind == "elasticity":
        assert_csr_values_close(an.m_matrix('sparse').tocsr(), fx.csr("m"), 1e-12)
        ke = elems[1].elementary_matrix(an.plane_strain, an.plane_stress)
        assert ke.shape == (36,)
        assert_close_rowscaled(ke[None], fx.ref("ke")[1:2], 1e-12)
        assert_close_rowscaled(elems[1].elementary_mass_matrix()[None], fx.ref("me")[1:2], 1e-12)
        with pytest.raises(ValueError):
            elems[0].elementary_matrix(True, True)
    else:
        ke = elems[1].elementary_matrix()
        assert isinstance(ke, tuple) and len(ke) == 9
        assert_close_rowscaled(np.array(ke)[None], fx.ref("ke")[1:2], 1e-12)
    fac = elems[0].element_to_node_factors()
    assert np.allclose(fac[1:], fx.ref("factors")[0][1:], rtol=1e-12)


@pytest.mark.parametrize("name", ["plate2_pstress", "beam18_pstrain", "semantics_elast", "gmsh_beam_0.3",
                                  "struct24x16_jit_pstress"])
def test_elasticity_results_vs_reference(name):
    import finite_elements_b200 as fe
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx, with_element_records=True)
    ref_x = fx.ref("x")
    res = fe.results.ElasticityResults2D(mesh, list(ref_x), an.plane_strain, an.plane_stress)
    assert_close_rowscaled(res.strain_array, fx.ref("strain"), 1e-11)
    assert_close_rowscaled(res.stress_array, fx.ref("stress"), 1e-11)
    energy = fx.ref("energy")
    assert np.max(np.abs(res.energy_array - energy)) <= 1e-10 * np.max(np.abs(energy))
    assert abs(res.energy - energy.sum()) <= 1e-10 * abs(energy.sum())
    # dictionary-shaped accessors of the reference
    assert np.allclose(res.strain[elems[1]], fx.ref("strain")[1], rtol=1e-9, atol=1e-11 * abs(fx.ref("strain")).max())
    assert np.allclose(elems[1].stress, res.stress[elems[1]])
    d = res.displacement_vectors_per_node[mesh.nodes[1]]
    assert (d.x, d.y) == (ref_x[2], ref_x[3])
    assert res.displacements_per_element[elems[0]] == [ref_x[2 * n + k] for n in fx.conn[0] for k in (0, 1)]
    assert res.axial_stress_x()[1] == res.stress_array[1, 0] and len(res.shear_strain_xy()) == len(elems)
    assert abs(res.energy_per_element[elems[0]] - energy[0]) <= 1e-10 * abs(energy).max()
    # the solver's own solution gives the same fields
    res2 = fe.results.ElasticityResults2D(mesh, an.solve().result_vector, an.plane_strain, an.plane_stress)
    assert np.max(np.abs(res2.stress_array - fx.ref("stress"))) <= 1e-6 * np.max(np.abs(fx.ref("stress")))


@pytest.mark.parametrize("name", ["magbar18", "semantics_mag", "struct24x16_jit_mag"])
def test_magnetic_results_vs_reference(name):
    import finite_elements_b200 as fe
    fx = Fixture(name)
    an, mesh, elems = build_object_analysis(fx, with_element_records=True)
    res = fe.results.MagneticResults(mesh, list(fx.ref("x")))
    assert_close_rowscaled(res.magnetic_field_array, fx.ref("bfield"), 1e-11)
    b = res.magnetic_field_per_element[elems[3]]
    assert np.allclose([b.x, b.y], fx.ref("bfield")[3], rtol=1e-10, atol=1e-11 * abs(fx.ref("bfield")).max())
    assert np.allclose(res.magnetic_field_norm, np.hypot(*fx.ref("bfield").T), rtol=1e-10)


def test_array_mesh_path_matches_object_path():
    import finite_elements_b200 as fe
    fx = Fixture("semantics_elast")
    mesh = fe.mesh.ArrayMesh(fx.coords, fx.conn, 'elasticity', fx.mat, fx.meta["group_bounds"])

    class Edge:
        def __init__(self, a, b):
            self.start, self.end = a, b

    an = fe.analysis.FiniteElementAnalysis(
        mesh,
        [fe.loads.ElementsLoad([mesh.element(j) for j in idx], v, d) for idx, v, d in fx.rec("elements_loads")],
        [fe.loads.EdgeLoad(Edge(a, b), v, d) for a, b, v, d in fx.rec("edge_loads")],
        [fe.loads.NodeLoad(n, v, d) for n, v, d in fx.rec("node_loads")], [], [],
        [fe.conditions.NodeBoundaryCondition(n, v, d) for n, v, d in fx.rec("node_bcs")],
        [fe.conditions.EdgeBoundaryCondition(Edge(a, b), v, d) for a, b, v, d in fx.rec("edge_bcs")],
        [fe.conditions.ElementBoundaryCondition(mesh.element(j), v, d) for j, v, d in fx.rec("element_bcs")],
        plane_strain=False, plane_stress=True)
    assert_csr_values_close(an.create_matrix(), fx.csr("kaug"), 1e-12)
    x = an.solve_arrays()
    ref_x = fx.ref("x")
    assert np.linalg.norm(x[:fx.ndof] - ref_x[:fx.ndof]) <= 1e-8 * np.linalg.norm(ref_x[:fx.ndof])
    res = fe.results.ElasticityResults2D(mesh, x, False, True)
    assert np.max(np.abs(res.stress_array - fx.ref("stress"))) <= 1e-6 * np.max(np.abs(fx.ref("stress")))
    with pytest.raises(TypeError):
        res.strain  # no element objects on an ArrayMesh


def test_singular_system_raises_like_the_reference():
    """No boundary condition at all: the reference's spsolve raises MatrixRankWarning ->
    NotImplementedError (analysis.py:824-826).  Here: PCG breakdown or an unattainable residual."""
    import finite_elements_b200 as fe
    fx = Fixture("plate2_pstress")
    an, mesh, elems = build_object_analysis(fx)
    an.node_boundary_conditions = []
    an.solver_maxit = 200
    with pytest.raises((NotImplementedError, fe._lib.NotConverged)):
        an.solve()


# ------------------------------------------------------------------ BASELINE configs[0] "as shipped"
_TIP_UY = {"0.8": -0.0842551, "0.5": -0.0926325, "0.3": -0.119740, "0.18": -0.127914, "0.1": -0.132526}


@pytest.mark.parametrize("lc", list(_TIP_UY))
def test_beam2d_example_3_from_msh_file(lc):
    """scripts/Elasticity/beam2d_example_3.py:52-121 driven from a .msh file through the drop-in classes
    exactly as the script does (GmshParser -> element objects -> Mesh -> keep gmsh node order -> NodeLoad
    at (10, 1), clamp x = 0, plane stress): solution vs the reference's spsolve (golden), tip deflection
    vs the values the reference gives for the five meshes."""
    import os
    import finite_elements_b200 as fe
    vmmesh = fe.mesh
    fx = Fixture("gmsh_beam_" + lc)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "msh", f"gmsh_beam_{lc}.msh")
    gmsh = fe.mesh.GmshParser.from_file(path)
    mesh = gmsh.define_triangular_element_mesh()
    group_elements = []
    for group in mesh.elements_groups:
        solid = [fe.elements.ElasticityTriangularElement2D(tri, 30 * 1e6, 0.25, 2.7, 1) for tri in group.elements]
        group_elements.append(vmmesh.ElementsGroup(solid, ''))
    mesh = vmmesh.Mesh(group_elements)
    mesh.nodes = gmsh.nodes['all_nodes']                                   # keep gmsh order (:72-73)
    mesh.node_to_index = {mesh.nodes[i]: i for i in range(len(mesh.nodes))}
    tip = mesh.node_to_index[vmmesh.Node2D(10, 1)]
    node_loads = [fe.loads.NodeLoad(mesh.nodes[tip], -1000, 2)]
    bcs = []
    for i, node in enumerate(mesh.nodes):
        if node.x == 0:
            bcs += [fe.conditions.NodeBoundaryCondition(mesh.nodes[i], 0, 1),
                    fe.conditions.NodeBoundaryCondition(mesh.nodes[i], 0, 2)]
    an = fe.analysis.FiniteElementAnalysis(mesh=mesh, element_loads=[], edge_loads=[], node_loads=node_loads,
                                           magnet_loads=[], continuity_conditions=[], node_boundary_conditions=bcs,
                                           edge_boundary_conditions=[], element_boundary_conditions=[],
                                           plane_strain=False, plane_stress=True)
    m = an.create_matrix()
    assert_csr_values_close(m, fx.csr("kaug"), 1e-12)
    results = an.solve()
    x, ref_x = np.array(results.result_vector), fx.ref("x")
    assert len(x) == len(ref_x)
    assert np.linalg.norm(x[:fx.ndof] - ref_x[:fx.ndof]) <= 1e-8 * np.linalg.norm(ref_x[:fx.ndof])
    er = fe.results.ElasticityResults2D(results.mesh, results.result_vector, False, True)
    assert abs(er.displacement_vectors_per_node[mesh.nodes[tip]][1] - _TIP_UY[lc]) < 2e-6
