"""CPU: the C-ABI library loads and exports what include/fe_b200.h declares; host-side
logic of the drop-in layer (no compute calls -- there is no GPU here and no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from tests.fixtures import Fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "fe_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fe_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import finite_elements_b200._lib as L
    names = _declared_symbols()
    assert len(names) >= 24
    raw = ctypes.CDLL(L.LIB_PATH)
    for name in names:
        assert hasattr(raw, name), f"{name} is declared in include/fe_b200.h but not exported"
        assert name in L.SIGNATURES, f"{name} has no ctypes signature in _lib.py"
    assert sorted(L.SIGNATURES) == names
    assert L.lib.fe_version() == 100
    assert L.lib.fe_pcg_work_len(10, 14) == 64   # 5 n_rows + n_cols


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from finite_elements_b200.device import Context, DeviceMesh
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Context(0)
    with pytest.raises(RuntimeError):
        DeviceMesh(np.zeros((3, 2)), np.array([[0, 1, 2]], np.int32))
    import finite_elements_b200._lib as L
    handle = ctypes.c_void_p()
    rc = L.lib.fe_ctx_create(0, ctypes.byref(handle))
    assert rc != 0 and L.last_error()
    with pytest.raises(Exception):
        L.check(rc)


def test_error_code_mapping():
    import finite_elements_b200._lib as L
    L.lib.fe_ctx_create(0, None)  # sets "out is NULL" -> FE_ERR_ARG
    with pytest.raises(ValueError):
        L.check(L.FE_ERR_ARG)
    with pytest.raises(NotImplementedError):
        L.check(L.FE_ERR_BREAKDOWN)
    with pytest.raises(L.NotConverged):
        L.check(L.FE_ERR_NOT_CONVERGED)
    with pytest.raises(RuntimeError):
        L.check(L.FE_ERR_CUDA)


# ------------------------------------------------------------------ mesh utilities
def test_structured_mesh_matches_oracle_generator():
    from finite_elements_b200.mesh import structured_mesh
    from oracle import numpy_oracle as no
    for nx, ny, jit in ((9, 1, 0.0), (24, 16, 0.2), (5, 7, 0.1)):
        c1, t1 = structured_mesh(nx, ny, jitter=jit, seed=4)
        c2, t2 = no.structured_mesh(nx, ny, jitter=jit, seed=4)
        assert np.array_equal(t1, t2) and np.array_equal(c1, c2)


def test_structured_mesh_torch_matches_numpy():
    import torch
    from finite_elements_b200.mesh import structured_mesh, structured_mesh_torch
    c1, t1 = structured_mesh(7, 5)
    c2, t2 = structured_mesh_torch(7, 5, torch.device("cpu"))
    assert np.array_equal(t1, t2.numpy()) and np.allclose(c1, c2.numpy(), rtol=0, atol=0)
    _, t3 = structured_mesh_torch(7, 5, torch.device("cpu"), row_lo=2, row_hi=4)
    assert np.array_equal(t3.numpy(), t1[2 * 7 * 2:2 * 7 * 4])


def test_read_gmsh41(tmp_path):
    from finite_elements_b200.mesh import read_gmsh41
    msh = """$MeshFormat
4.1 0 8
$EndMeshFormat
$Nodes
2 4 1 4
0 1 0 2
1
2
0 0 0
1 0 0
2 1 0 2
3
4
1 1 0
0 1 0
$EndNodes
$Elements
2 3 1 3
1 1 1 1
1 1 2
2 1 2 2
2 1 2 3
3 1 3 4
$EndElements
"""
    p = tmp_path / "t.msh"
    p.write_text(msh)
    coords, conn = read_gmsh41(str(p))
    assert np.array_equal(coords, [[0, 0], [1, 0], [1, 1], [0, 1]])
    assert np.array_equal(conn, [[0, 1, 2], [0, 2, 3]])
    (tmp_path / "bad.msh").write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError):
        read_gmsh41(str(tmp_path / "bad.msh"))


def test_mesh_lookalike_first_seen_order_and_approx_equality():
    import finite_elements_b200 as fe
    m = fe.mesh
    tris = [m.TriangularElement2D([m.Node2D(3, 0), m.Node2D(3, 2), m.Node2D(0, 0)]),
            m.TriangularElement2D([m.Node2D(0, 2), m.Node2D(0, 0), m.Node2D(3, 2)])]
    mesh = m.Mesh([m.ElementsGroup(tris, '')])
    assert [(n.x, n.y) for n in mesh.nodes] == [(3, 0), (3, 2), (0, 0), (0, 2)]
    assert mesh.node_to_index[m.Node2D(3 + 1e-9, 2)] == 1
    assert tris[0].area == 3.0
    a, b, c = zip(*tris[0].form_functions)
    for i, p in enumerate(tris[0].points):  # N_i(p_j) = delta_ij
        vals = [a[k] + b[k] * p.x + c[k] * p.y for k in range(3)]
        assert np.allclose(vals, np.eye(3)[i])
    assert len(tris[0].linear_elements) == 3


# ------------------------------------------------------------------ host logic of the analysis
from tests.fixtures import build_object_analysis as _object_analysis  # noqa: E402


def test_flatten_positions_and_triplet_indices():
    from oracle import numpy_oracle as no
    fx = Fixture("semantics_elast")
    an, mesh, elems = _object_analysis(fx)
    flat = an._flatten()
    assert np.array_equal(flat["conn"], fx.conn) and np.array_equal(flat["coords"], fx.coords)
    assert np.array_equal(flat["mat"][flat["mat_id"]], fx.mat[fx.mat_id])
    assert an.dimension == 2 and an.positions[(3, 2)] == 7 and len(an.positions) == 2 * len(fx.coords)
    with pytest.raises(KeyError):
        an.positions[(0, 3)]
    rows, cols = no.triplet_indices(fx.conn, 2)
    r, c = an.get_row_col_indices(elems[5])
    assert r == list(rows[5]) and c == list(cols[5])


def test_last_wins_dedup_of_loads_and_conditions_matches_reference():
    """Node + edge records only (element records need the device): the (dof, value) lists
    must equal the reference's (golden bc_dofs / bc_vals restricted to those records)."""
    from oracle import numpy_oracle as no
    fx = Fixture("semantics_elast")
    an, _, _ = _object_analysis(fx)
    dofs, vals = an._bc_arrays()
    ed, ev = no.bcs_to_dof_records(fx.coords, fx.conn, 2, fx.rec("node_bcs"), (), fx.rec("edge_bcs"))
    assert np.array_equal(dofs, ed) and np.array_equal(vals, ev)
    data, rows = an.source_c_matrix_loads()
    ld, lv = no.loads_to_dof_records(fx.coords, fx.conn, 2, fx.rec("node_loads"), (), fx.rec("edge_loads"))
    assert rows == list(ld) and np.array_equal(data, lv)
    # duplicate key: last value, first position
    assert dict(zip(rows, data))[2 * (4 * 7 + 6) + 1] == -400.0
    d, r, c = an.c_matrix_boundary_conditions()
    ndof = 2 * len(fx.coords)
    assert d == [1, 1] * len(dofs) and r[0::2] == [ndof + i for i in range(len(dofs))] and c[0::2] == list(dofs)
    src, srow = an.source_c_matrix_boundary_conditions()
    assert srow == [ndof + i for i in range(len(dofs))] and np.array_equal(src, vals)


def test_array_mesh_surface():
    import finite_elements_b200 as fe
    fx = Fixture("magbar18")
    mesh = fe.mesh.ArrayMesh(fx.coords, fx.conn, 'magnetic', fx.mat, fx.meta["group_bounds"])
    assert mesh.dimension == 1 and len(mesh.nodes) == 20 and mesh.node_to_index[7] == 7
    assert mesh.node_to_index[fe.mesh.Point2D(3.0, 1.0)] == 13
    assert np.array_equal(mesh.mat_id, fx.mat_id)
    el = mesh.element(3)
    assert int(el) == 3 and el.area == 0.5 and el.points == [int(v) for v in fx.conn[3]]
    load = fe.loads.ElementsLoad([mesh.element(0), mesh.element(1)], 1e10, 1)
    assert load.value_per_element == [5e9, 5e9]
    groups = mesh.elements_groups
    assert len(groups) == 3 and groups[1].elements[0].mu_total == fx.mat[1, 0]


def test_api_errors_without_device():
    import finite_elements_b200 as fe
    fx = Fixture("plate2_pstress")
    an, _, elems = _object_analysis(fx)
    an.plane_stress, an.plane_strain = True, True
    with pytest.raises(ValueError):
        an._kind()
    an.plane_stress, an.plane_strain = False, False
    with pytest.raises(ValueError):
        an._kind()
    with pytest.raises(ValueError):
        elems[0].d_matrix(True, True)
    with pytest.raises(NotImplementedError):
        an.k_matrix('banded')
    with pytest.raises(NotImplementedError):
        an.m_matrix('banded')
    an.continuity_conditions = [fe.conditions.ContinuityCondition(0, 1, 1)]
    with pytest.raises(NotImplementedError):
        an.c_matrix_continuity_conditions()
    # post-processing helpers agree with the oracle's B and D
    from oracle import numpy_oracle as no
    b, _ = no.b_matrix(fx.coords, fx.conn)
    assert np.allclose(elems[1].b_matrix, b[1], rtol=1e-15)
    assert np.allclose(elems[0].d_matrix(False, True), no.d_matrix(30e6, 0.25, False, True))


def test_magnet_load_rhs_matches_reference():
    """MagnetLoad -> right-hand side (loads.py:129-147, analysis.py:556-577) against the reference's own
    (data, rows) and source vector; pure host work, no device call (node BCs + magnet loads only)."""
    fx = Fixture("semantics_mag_magnet")
    an, mesh, elems = _object_analysis(fx)          # element records (ElementsLoad) need the device: left out
    data, rows = an.source_c_matrix_magnet_loads()
    assert rows == list(fx.ref("magnet_rows"))
    ref = fx.ref("magnet_data")
    assert np.allclose(data, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    contour = an.magnet_loads[1].contour_linear_elements()
    assert len(contour) == 3                         # the edge between the two non-contour nodes dropped out
    f = an.create_source_matrix()
    expect = np.zeros_like(f)
    np.add.at(expect[:, 0], fx.ref("magnet_rows"), ref)
    assert f.shape == fx.ref("f").shape and np.allclose(f, expect, rtol=1e-13, atol=1e-13 * np.abs(ref).max())


# ------------------------------------------------------------------ gmsh 4.1 ingestion
_MSH2D = ["gmsh_beam_0.8", "gmsh_beam_0.5", "gmsh_beam_0.3", "gmsh_beam_0.18", "gmsh_beam_0.1"]
_MSH3D = ["gmsh_beam3d_1", "gmsh_beam3d_0.5"]
_REF_INPUT = "/root/reference/scripts/InputFiles"


@pytest.mark.parametrize("name", _MSH2D + _MSH3D)
def test_read_gmsh41_committed_files(name):
    """tests/golden/msh/*.msh (several node blocks, non-contiguous tags, a line-element block) ->
    exactly the coords / conn of the golden fixture the reference ran on."""
    import finite_elements_b200 as fe
    fx = Fixture(name)
    path = os.path.join(ROOT, "tests", "golden", "msh", name + ".msh")
    coords, conn = fe.mesh.read_gmsh41(path, tetrahedra=name in _MSH3D)
    assert coords.dtype == np.float64 and conn.dtype == np.int32
    assert np.array_equal(coords, fx.coords) and np.array_equal(conn, fx.conn)
    parser = fe.mesh.GmshParser.from_file(path)
    assert parser.tetrahedra == (name in _MSH3D) and len(parser.nodes['all_nodes']) == len(fx.coords)
    mesh = parser.define_tetrahedron_element_mesh() if parser.tetrahedra else parser.define_triangular_element_mesh()
    assert sum(len(g.elements) for g in mesh.elements_groups) == len(fx.conn)


@pytest.mark.skipif(not os.path.isdir(_REF_INPUT), reason="reference tree not present (GPU box)")
@pytest.mark.parametrize("name", _MSH2D + _MSH3D)
def test_read_gmsh41_reference_files(name):
    """The reader on the reference's OWN .msh files (scripts/InputFiles, read in place)."""
    import finite_elements_b200 as fe
    fx = Fixture(name)
    if name in _MSH3D:
        path = os.path.join(_REF_INPUT, "3D", name.replace("gmsh_", "") + ".msh")
    else:
        path = os.path.join(_REF_INPUT, "2D", name.replace("gmsh_beam_", "beam_2d_") + ".msh")
    coords, conn = fe.mesh.read_gmsh41(path, tetrahedra=name in _MSH3D)
    assert np.array_equal(coords, fx.coords) and np.array_equal(conn, fx.conn)
    parser = fe.mesh.GmshParser.from_file(path)
    assert parser.tetrahedra == (name in _MSH3D)
    assert np.array_equal(parser.conn, fx.conn)


def test_read_gmsh41_rejects_other_formats(tmp_path):
    import finite_elements_b200 as fe
    p = tmp_path / "old.msh"
    p.write_text("$MeshFormat\n2.2 0 8\n$EndMeshFormat\n")
    with pytest.raises(ValueError):
        fe.mesh.read_gmsh41(str(p))
