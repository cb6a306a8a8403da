"""Loads tests/golden/*.npz (minted by oracle/make_golden.py from the reference's own code)."""
import glob
import json
import os

import numpy as np
import scipy.sparse as sp

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _all_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "*.npz")))


def _kind_of(name):
    return json.loads(str(np.load(os.path.join(GOLDEN, name + ".npz"))["meta"]))["kind"]


def names():
    """The 2D fixtures (triangles: elasticity and magnetics)."""
    return [n for n in _all_names() if _kind_of(n) != "elasticity3d"]


def names3d():
    """The tetrahedral fixtures."""
    return [n for n in _all_names() if _kind_of(n) == "elasticity3d"]


class Fixture:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        self.name = name
        self.z = z
        self.meta = json.loads(str(z["meta"]))
        self.kind = self.meta["kind"]
        self.plane = self.meta["plane"]
        self.dim = {"elasticity": 2, "elasticity3d": 3}.get(self.kind, 1)
        self.coords, self.conn = z["coords"], z["conn"]
        self.mat_id, self.mat = z["mat_id"], z["mat"]
        self.records = self.meta["records"]
        self.ndof = len(self.coords) * self.dim

    def rec(self, key):
        out = []
        for r in self.records.get(key, []):
            out.append(tuple(r))
        return out

    def csr(self, prefix):
        z = self.z
        return sp.csr_matrix((z[f"ref_{prefix}_data"], z[f"ref_{prefix}_indices"], z[f"ref_{prefix}_indptr"]),
                             shape=tuple(z[f"ref_{prefix}_shape"]))

    def ref(self, key):
        return self.z["ref_" + key]

    @property
    def oracle_kind(self):
        from oracle import numpy_oracle as no
        if self.kind == "magnetic":
            return no.KIND_MAGNETIC
        if self.kind == "elasticity3d":
            return no.KIND_ELAST_TET
        return no.KIND_ELAST_PSTRAIN if self.plane == "strain" else no.KIND_ELAST_PSTRESS


def assert_close_rowscaled(a, b, rtol=1e-12):
    """|a-b| <= rtol * max|row of b| (structurally-present zeros: SURVEY §7 hard part 7)."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape
    scale = np.max(np.abs(b.reshape(b.shape[0], -1)), axis=1)
    scale = np.where(scale == 0, 1.0, scale).reshape((-1,) + (1,) * (b.ndim - 1))
    err = np.max(np.abs(a - b) / scale)
    assert err <= rtol, f"row-scaled error {err:.3e} > {rtol:.1e}"


def assert_csr_values_close(a, b, rtol=1e-12):
    """Same pattern (bit-exact) and values within rtol * max|row|."""
    assert a.shape == b.shape
    assert np.array_equal(a.indptr, b.indptr), "indptr differs"
    assert np.array_equal(a.indices, b.indices), "indices differ"
    rows = np.repeat(np.arange(b.shape[0]), np.diff(b.indptr))
    scale = np.zeros(b.shape[0])
    np.maximum.at(scale, rows, np.abs(b.data))
    scale[scale == 0] = 1.0
    err = np.max(np.abs(a.data - b.data) / scale[rows]) if len(rows) else 0.0
    assert err <= rtol, f"row-scaled value error {err:.3e} > {rtol:.1e}"


def build_object_analysis(fx, with_element_records=False, **kw):
    """The fixture as the reference's scripts would build it: one Python object per node /
    element, load and condition records on those objects (finite_elements_b200 classes)."""
    import finite_elements_b200 as fe
    m = fe.mesh
    if fx.kind == "elasticity3d":
        nodes = [m.Node3D(float(x), float(y), float(z)) for x, y, z in fx.coords]
    else:
        nodes = [m.Node2D(float(x), float(y)) for x, y in fx.coords]
    groups, elems = [], []
    bounds = fx.meta["group_bounds"]
    for g in range(len(bounds) - 1):
        ge = []
        for e in range(bounds[g], bounds[g + 1]):
            p = fx.mat[g]
            if fx.kind == "elasticity3d":
                tet = m.TetrahedralElement([nodes[i] for i in fx.conn[e]])
                ge.append(fe.elements.ElasticityTetrahedralElement3D(tet, p[0], p[1], p[3]))
                continue
            tri = m.TriangularElement2D([nodes[i] for i in fx.conn[e]])
            if fx.kind == "elasticity":
                ge.append(fe.elements.ElasticityTriangularElement2D(tri, p[0], p[1], p[3], p[2]))
            else:
                ge.append(fe.elements.MagneticElement2D(tri, p[0]))
        elems.extend(ge)
        groups.append(m.ElementsGroup(ge, ''))
    mesh = m.Mesh(groups)
    mesh.nodes = nodes  # keep the fixture numbering, as beam2d_example_3.py:72-73 does
    mesh.node_to_index = {nodes[i]: i for i in range(len(nodes))}

    class Edge:
        def __init__(self, a, b):
            self.start, self.end = a, b

    nl = [fe.loads.NodeLoad(nodes[n], v, d) for n, v, d in fx.rec("node_loads")]
    edl = [fe.loads.EdgeLoad(Edge(nodes[a], nodes[b]), v, d) for a, b, v, d in fx.rec("edge_loads")]
    nb = [fe.conditions.NodeBoundaryCondition(nodes[n], v, d) for n, v, d in fx.rec("node_bcs")]
    edb = [fe.conditions.EdgeBoundaryCondition(Edge(nodes[a], nodes[b]), v, d) for a, b, v, d in fx.rec("edge_bcs")]
    el, elb = [], []
    if with_element_records:
        el = [fe.loads.ElementsLoad([elems[j] for j in idx], v, d) for idx, v, d in fx.rec("elements_loads")]
        elb = [fe.conditions.ElementBoundaryCondition(elems[j], v, d) for j, v, d in fx.rec("element_bcs")]
    ml = [fe.loads.MagnetLoad([elems[j] for j in idx], [nodes[i] for i in ncn], m.Vector2D(mx, my))
          for idx, ncn, mx, my in fx.records.get("magnet_loads", [])]
    ps = fx.plane
    an = fe.analysis.FiniteElementAnalysis(mesh, el, edl, nl, ml, [], nb, edb, elb,
                                           None if ps is None else ps == "strain",
                                           None if ps is None else ps == "stress", **kw)
    return an, mesh, elems
