"""TEST INFRASTRUCTURE -- not product code.

Loads the reference's *unmodified* modules from /root/reference/finite_elements
behind stand-ins for the three third-party packages that are absent from this
image (volmdlr, dessia_common, matplotlib).  Used ONLY by oracle/make_golden.py
and tests/test_oracle.py's literal-reference checks in the build container, where /root/reference
exists; nothing here runs on the GPU box and nothing in the product imports it.

What is the reference's own code and what is restated
------------------------------------------------------
* finite_elements/{core,elements,loads,conditions,analysis,results}.py are
  imported as they lie (sys.modules['finite_elements'] is pre-seeded so that
  finite_elements/__init__.py:5-7, which needs an installed dist, is skipped).
* volmdlr geometry is RESTATED here (volmdlr>=0.10.0 is a setup.py:112
  dependency that is not vendored): Point2D arithmetic, TriangularElement2D
  .area (=|u x v|/2) and .form_functions (three 3x3 solves of rows [1, x, y]),
  Mesh.nodes in first-seen order.  The reference duplicates these formulas at
  elements.py:33 (det), :100-116 (use of b_i, c_i), :406-408 (detJ).  No
  reference test pins values at that seam => node numbering is compared by
  coordinates / explicit index arrays, never by volmdlr's implicit ordering.
"""
import importlib
import math
import os
import sys
import types

import numpy as np

REFERENCE_ROOT = os.environ.get("FE_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "finite_elements"))


# --------------------------------------------------------------------------
# volmdlr stand-in
# --------------------------------------------------------------------------
class Vector2D:
    def __init__(self, x, y=None, name=''):
        if y is None:  # analysis.py:570 passes a 2-list
            x, y = x[0], x[1]
        self.x = x
        self.y = y
        self.name = name

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __iter__(self):
        return iter((self.x, self.y))

    def __len__(self):
        return 2

    def __add__(self, o):
        return self.__class__(self.x + o[0], self.y + o[1])

    def __sub__(self, o):
        return self.__class__(self.x - o[0], self.y - o[1])

    def __mul__(self, s):
        return self.__class__(self.x * s, self.y * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return self.__class__(self.x / s, self.y / s)

    def __neg__(self):
        return self.__class__(-self.x, -self.y)

    def dot(self, o):
        return self.x * o[0] + self.y * o[1]

    Dot = dot

    def cross(self, o):
        return self.x * o[1] - self.y * o[0]

    def norm(self):
        return math.hypot(self.x, self.y)

    def normalize(self):
        n = self.norm()
        self.x /= n
        self.y /= n

    def _key(self):
        return (int(round(self.x * 1e6)), int(round(self.y * 1e6)))

    def __eq__(self, o):
        return isinstance(o, Vector2D) and self._key() == o._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"{self.__class__.__name__}({self.x}, {self.y})"


class Point2D(Vector2D):
    pass


class Node2D(Point2D):
    pass


class LinearElement:
    def __init__(self, points, interior_normal, name=''):
        self.points = points
        self.interior_normal = interior_normal

    def length(self):
        return (self.points[1] - self.points[0]).norm()

    def _key(self):
        return frozenset(p._key() for p in self.points)

    def __eq__(self, o):
        return self._key() == o._key()

    def __hash__(self):
        return hash(self._key())


class TriangularElement:
    pass


class TriangularElement2D(TriangularElement):
    def __init__(self, points, name=''):
        self.points = points
        self.name = name
        u = points[1] - points[0]
        v = points[2] - points[0]
        self.area = 0.5 * abs(u.cross(v))
        self.center = (points[0] + points[1] + points[2]) / 3
        self.form_functions = self._form_functions()
        self.linear_elements = self._linear_elements()

    def _form_functions(self):
        a = np.array([[1.0, p[0], p[1]] for p in self.points])
        return tuple(list(np.linalg.solve(a, e)) for e in np.eye(3))

    def _linear_elements(self):
        out = []
        for i in range(3):
            p, q, r = self.points[i], self.points[(i + 1) % 3], self.points[(i + 2) % 3]
            t = q - p
            n = Vector2D(-t.y, t.x)
            if n.dot(r - p) < 0:
                n = -n
            nn = n.norm()
            out.append(LinearElement([p, q], Vector2D(n.x / nn, n.y / nn)))
        return out

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        return self is o


class Vector3D:
    def __init__(self, x, y, z, name=''):
        self.x, self.y, self.z = x, y, z
        self.name = name

    def __getitem__(self, i):
        return (self.x, self.y, self.z)[i]

    def __iter__(self):
        return iter((self.x, self.y, self.z))

    def __len__(self):
        return 3

    def __add__(self, o):
        return self.__class__(self.x + o[0], self.y + o[1], self.z + o[2])

    def __sub__(self, o):
        return self.__class__(self.x - o[0], self.y - o[1], self.z - o[2])

    def __mul__(self, s):
        return self.__class__(self.x * s, self.y * s, self.z * s)

    __rmul__ = __mul__

    def __truediv__(self, s):
        return self.__class__(self.x / s, self.y / s, self.z / s)

    def _key(self):
        return (int(round(self.x * 1e6)), int(round(self.y * 1e6)), int(round(self.z * 1e6)))

    def __eq__(self, o):
        return isinstance(o, Vector3D) and self._key() == o._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"{self.__class__.__name__}({self.x}, {self.y}, {self.z})"


class Point3D(Vector3D):
    pass


class Node3D(Point3D):
    pass


class TetrahedralElement:
    """volmdlr.mesh.TetrahedralElement RESTATED (third-party, not vendored).  The reference
    consumes `.points`, `.volume` and `.form_functions` (elements.py:726-749, :823, :854): with
    A = rows [1, x_j, y_j, z_j], volume = |det A| / 6 and form_functions[i] = (alpha_i, a_i, b_i, c_i)
    such that N_i = (alpha_i + a_i x + b_i y + c_i z) / (6 volume) -- the only reading under which
    the reference's B = 1/(6 V) [a_i ...] (:749) is the strain-displacement matrix of a linear
    tetrahedron.  Ke = V B^T D B does not depend on the sign convention of the cofactors."""

    def __init__(self, points, name=''):
        self.points = points
        self.name = name
        a = np.array([[1.0, p[0], p[1], p[2]] for p in points])
        det = np.linalg.det(a)
        self.volume = abs(det) / 6.0
        inv = np.linalg.inv(a)
        self.form_functions = tuple(tuple(abs(det) * inv[:, i]) for i in range(4))
        self.center = (points[0] + points[1] + points[2] + points[3]) / 4

    def __hash__(self):
        return id(self)

    def __eq__(self, o):
        return self is o


class ElementsGroup:
    def __init__(self, elements, name=''):
        self.elements = elements
        self.name = name


class Mesh:
    def __init__(self, elements_groups):
        self.elements_groups = elements_groups
        self.nodes = []
        self.node_to_index = {}
        for g in elements_groups:
            for e in g.elements:
                for p in e.points:
                    if p not in self.node_to_index:
                        self.node_to_index[p] = len(self.nodes)
                        self.nodes.append(p)


class _Anything:
    """Import-time placeholder for matplotlib names the reference binds."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()


class DessiaObject:
    def __init__(self, name='', **kwargs):
        self.name = name


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


_loaded = {}


def load():
    """Return a namespace with the reference's modules (.analysis, .elements, ...)
    and the geometry stand-ins (.vm, .vmmesh)."""
    if _loaded:
        return _loaded["ns"]
    if not available():
        raise RuntimeError("reference tree not present at " + REFERENCE_ROOT)
    vm = _module("volmdlr", Point2D=Point2D, Vector2D=Vector2D, Point3D=Point3D, Vector3D=Vector3D)
    vmmesh = _module("volmdlr.mesh", Node2D=Node2D, Node3D=Node3D, TriangularElement=TriangularElement,
                     TriangularElement2D=TriangularElement2D,
                     TetrahedralElement=TetrahedralElement,
                     ElementsGroup=ElementsGroup, Mesh=Mesh, LinearElement=LinearElement)
    vmcore = _module("volmdlr.core", EdgeStyle=_Anything)
    vm.mesh, vm.core = vmmesh, vmcore
    dc = _module("dessia_common")
    dc.core = _module("dessia_common.core", DessiaObject=DessiaObject)
    if "matplotlib" not in sys.modules:
        mpl = _module("matplotlib")
        mpl.pyplot = _module("matplotlib.pyplot", subplots=_Anything(), gcf=_Anything(),
                             cm=_Anything())
        mpl.colors = _module("matplotlib.colors", LinearSegmentedColormap=_Anything,
                             Normalize=_Anything)
        mpl.tri = _module("matplotlib.tri", Triangulation=_Anything, TriAnalyzer=_Anything,
                          UniformTriRefiner=_Anything)
    pkg = types.ModuleType("finite_elements")
    pkg.__path__ = [os.path.join(REFERENCE_ROOT, "finite_elements")]
    sys.modules["finite_elements"] = pkg
    ns = types.SimpleNamespace(vm=vm, vmmesh=vmmesh)
    for sub in ("core", "elements", "loads", "conditions", "results", "analysis"):
        mod = importlib.import_module("finite_elements." + sub)
        setattr(pkg, sub, mod)
        setattr(ns, sub, mod)
    _loaded["ns"] = ns
    return ns


# --------------------------------------------------------------------------
# Convenience: build a reference problem from flat arrays (fixture description)
# --------------------------------------------------------------------------
def build_reference_analysis(ns, coords, conn, groups, kind, node_loads=(), node_bcs=(),
                             elements_loads=(), edge_loads=(), edge_bcs=(), element_bcs=(),
                             plane_strain=None, plane_stress=None, magnet_loads=()):
    """coords (N,2), conn (E,3) int, groups = list of dict(start, stop, params).
    kind 'elasticity': params = (E, nu, rho, t); 'magnetic': params = (mu,).
    node_loads / node_bcs: iterables of (node_index, value, dimension(1-based)).
    elements_loads: iterables of (list_of_element_indices, value, dimension).
    edge_loads / edge_bcs: (node_index_start, node_index_end, value, dimension).
    element_bcs: (element_index, value, dimension).
    magnet_loads: (list_of_element_indices, list_of_non_contour_node_indices, m_x, m_y)
    (loads.py:105-147; the contour edges come from the restated volmdlr `linear_elements`).
    mesh.nodes is forced to the given numbering (as beam2d_example_3.py:72-73 does)."""
    if kind == "elasticity3d":
        nodes = [ns.vmmesh.Node3D(float(x), float(y), float(z)) for x, y, z in coords]
    else:
        nodes = [ns.vmmesh.Node2D(float(x), float(y)) for x, y in coords]
    all_elems, egroups = [], []
    for g in groups:
        elems = []
        for e in range(g["start"], g["stop"]):
            if kind == "elasticity3d":
                tet = ns.vmmesh.TetrahedralElement([nodes[i] for i in conn[e]])
                em, nu, rho, _t = g["params"]
                elems.append(ns.elements.ElasticityTetrahedralElement3D(tet, em, nu, rho))
                continue
            tri = ns.vmmesh.TriangularElement2D([nodes[i] for i in conn[e]])
            if kind == "elasticity":
                em, nu, rho, t = g["params"]
                elems.append(ns.elements.ElasticityTriangularElement2D(tri, em, nu, rho, t))
            else:
                elems.append(ns.elements.MagneticElement2D(tri, g["params"][0]))
        all_elems.extend(elems)
        egroups.append(ns.vmmesh.ElementsGroup(elems, g.get("name", "")))
    mesh = ns.vmmesh.Mesh(egroups)
    mesh.nodes = nodes
    mesh.node_to_index = {nodes[i]: i for i in range(len(nodes))}

    class _Edge:
        def __init__(self, a, b):
            self.start, self.end = a, b

    nl = [ns.loads.NodeLoad(nodes[i], v, d) for i, v, d in node_loads]
    nb = [ns.conditions.NodeBoundaryCondition(nodes[i], v, d) for i, v, d in node_bcs]
    el = [ns.loads.ElementsLoad([all_elems[j] for j in idx], v, d) for idx, v, d in elements_loads]
    edl = [ns.loads.EdgeLoad(_Edge(nodes[a], nodes[b]), v, d) for a, b, v, d in edge_loads]
    edb = [ns.conditions.EdgeBoundaryCondition(_Edge(nodes[a], nodes[b]), v, d)
           for a, b, v, d in edge_bcs]
    elb = [ns.conditions.ElementBoundaryCondition(all_elems[j], v, d) for j, v, d in element_bcs]
    ml = [ns.loads.MagnetLoad([all_elems[j] for j in idx], [nodes[i] for i in ncn], ns.vm.Vector2D(mx, my))
          for idx, ncn, mx, my in magnet_loads]
    an = ns.analysis.FiniteElementAnalysis(mesh, el, edl, nl, ml, [], nb, edb, elb,
                                           plane_strain, plane_stress)
    return an, mesh, all_elems
