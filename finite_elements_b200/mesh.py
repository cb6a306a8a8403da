"""Mesh containers for the host layer.

The reference takes `volmdlr.mesh` objects (third-party, not vendored).
`FiniteElementAnalysis` here is duck-typed on the attributes the reference reads
(`mesh.nodes`, `mesh.node_to_index`, `mesh.elements_groups[*].elements[*].points`), so a real
volmdlr mesh works unchanged.  This module provides
  * look-alikes of the few volmdlr classes the reference's scripts use, for environments
    without volmdlr (Point2D / Node2D / TriangularElement2D / ElementsGroup / Mesh, and
    Point3D / Node3D / TetrahedralElement for the tetrahedral path);
  * ArrayMesh: the same surface backed by flat arrays, so that million-element meshes
    never become Python objects;
  * structured_mesh(): the synthetic triangulations of SURVEY §8d;
  * read_gmsh41(): gmsh 4.1 ASCII reader (scripts/InputFiles/2D/*.msh) -> arrays.
"""
import math

import numpy as np


class Vector2D:
    """x, y with the arithmetic the reference calls (analysis.py:570-573, results.py:94-101)."""

    __slots__ = ("x", "y", "name")

    def __init__(self, x, y=None, name=''):
        if y is None:
            x, y = x[0], x[1]
        self.x, self.y, self.name = x, y, name

    def __getitem__(self, i):
        return (self.x, self.y)[i]

    def __iter__(self):
        yield self.x
        yield self.y

    def __len__(self):
        return 2

    def __add__(self, other):
        return type(self)(self.x + other[0], self.y + other[1])

    def __sub__(self, other):
        return type(self)(self.x - other[0], self.y - other[1])

    def __mul__(self, k):
        return type(self)(self.x * k, self.y * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return type(self)(self.x / k, self.y / k)

    def __neg__(self):
        return type(self)(-self.x, -self.y)

    def dot(self, other):
        return self.x * other[0] + self.y * other[1]

    Dot = dot

    def cross(self, other):
        return self.x * other[1] - self.y * other[0]

    def norm(self):
        return math.hypot(self.x, self.y)

    def normalize(self):
        n = self.norm()
        self.x, self.y = self.x / n, self.y / n

    def _key(self):  # volmdlr compares points with a ~1e-6 tolerance
        return (round(self.x * 1e6), round(self.y * 1e6))

    def __eq__(self, other):
        return isinstance(other, Vector2D) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"{type(self).__name__}({self.x}, {self.y})"


class Point2D(Vector2D):
    __slots__ = ()


class Node2D(Point2D):
    __slots__ = ()


class Vector3D:
    """x, y, z (volmdlr.Vector3D look-alike; approximate equality like the 2D points)."""

    __slots__ = ("x", "y", "z", "name")

    def __init__(self, x, y=None, z=None, name=''):
        if y is None:
            x, y, z = x[0], x[1], x[2]
        self.x, self.y, self.z, self.name = x, y, z, name

    def __getitem__(self, i):
        return (self.x, self.y, self.z)[i]

    def __iter__(self):
        yield self.x
        yield self.y
        yield self.z

    def __len__(self):
        return 3

    def __add__(self, other):
        return type(self)(self.x + other[0], self.y + other[1], self.z + other[2])

    def __sub__(self, other):
        return type(self)(self.x - other[0], self.y - other[1], self.z - other[2])

    def __mul__(self, k):
        return type(self)(self.x * k, self.y * k, self.z * k)

    __rmul__ = __mul__

    def __truediv__(self, k):
        return type(self)(self.x / k, self.y / k, self.z / k)

    def dot(self, other):
        return self.x * other[0] + self.y * other[1] + self.z * other[2]

    def norm(self):
        return math.sqrt(self.x * self.x + self.y * self.y + self.z * self.z)

    def _key(self):
        return (round(self.x * 1e6), round(self.y * 1e6), round(self.z * 1e6))

    def __eq__(self, other):
        return isinstance(other, Vector3D) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())

    def __repr__(self):
        return f"{type(self).__name__}({self.x}, {self.y}, {self.z})"


class Point3D(Vector3D):
    __slots__ = ()


class Node3D(Point3D):
    __slots__ = ()


class LinearElement:
    def __init__(self, points, interior_normal, name=''):
        self.points = points
        self.interior_normal = interior_normal
        self.name = name

    def length(self):
        return (self.points[1] - self.points[0]).norm()

    def _key(self):
        return frozenset(p._key() for p in self.points)

    def __eq__(self, other):
        return isinstance(other, LinearElement) and self._key() == other._key()

    def __hash__(self):
        return hash(self._key())


class TriangularElement:
    pass


class TriangularElement2D(TriangularElement):
    """points, area, center, form_functions, linear_elements (what the reference reads)."""

    def __init__(self, points, name=''):
        self.points = list(points)
        self.name = name
        (x1, y1), (x2, y2), (x3, y3) = ((p[0], p[1]) for p in self.points)
        self._cross = (x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)
        self.area = 0.5 * abs(self._cross)

    @property
    def center(self):
        p = self.points
        return Point2D((p[0][0] + p[1][0] + p[2][0]) / 3, (p[0][1] + p[1][1] + p[2][1]) / 3)

    @property
    def form_functions(self):
        """((a_i, b_i, c_i))_i with N_i = a_i + b_i x + c_i y, N_i(p_j) = delta_ij."""
        (x1, y1), (x2, y2), (x3, y3) = ((p[0], p[1]) for p in self.points)
        d = self._cross
        return ([(x2 * y3 - x3 * y2) / d, (y2 - y3) / d, (x3 - x2) / d],
                [(x3 * y1 - x1 * y3) / d, (y3 - y1) / d, (x1 - x3) / d],
                [(x1 * y2 - x2 * y1) / d, (y1 - y2) / d, (x2 - x1) / d])

    @property
    def linear_elements(self):
        out = []
        for i in range(3):
            p, q, r = self.points[i], self.points[(i + 1) % 3], self.points[(i + 2) % 3]
            t = Vector2D(q[0] - p[0], q[1] - p[1])
            n = Vector2D(-t.y, t.x)
            if n.dot(Vector2D(r[0] - p[0], r[1] - p[1])) < 0:
                n = -n
            out.append(LinearElement([p, q], n / n.norm()))
        return out


class TetrahedralElement:
    """points, volume, center, form_functions as ElasticityTetrahedralElement3D reads them
    (elements.py:726-749, :823, :854): volume = |det [1 x y z]| / 6 and form_functions[i] =
    (alpha_i, a_i, b_i, c_i) with N_i = (alpha_i + a_i x + b_i y + c_i z) / (6 volume)."""

    def __init__(self, points, name=''):
        self.points = list(points)
        self.name = name
        a = np.array([[1.0, p[0], p[1], p[2]] for p in self.points])
        self._det = float(np.linalg.det(a))
        self.volume = abs(self._det) / 6.0
        self._a = a

    @property
    def center(self):
        p = self.points
        return Point3D(*(sum(q[d] for q in p) / 4 for d in range(3)))

    @property
    def form_functions(self):
        inv = np.linalg.inv(self._a)
        return tuple(tuple(abs(self._det) * inv[:, i]) for i in range(4))


class ElementsGroup:
    def __init__(self, elements, name=''):
        self.elements = elements
        self.name = name


class Mesh:
    """nodes in first-seen order; node_to_index maps a point (approximate equality) to it."""

    def __init__(self, elements_groups):
        self.elements_groups = elements_groups
        self.nodes = []
        self.node_to_index = {}
        for group in elements_groups:
            for element in group.elements:
                for point in element.points:
                    if point not in self.node_to_index:
                        self.node_to_index[point] = len(self.nodes)
                        self.nodes.append(point)


# ---------------------------------------------------------------------------------------
# array-backed mesh
# ---------------------------------------------------------------------------------------
class _LazyNodes:
    def __init__(self, coords):
        self._c = coords

    def __len__(self):
        return len(self._c)

    def _node(self, row):
        return Node3D(*(float(v) for v in row)) if len(row) == 3 else Node2D(float(row[0]), float(row[1]))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self._node(row) for row in self._c[i]]
        return self._node(self._c[i])

    def __iter__(self):
        for row in self._c:
            yield self._node(row)


class _NodeIndex:
    """node_to_index for ArrayMesh: accepts an int node id or a point (coordinate look-up)."""

    def __init__(self, coords):
        self._c = coords
        self._map = None

    def __getitem__(self, node):
        if isinstance(node, (int, np.integer)):
            return int(node)
        if self._map is None:
            keys = np.round(self._c * 1e6).astype(np.int64)
            self._map = {tuple(int(v) for v in row): i for i, row in enumerate(keys)}
        return self._map[tuple(round(node[d] * 1e6) for d in range(self._c.shape[1]))]


class ArrayElement:
    """Handle on element `index` of an ArrayMesh: what load / condition records need
    (.area for ElementsLoad, .points as node indices, int() for the flat index)."""

    __slots__ = ("index", "points", "area")

    def __init__(self, index, points, area):
        self.index, self.points, self.area = index, points, area

    def __int__(self):
        return self.index

    __index__ = __int__


class ArrayMesh:
    """Flat-array mesh with the attribute surface of volmdlr.mesh.Mesh.

    coords f64[N,2]; conn i32[E,3]; group_bounds [0, e1, ..., E] (contiguous element groups);
    (kind 'elasticity3d': coords f64[N,3], conn i32[E,4] -- linear tetrahedra);
    per-element objects are built on demand (only small meshes ever need them); `kind` is
    'elasticity', 'magnetic' or 'elasticity3d' and
    `group_params` holds, per group, the flat material row the device uses
    ((E, nu, thickness, rho) or (mu, 0, 0, 0))."""

    def __init__(self, coords, conn, kind, group_params, group_bounds=None, group_names=None):
        self.coords = np.ascontiguousarray(coords, dtype=np.float64)
        self.conn = np.ascontiguousarray(conn, dtype=np.int32)
        self.kind = kind
        self.group_bounds = [0, len(self.conn)] if group_bounds is None else [int(b) for b in group_bounds]
        self.group_params = np.atleast_2d(np.asarray(group_params, dtype=np.float64))
        if self.group_params.shape != (len(self.group_bounds) - 1, 4):
            raise ValueError("group_params must have one (p0, p1, p2, p3) row per group")
        self.group_names = group_names or [''] * (len(self.group_bounds) - 1)
        self.nodes = _LazyNodes(self.coords)
        self.node_to_index = _NodeIndex(self.coords)
        self._groups = None

    @property
    def dimension(self):
        return {'elasticity': 2, 'elasticity3d': 3}.get(self.kind, 1)

    def element(self, index):
        """ArrayElement handle (for ElementsLoad / ElementBoundaryCondition records)."""
        n = [int(v) for v in self.conn[index]]
        if self.kind == 'elasticity3d':
            p = self.coords[n]
            return ArrayElement(int(index), n, abs(float(np.linalg.det(p[1:] - p[0]))) / 6.0)
        (x1, y1), (x2, y2), (x3, y3) = self.coords[n]
        return ArrayElement(int(index), n, 0.5 * abs((x2 - x1) * (y3 - y1) - (y2 - y1) * (x3 - x1)))

    @property
    def mat_id(self):
        out = np.empty(len(self.conn), dtype=np.int32)
        for g in range(len(self.group_bounds) - 1):
            out[self.group_bounds[g]:self.group_bounds[g + 1]] = g
        return out

    @property
    def elements_groups(self):
        if self._groups is None:
            from . import elements as fe_elements
            groups = []
            for g in range(len(self.group_bounds) - 1):
                elems = []
                p = self.group_params[g]
                for e in range(self.group_bounds[g], self.group_bounds[g + 1]):
                    if self.kind == 'elasticity3d':
                        tet = TetrahedralElement([self.nodes[int(i)] for i in self.conn[e]])
                        elems.append(fe_elements.ElasticityTetrahedralElement3D(tet, p[0], p[1], p[3]))
                        continue
                    tri = TriangularElement2D([self.nodes[int(i)] for i in self.conn[e]])
                    if self.kind == 'elasticity':
                        elems.append(fe_elements.ElasticityTriangularElement2D(tri, p[0], p[1], p[3], p[2]))
                    else:
                        elems.append(fe_elements.MagneticElement2D(tri, p[0]))
                groups.append(ElementsGroup(elems, self.group_names[g]))
            self._groups = groups
        return self._groups


def flatten_mesh(mesh):
    """Mesh object -> flat arrays, once: coords f64[N,2], conn i32[E,3], mat_id i32[E], mat f64[G,4]
    ((E, nu, thickness, rho) or (mu, 0, 0, 0) rows, de-duplicated), `elements` (the element objects
    in flat order, None for an ArrayMesh), `element_index` ({id(element): flat index}) and
    `magnetic`.  Duck-typed on what the reference reads from volmdlr meshes."""
    if isinstance(mesh, ArrayMesh):
        return dict(coords=mesh.coords, conn=mesh.conn, mat_id=mesh.mat_id, mat=mesh.group_params,
                    elements=None, element_index=None, magnetic=mesh.kind == 'magnetic',
                    space_dim=mesh.coords.shape[1])
    first = mesh.elements_groups[0].elements[0] if mesh.elements_groups and mesh.elements_groups[0].elements else None
    sdim = 3 if first is not None and len(first.points) == 4 else 2
    coords = np.array([[node[d] for d in range(sdim)] for node in mesh.nodes], dtype=np.float64).reshape(-1, sdim)
    conn, mat_id, rows, row_of, element_index, elements = [], [], [], {}, {}, []
    magnetic = None
    for group in mesh.elements_groups:
        for element in group.elements:
            is_mag = hasattr(element, 'mu_total')
            if magnetic is None:
                magnetic = is_mag
            elif magnetic != is_mag:
                raise NotImplementedError('a mesh mixing magnetic and elasticity elements is not supported')
            if is_mag:
                row = (float(element.mu_total), 0.0, 0.0, 0.0)
            else:
                row = (float(element.elasticity_modulus), float(element.poisson_ratio),
                       float(getattr(element, 'thickness', 1.0)), float(element.mass_density))
            if row not in row_of:
                row_of[row] = len(rows)
                rows.append(row)
            element_index[id(element)] = len(conn)
            elements.append(element)
            conn.append([mesh.node_to_index[point] for point in element.points])
            mat_id.append(row_of[row])
    return dict(coords=coords, conn=np.array(conn, dtype=np.int32).reshape(-1, sdim + 1),
                mat_id=np.array(mat_id, dtype=np.int32), mat=np.array(rows, dtype=np.float64).reshape(-1, 4),
                elements=elements, element_index=element_index, magnetic=bool(magnetic), space_dim=sdim)


# ---------------------------------------------------------------------------------------
# synthetic meshes and gmsh input
# ---------------------------------------------------------------------------------------
def structured_mesh(nx, ny, h=None, jitter=0.0, seed=0):
    """Structured triangulation of [0, nx h] x [0, ny h], h = 1/ny: nodes
    row-major id = j (nx+1) + i; cell (i, j) -> T0 = [(i,j),(i+1,j),(i,j+1)],
    T1 = [(i+1,j+1),(i+1,j),(i,j+1)] (the orientation of beam2d_example_2.py:39-40);
    element id 2 (j nx + i) + {0, 1}.  Returns coords f64[N,2], conn i32[E,3]."""
    h = 1.0 / ny if h is None else h
    i = np.arange(nx + 1, dtype=np.float64)
    j = np.arange(ny + 1, dtype=np.float64)
    coords = np.empty(((ny + 1) * (nx + 1), 2))
    coords[:, 0] = np.tile(i * h, ny + 1)
    coords[:, 1] = np.repeat(j * h, nx + 1)
    if jitter:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter * h, jitter * h, size=coords.shape)
        gi = np.tile(np.arange(nx + 1), ny + 1)
        gj = np.repeat(np.arange(ny + 1), nx + 1)
        inner = (gi > 0) & (gi < nx) & (gj > 0) & (gj < ny)
        coords[inner] += d[inner]
    cell = (np.arange(ny)[:, None] * (nx + 1) + np.arange(nx)[None, :]).reshape(-1)
    conn = np.empty((2 * nx * ny, 3), dtype=np.int32)
    conn[0::2, 0], conn[0::2, 1], conn[0::2, 2] = cell, cell + 1, cell + nx + 1
    conn[1::2, 0], conn[1::2, 1], conn[1::2, 2] = cell + nx + 2, cell + 1, cell + nx + 1
    return coords, conn


def structured_mesh_torch(nx, ny, device, h=None, row_lo=0, row_hi=None):
    """Same mesh generated directly on the device (no host arrays at 16 M triangles).
    row_lo/row_hi select the cell rows [row_lo, row_hi) and the node rows [row_lo, row_hi]
    (global numbering is kept; used by the partitioner in dist.py)."""
    import torch
    h = 1.0 / ny if h is None else h
    row_hi = ny if row_hi is None else row_hi
    i = torch.arange(nx + 1, device=device, dtype=torch.float64) * h
    j = torch.arange(ny + 1, device=device, dtype=torch.float64) * h
    coords = torch.stack([i.repeat(ny + 1), j.repeat_interleave(nx + 1)], dim=1).contiguous()
    cj = torch.arange(row_lo, row_hi, device=device, dtype=torch.int64)
    ci = torch.arange(nx, device=device, dtype=torch.int64)
    cell = (cj[:, None] * (nx + 1) + ci[None, :]).reshape(-1)
    t0 = torch.stack([cell, cell + 1, cell + nx + 1], dim=1)
    t1 = torch.stack([cell + nx + 2, cell + 1, cell + nx + 1], dim=1)
    conn = torch.stack([t0, t1], dim=1).reshape(-1, 3).to(torch.int32).contiguous()
    return coords, conn


def read_gmsh41(path, tetrahedra=False):
    """gmsh 4.1 ASCII: nodes in file order (gmsh.nodes['all_nodes'] order in
    beam2d_example_3.py:72), 3-node triangles (element type 2) -> coords (N,2), conn (E,3); with
    tetrahedra=True the 4-node tetrahedra (type 4, beam3d_example_2.py:59-61) -> coords (N,3), conn (E,4)."""
    with open(path) as fh:
        lines = [ln.strip() for ln in fh]
    if "$MeshFormat" not in lines or not lines[lines.index("$MeshFormat") + 1].startswith("4.1"):
        raise ValueError(f"{path}: only gmsh 4.1 ASCII is supported")
    k = lines.index("$Nodes") + 1
    n_blocks, n_nodes = (int(t) for t in lines[k].split()[:2])
    k += 1
    tags, xy = [], []
    for _ in range(n_blocks):
        count = int(lines[k].split()[3])
        k += 1
        tags.extend(int(lines[k + r]) for r in range(count))
        k += count
        for r in range(count):
            xy.append(tuple(float(t) for t in lines[k + r].split()[:(3 if tetrahedra else 2)]))
        k += count
    if len(tags) != n_nodes:
        raise ValueError(f"{path}: node count mismatch")
    index_of = {t: i for i, t in enumerate(tags)}
    k = lines.index("$Elements") + 1
    n_blocks = int(lines[k].split()[0])
    k += 1
    tris = []
    for _ in range(n_blocks):
        _, _, etype, count = (int(t) for t in lines[k].split())
        k += 1
        if etype == (4 if tetrahedra else 2):
            for r in range(count):
                t = lines[k + r].split()
                tris.append([index_of[int(v)] for v in t[1:(5 if tetrahedra else 4)]])
        k += count
    return np.array(xy, dtype=np.float64), np.array(tris, dtype=np.int32).reshape(-1, 4 if tetrahedra else 3)


def write_gmsh41(path, coords, conn, node_blocks=1, boundary_edges=(), tag_of=lambda i: i + 1):
    """Writes coords (N,2|3) / conn (E,3|4) as gmsh 4.1 ASCII: `node_blocks` entity blocks of nodes
    (node i of the file order carries tag `tag_of(i)`; gmsh tags need not be contiguous), an optional block of 2-node line elements (type 1, skipped by every
    reader of surface / volume meshes) and one block of triangles (type 2) or tetrahedra (type 4)."""
    coords = np.asarray(coords, dtype=np.float64)
    conn = np.asarray(conn)
    n, sdim = coords.shape
    tets = conn.shape[1] == 4
    xyz = np.zeros((n, 3))
    xyz[:, :sdim] = coords
    bounds = [n * b // node_blocks for b in range(node_blocks + 1)]
    out = ["$MeshFormat", "4.1 0 8", "$EndMeshFormat", "$Nodes", f"{node_blocks} {n} {tag_of(0)} {tag_of(n - 1)}"]
    for b in range(node_blocks):
        lo, hi = bounds[b], bounds[b + 1]
        out.append(f"{2 if b else 0} {b + 1} 0 {hi - lo}")
        out.extend(str(tag_of(t)) for t in range(lo, hi))
        out.extend(" ".join(repr(float(v)) for v in xyz[t]) for t in range(lo, hi))
    out.append("$EndNodes")
    n_lines = len(boundary_edges)
    n_el = n_lines + len(conn)
    out += ["$Elements", f"{2 if n_lines else 1} {n_el} 1 {n_el}"]
    tag = 1
    if n_lines:
        out.append(f"1 1 1 {n_lines}")
        for a, b in boundary_edges:
            out.append(f"{tag} {tag_of(a)} {tag_of(b)}")
            tag += 1
    out.append(f"{3 if tets else 2} 1 {4 if tets else 2} {len(conn)}")
    for row in conn:
        out.append(f"{tag} " + " ".join(str(tag_of(int(v))) for v in row))
        tag += 1
    out.append("$EndElements")
    with open(path, "w") as fh:
        fh.write("\n".join(out) + "\n")


class GmshParser:
    """The part of volmdlr.gmsh_vm.GmshParser the reference's scripts use
    (scripts/Elasticity/beam2d_example_3.py:52-73, beam3d_example_2.py:52-66): `from_file`, the
    `nodes['all_nodes']` list in file order and the two `define_*_element_mesh` builders."""

    def __init__(self, coords, conn, tetrahedra):
        self.coords, self.conn, self.tetrahedra = coords, conn, tetrahedra
        cls = Node3D if tetrahedra else Node2D
        self.nodes = {'all_nodes': [cls(*(float(v) for v in row)) for row in coords]}

    @classmethod
    def from_file(cls, file_path):
        """A file holding tetrahedra is a 3D mesh (Node3D, its surface triangles are ignored);
        otherwise the triangles define a 2D mesh (Node2D)."""
        coords, conn = read_gmsh41(file_path, tetrahedra=True)
        if len(conn):
            return cls(coords, conn, True)
        coords, conn = read_gmsh41(file_path, tetrahedra=False)
        return cls(coords, conn, False)

    def define_triangular_element_mesh(self):
        nodes = self.nodes['all_nodes']
        return Mesh([ElementsGroup([TriangularElement2D([nodes[int(i)] for i in row]) for row in self.conn], '')])

    def define_tetrahedron_element_mesh(self):
        nodes = self.nodes['all_nodes']
        return Mesh([ElementsGroup([TetrahedralElement([nodes[int(i)] for i in row]) for row in self.conn], '')])


def structured_tet_mesh(nx, ny, nz, h=1.0, jitter=0.0, seed=0):
    """Box of nx x ny x nz cells of size h, nodes id = (k (ny+1) + j) (nx+1) + i, every cell cut into
    the 6 tetrahedra of the Kuhn triangulation (conforming).  Returns coords f64[N,3], conn i32[E,4]."""
    ii, jj, kk = np.meshgrid(np.arange(nx + 1), np.arange(ny + 1), np.arange(nz + 1), indexing="ij")
    nid = (kk * (ny + 1) + jj) * (nx + 1) + ii
    coords = np.zeros(((nx + 1) * (ny + 1) * (nz + 1), 3))
    coords[nid.reshape(-1)] = np.stack([ii, jj, kk], axis=-1).reshape(-1, 3) * h
    if jitter:
        rng = np.random.default_rng(seed)
        d = rng.uniform(-jitter * h, jitter * h, size=coords.shape)
        inner = nid[(ii > 0) & (ii < nx) & (jj > 0) & (jj < ny) & (kk > 0) & (kk < nz)]
        coords[inner] += d[inner]
    c = nid[:-1, :-1, :-1].reshape(-1)
    dx, dy, dz = 1, nx + 1, (nx + 1) * (ny + 1)
    tets = [np.stack([c, c + a, c + a + b, c + a + b + d3], axis=1)
            for a, b, d3 in ((dx, dy, dz), (dx, dz, dy), (dy, dx, dz), (dy, dz, dx), (dz, dx, dy), (dz, dy, dx))]
    return coords, np.stack(tets, axis=1).reshape(-1, 4).astype(np.int32)
