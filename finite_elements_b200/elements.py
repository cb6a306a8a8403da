"""Finite-element records (reference: finite_elements/elements.py).

In the reference every element object precomputes B (3x6) and both D (3x3) with numpy in
its constructor (elements.py:241-243) and `elementary_matrix` multiplies them per element
(elements.py:508-509) -- that is HOT LOOP 1.  Here the objects are plain
records (points + material); FiniteElementAnalysis flattens them once and the matrices are
computed by the CUDA kernels for the whole mesh.  The per-element methods below keep the
reference's names / return shapes and evaluate through the same kernels
(fe_elem_matrices, fe_source_factors) on a one-element batch.
"""
import numpy as np

from . import core
from .core import DessiaObject
from .mesh import TriangularElement2D, TetrahedralElement


def _points_array(points):
    return np.array([[p[0], p[1]] for p in points], dtype=np.float64)


def _one_element_matrix(kind, points, mat_row):
    from .device import DeviceMesh
    dm = DeviceMesh(_points_array(points), np.array([[0, 1, 2]], dtype=np.int32), None,
                    dim=1 if kind == 2 else 2)
    return dm.element_matrices(kind, np.array([mat_row], dtype=np.float64)).cpu().numpy()[0]


def check_plane_flags(plane_strain, plane_stress):
    """Exactly one flag must be set (elements.py:265-273)."""
    if plane_strain and plane_stress:
        raise ValueError('just one of plane_strain or plane_stress can be True')
    if not plane_strain and not plane_stress:
        raise ValueError('one of plane_strain or plane_stress must be True')


class Element2D(TriangularElement2D):
    """P1 triangle; element_to_node_factors as the reference defines them
    (elements.py:18-53): |det| * N_i(midpoint of points[1], points[2]), i.e. (~0, A, A)."""

    def element_to_node_factors(self):
        from .device import DeviceMesh
        dm = DeviceMesh(_points_array(self.points), np.array([[0, 1, 2]], dtype=np.int32), None, dim=1)
        fac, _ = dm.source_factors()
        return tuple(fac.cpu().numpy()[0])


class MagneticElement2D(Element2D):
    """1 DOF per node, permeability mu_total (elements.py:56-191)."""

    def __init__(self, triangular_element, mu_total, name=''):
        self.triangular_element = triangular_element
        TriangularElement2D.__init__(self, points=triangular_element.points, name=name)
        self.mu_total = mu_total

    @property
    def dimension(self):
        return 1

    def elementary_matrix(self):
        """9-tuple, row-major: (1/mu) (b_i b_j + c_i c_j) * area (elements.py:93-118)."""
        return tuple(_one_element_matrix(2, self.points, (self.mu_total, 0.0, 0.0, 0.0)))


class ElasticityElement(DessiaObject):
    """Material + state shared by the elasticity elements (elements.py:194-330)."""

    def __init__(self, mesh_element, elasticity_modulus, poisson_ratio, mass_density,
                 displacements=None, stress=None, strain=None, name=''):
        self.mesh_element = mesh_element
        self.elasticity_modulus = elasticity_modulus
        self.poisson_ratio = poisson_ratio
        self.mass_density = mass_density
        self.points = self.mesh_element.points
        self.displacements = displacements
        self.stress = stress
        self.strain = strain
        DessiaObject.__init__(self, name=name)

    # The reference stores these three as constructor-time numpy arrays; here they are
    # derived on demand for post-processing only -- assembly never reads them.
    @property
    def b_matrix(self):
        return self._b_matrix()

    @property
    def d_matrix_plane_strain(self):
        return self._d_matrix(plane_strain=True)

    @property
    def d_matrix_plane_stress(self):
        return self._d_matrix(plane_strain=False)

    def _d_matrix(self, plane_strain):
        e_mod, nu = self.elasticity_modulus, self.poisson_ratio
        lam = e_mod * nu / ((1 + nu) * (1 - 2 * nu)) if plane_strain else e_mod * nu / (1 - nu ** 2)
        shear = e_mod / (2 * (1 + nu))
        return np.array([[lam + 2 * shear, lam, 0.0], [lam, lam + 2 * shear, 0.0], [0.0, 0.0, shear]])

    def d_matrix(self, plane_strain, plane_stress):
        check_plane_flags(plane_strain, plane_stress)
        return self.d_matrix_plane_strain if plane_strain else self.d_matrix_plane_stress

    def energy(self, plane_strain, plane_stress):
        """0.5 u_e^T Ke u_e (elements.py:275-292)."""
        size = self.dimension * len(self.mesh_element.points)
        u = np.asarray(self.displacements, dtype=np.float64)
        return 0.5 * u @ self.elementary_matrix(plane_strain, plane_stress).reshape(size, size) @ u

    @classmethod
    def with_material_object(cls, mesh_element, material: core.Material, displacements=None, stress=None,
                             strain=None, name=''):
        return cls(mesh_element=mesh_element, elasticity_modulus=material.elasticity_modulus,
                   poisson_ratio=material.poisson_ratio, mass_density=material.mass_density,
                   displacements=displacements, stress=stress, strain=strain, name=name)


class ElasticityTriangularElement2D(ElasticityElement, Element2D):
    """2 DOF per node, local DOF order [u0, v0, u1, v1, u2, v2] (elements.py:333-660)."""

    def __init__(self, mesh_element, elasticity_modulus, poisson_ratio, mass_density, thickness=1.0,
                 displacements=None, stress=None, strain=None, name=''):
        self.thickness = thickness
        ElasticityElement.__init__(self, mesh_element, elasticity_modulus, poisson_ratio, mass_density,
                                   displacements=displacements, stress=stress, strain=strain, name=name)
        TriangularElement2D.__init__(self, points=mesh_element.points, name=name)

    @property
    def dimension(self):
        return 2

    def _b_matrix(self):
        """Strain-displacement matrix (elements.py:395-416), post-processing helper."""
        (x0, y0), (x1, y1), (x2, y2) = ((p[0], p[1]) for p in self.points)
        beta = (y1 - y2, y2 - y0, y0 - y1)
        gamma = (x2 - x1, x0 - x2, x1 - x0)
        det = (x0 - x2) * (y1 - y2) - (y0 - y2) * (x1 - x2)
        out = np.zeros((3, 6))
        out[0, 0::2], out[1, 1::2] = beta, gamma
        out[2, 0::2], out[2, 1::2] = gamma, beta
        return out / det

    def _material_row(self):
        return (self.elasticity_modulus, self.poisson_ratio, self.thickness, self.mass_density)

    def elementary_matrix(self, plane_strain, plane_stress):
        """thickness * area * B^T D B flattened to 36 (elements.py:466-511)."""
        check_plane_flags(plane_strain, plane_stress)
        return _one_element_matrix(1 if plane_strain else 0, self.points, self._material_row())

    def elementary_mass_matrix(self):
        """Consistent mass flattened to 36 (elements.py:513-536)."""
        return _one_element_matrix(3, self.points, self._material_row())

    @classmethod
    def from_element(cls, mesh_element, elasticity_element):
        return cls(mesh_element, elasticity_element.elasticity_modulus, elasticity_element.poisson_ratio,
                   elasticity_element.mass_density, elasticity_element.thickness)


class ElasticityTetrahedralElement3D(ElasticityElement, TetrahedralElement):
    """Linear tetrahedron, 3 DOF per node, local DOF order [u0, v0, w0, u1, ...]
    (elements.py:663-876).  The plane flags are accepted and ignored, as in the reference
    (both `_d_matrix_plane_*` return the 3D matrix, :753-771)."""

    def __init__(self, mesh_element, elasticity_modulus, poisson_ratio, mass_density,
                 displacements=None, stress=None, strain=None, name=''):
        ElasticityElement.__init__(self, mesh_element, elasticity_modulus, poisson_ratio, mass_density,
                                   displacements=displacements, stress=stress, strain=strain, name=name)
        TetrahedralElement.__init__(self, points=mesh_element.points, name=name)

    @property
    def dimension(self):
        return 3

    def _b_matrix(self):
        """6 x 12 strain-displacement matrix (elements.py:719-751), post-processing helper."""
        f = self.form_functions
        out = np.zeros((6, 12))
        for i in range(4):
            a, b, c = f[i][1], f[i][2], f[i][3]
            out[0, 3 * i], out[1, 3 * i + 1], out[2, 3 * i + 2] = a, b, c
            out[3, 3 * i], out[3, 3 * i + 1] = b, a
            out[4, 3 * i + 1], out[4, 3 * i + 2] = c, b
            out[5, 3 * i], out[5, 3 * i + 2] = c, a
        return out / (6 * self.volume)

    def _d_matrix(self, plane_strain=None):
        """elements.py:773-797."""
        e_mod, nu = self.elasticity_modulus, self.poisson_ratio
        d = np.zeros((6, 6))
        d[:3, :3] = nu
        d[np.arange(3), np.arange(3)] = 1 - nu
        d[np.arange(3, 6), np.arange(3, 6)] = (1 - 2 * nu) / 2
        return e_mod / ((1 + nu) * (1 - 2 * nu)) * d

    def d_matrix(self, plane_strain, plane_stress):
        check_plane_flags(plane_strain, plane_stress)
        return self._d_matrix()

    def _one(self, kind):
        from .device import DeviceMesh3D
        pts = np.array([[p[0], p[1], p[2]] for p in self.points], dtype=np.float64)
        dm = DeviceMesh3D(pts, np.array([[0, 1, 2, 3]], dtype=np.int32), None)
        row = np.array([[self.elasticity_modulus, self.poisson_ratio, 1.0, self.mass_density]], dtype=np.float64)
        return dm.element_matrices(kind, row).cpu().numpy()[0]

    def elementary_matrix(self, plane_strain, plane_stress):
        """volume * B^T D B flattened to 144 (elements.py:809-828)."""
        return self._one(4)

    def elementary_mass_matrix(self):
        """(rho V / 20) ((1 + delta_ij) (x) I3) flattened to 144 (elements.py:830-857)."""
        return self._one(5)

    @classmethod
    def from_element(cls, mesh_element, elasticity_element):
        return cls(mesh_element, elasticity_element.elasticity_modulus, elasticity_element.poisson_ratio,
                   elasticity_element.mass_density)
