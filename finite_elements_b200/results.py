"""Result containers and element post-processing (reference: finite_elements/results.py).

`Result` is the return type of the solve boundary (results.py:23-54).  The post-processing
classes keep the reference's names and dictionary-shaped accessors, but the per-element
arithmetic -- strain = B u_e, stress = D B u_e, energy = 1/2 u_e^T Ke u_e (results.py:809-830,
:769-781), magnetic B = (sum c_i A_i, -sum b_i A_i) (results.py:121-152) -- runs once for the
whole mesh in fe_elem_post on the device; `*_array` accessors expose the flat results.
Plotting / VTK output of the reference are out of scope.
"""
import numpy as np

from .core import DessiaObject
from .mesh import Vector2D, Vector3D, flatten_mesh


class Result(DessiaObject):
    """mesh + result_vector (length ndof + n_bc: solution, then the multipliers of the
    boundary conditions, exactly as the reference's augmented solve returns them)."""

    def __init__(self, mesh, result_vector):
        self.mesh = mesh
        self.result_vector = result_vector
        DessiaObject.__init__(self, name='')

    @property
    def dimension(self):
        if hasattr(self.mesh, 'dimension'):
            return self.mesh.dimension
        return self.mesh.elements_groups[0].elements[0].dimension


class _DevicePost(Result):
    """Shared plumbing: flatten the mesh once, run fe_elem_post once."""

    def __init__(self, mesh, result_vector, device=0):
        Result.__init__(self, mesh, result_vector)
        self._device = device
        self._flat = None
        self._post = None

    def _flatten(self):
        if self._flat is None:
            self._flat = flatten_mesh(self.mesh)
        return self._flat

    def _kind(self):
        raise NotImplementedError

    def _post_array(self):
        if self._post is None:
            from .device import element_post_arrays   # no plan: post-processing never needs the CSR pattern
            flat = self._flatten()
            ndof = len(flat['coords']) * self.dimension
            u = np.real(np.asarray(self.result_vector[:ndof], dtype=np.complex128)).astype(np.float64)
            self._post = element_post_arrays(self._kind(), flat['coords'], flat['conn'], flat['mat_id'], flat['mat'],
                                             u, device=self._device).cpu().numpy()
        return self._post

    def _elements(self):
        elements = self._flatten()['elements']
        if elements is None:
            raise TypeError('an ArrayMesh has no element objects: use the *_array accessors')
        return elements


class ElasticityResults(_DevicePost):
    """Displacements, strain, stress and energy of an elasticity solution (results.py:490-830)."""

    def __init__(self, mesh, result_vector, plane_strain, plane_stress, *, device=0):
        self.plane_strain = plane_strain
        self.plane_stress = plane_stress
        _DevicePost.__init__(self, mesh, result_vector, device)
        self._displacement_vectors_per_node = None
        self._displacements_per_element = None
        self._energy_per_element = None
        self._strain, self._stress = None, None

    def _kind(self):
        from . import _lib
        from .elements import check_plane_flags
        check_plane_flags(self.plane_strain, self.plane_stress)
        if self.dimension == 3:
            return _lib.KIND_ELAST_TET
        return _lib.KIND_ELAST_PSTRAIN if self.plane_strain else _lib.KIND_ELAST_PSTRESS

    @property
    def _ncomp(self):
        """Voigt components per element: 3 in the plane, 6 for tetrahedra."""
        return 6 if self.dimension == 3 else 3

    # ---- flat accessors ------------------------------------------------------------------
    @property
    def displacement_array(self):
        n = len(self.mesh.nodes)
        return np.real(np.asarray(self.result_vector[:n * self.dimension])).reshape(n, self.dimension)

    @property
    def strain_array(self):
        """f64[E,3] = (eps_xx, eps_yy, gamma_xy); tetrahedra: f64[E,6] = (xx, yy, zz, xy, yz, zx)."""
        return self._post_array()[:, 0:self._ncomp]

    @property
    def stress_array(self):
        """f64[E,3] = (sig_xx, sig_yy, tau_xy); tetrahedra: the six components in the strain order."""
        return self._post_array()[:, self._ncomp:2 * self._ncomp]

    @property
    def energy_array(self):
        return self._post_array()[:, 2 * self._ncomp]

    # ---- the reference's dictionary-shaped accessors -------------------------------------
    @property
    def displacement_vectors_per_node(self):
        """{node: Vector2D(u_x, u_y)} (results.py:639-674)."""
        if not self._displacement_vectors_per_node:
            d = self.displacement_array
            vec = Vector3D if self.dimension == 3 else Vector2D
            self._displacement_vectors_per_node = {node: vec(*(float(v) for v in d[n]))
                                                   for n, node in enumerate(self.mesh.nodes)}
        return self._displacement_vectors_per_node

    def displacement_per_node_x(self):
        return [float(v) for v in self.displacement_array[:, 0]]

    def displacement_per_node_y(self):
        return [float(v) for v in self.displacement_array[:, 1]]

    @property
    def displacements_per_element(self):
        """{element: [u0, v0, u1, v1, u2, v2]}; also sets element.displacements (results.py:677-714)."""
        if not self._displacements_per_element:
            conn = self._flatten()['conn']
            d = self.displacement_array
            out = {}
            for e, element in enumerate(self._elements()):
                element.displacements = [float(v) for v in d[conn[e]].reshape(-1)]
                out[element] = element.displacements
            self._displacements_per_element = out
        return self._displacements_per_element

    @property
    def energy_per_element(self):
        if not self._energy_per_element:
            energy = self.energy_array
            self._energy_per_element = {element: float(energy[e]) for e, element in enumerate(self._elements())}
        return self._energy_per_element

    @property
    def energy(self):
        return float(self.energy_array.sum())

    def _strain_stress(self):
        if not self._strain:
            strain, stress = self.strain_array, self.stress_array
            self._strain, self._stress = {}, {}
            for e, element in enumerate(self._elements()):
                element.strain = self._strain[element] = strain[e].copy()
                element.stress = self._stress[element] = stress[e].copy()
        return self._strain, self._stress

    @property
    def strain(self):
        return self._strain_stress()[0]

    @property
    def stress(self):
        return self._strain_stress()[1]


class ElasticityResults2D(ElasticityResults):
    """Component lists in element order (results.py:930-1016, :1435-1466)."""

    def axial_strain_x(self):
        return [float(v) for v in self.strain_array[:, 0]]

    def axial_strain_y(self):
        return [float(v) for v in self.strain_array[:, 1]]

    def shear_strain_xy(self):
        return [float(v) for v in self.strain_array[:, 2]]

    def axial_stress_x(self):
        return [float(v) for v in self.stress_array[:, 0]]

    def axial_stress_y(self):
        return [float(v) for v in self.stress_array[:, 1]]

    def shear_stress_xy(self):
        return [float(v) for v in self.stress_array[:, 2]]


class ElasticityResults3D(ElasticityResults):
    """Tetrahedra: component dictionaries {element: value} (results.py:1468-1720)."""

    def _component(self, which, k):
        data = self.strain if which == 'strain' else self.stress
        return {element: float(data[element][k]) for element in self._elements()}

    def axial_strain_x(self):
        return self._component('strain', 0)

    def axial_strain_y(self):
        return self._component('strain', 1)

    def axial_strain_z(self):
        return self._component('strain', 2)

    def shear_strain_xy(self):
        return self._component('strain', 3)

    def shear_strain_yz(self):
        return self._component('strain', 4)

    def shear_strain_zx(self):
        return self._component('strain', 5)

    def axial_stress_x(self):
        return self._component('stress', 0)

    def axial_stress_y(self):
        return self._component('stress', 1)

    def axial_stress_z(self):
        return self._component('stress', 2)

    def shear_stress_xy(self):
        return self._component('stress', 3)

    def shear_stress_yz(self):
        return self._component('stress', 4)

    def shear_stress_zx(self):
        return self._component('stress', 5)

    def displacement_per_node_z(self):
        return [float(v) for v in self.displacement_array[:, 2]]


class MagneticResults(_DevicePost):
    """Magnetic flux density per element from the nodal potentials (results.py:57-152)."""

    def __init__(self, mesh, result_vector, *, device=0):
        _DevicePost.__init__(self, mesh, result_vector, device)
        self._field = None

    def _kind(self):
        from . import _lib
        return _lib.KIND_MAGNETIC

    @property
    def magnetic_field_array(self):
        """f64[E,2] = (B_x, B_y)."""
        return self._post_array()

    @property
    def magnetic_field_per_element(self):
        if self._field is None:
            b = self.magnetic_field_array
            self._field = {element: Vector2D(float(b[e, 0]), float(b[e, 1]))
                           for e, element in enumerate(self._elements())}
        return self._field

    @property
    def magnetic_field_norm(self):
        return [float(v) for v in np.hypot(self.magnetic_field_array[:, 0], self.magnetic_field_array[:, 1])]
