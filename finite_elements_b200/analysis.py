"""FiniteElementAnalysis -- the drop-in class surface of the reference's
finite_elements/analysis.py, with the hot path on the GPU.

What stays on the host (Python, O(boundary) work, same semantics as the reference):
  * flattening the mesh ONCE into coords / conn / material arrays;
  * load and boundary-condition records -> (dof, value) lists, including the reference's
    last-writer-wins de-duplication (analysis.py:28-47, :71-90) and its element->node /
    edge->node conversions (analysis.py:201-239, :407-447);
  * returning scipy / numpy containers of the reference's shapes.
What runs on the device through libfe_b200.so:
  * element matrices, CSR pattern, deterministic assembly (replaces analysis.py:324-339,
    :357-365, :387-405 and scipy's COO->CSR at :661);
  * element_to_node_factors for loaded elements (elements.py:18-53);
  * Dirichlet elimination + Jacobi-PCG (replaces spsolve, analysis.py:820-822) and the
    recovery of the Lagrange-multiplier tail so that Result.result_vector has the
    reference's length and meaning.
"""
import numpy as np

from . import conditions as fe_conditions
from . import core
from . import elements as fe_elements
from .core import DessiaObject
from .loads import NodeLoad
from .mesh import ArrayMesh, flatten_mesh
from .results import Result


def node_boundary_conditions_to_dict(node_boundary_conditions):
    """{(application, dimension): value}; a repeated key keeps its first position and takes
    the LAST value (the reference's `d[key] = + value` is an assignment, analysis.py:42-43)."""
    out = {}
    for condition in node_boundary_conditions:
        out[(condition.application, condition.dimension)] = condition.value
    return out


def node_boundary_from_dict(node_boundary_conditions_dict):
    return [fe_conditions.NodeBoundaryCondition(application=key[0], value=value, dimension=key[1])
            for key, value in node_boundary_conditions_dict.items()]


def node_loads_to_dict(node_loads):
    """Same last-writer-wins rule for loads (analysis.py:85-86)."""
    out = {}
    for load in node_loads:
        out[(load.node, load.dimension)] = load.value
    return out


def node_loads_from_dict(node_loads_dict):
    return [NodeLoad(node=key[0], value=value, dimension=key[1]) for key, value in node_loads_dict.items()]


class FiniteElements(DessiaObject):
    """Assembly-side half of the reference class (analysis.py:113-577)."""

    def __init__(self, mesh, element_loads, edge_loads, node_loads, magnet_loads, continuity_conditions,
                 node_boundary_conditions, edge_boundary_conditions, element_boundary_conditions,
                 plane_strain: bool = None, plane_stress: bool = None, *, device=0, solver_rtol=1e-12,
                 solver_maxit=None, assembly_variant=0):
        self.mesh = mesh
        self.element_loads = element_loads
        self.edge_loads = edge_loads
        self.node_loads = node_loads
        self.magnet_loads = magnet_loads
        self.continuity_conditions = continuity_conditions
        self.node_boundary_conditions = node_boundary_conditions
        self.edge_boundary_conditions = edge_boundary_conditions
        self.element_boundary_conditions = element_boundary_conditions
        self.plane_strain = plane_strain
        self.plane_stress = plane_stress
        # solver knobs (keyword-only extras; the reference has none because spsolve is direct)
        self.device = device
        self.solver_rtol = solver_rtol
        self.solver_maxit = solver_maxit
        self.assembly_variant = assembly_variant
        self.last_solve_info = None

        self._boundary_conditions = None
        self._node_loads = None
        self._positions = None
        self._flat = None
        self._device_mesh = None
        self._factor_cache = {}
        DessiaObject.__init__(self, name='')

    # ------------------------------------------------------------------ mesh flattening
    @property
    def dimension(self):
        if isinstance(self.mesh, ArrayMesh):
            return self.mesh.dimension
        return self.mesh.elements_groups[0].elements[0].dimension

    @property
    def elements_name(self):
        return self.mesh.elements_groups[0].elements[0].__class__.__name__

    def elements_permeability(self):
        return [group.elements[0].mu_total for group in self.mesh.elements_groups]

    def _flatten(self):
        """coords f64[N,2], conn i32[E,3], mat_id i32[E], mat f64[G,4], element -> flat index."""
        if self._flat is None:
            self._flat = flatten_mesh(self.mesh)
        return self._flat

    def _kind(self):
        from ._lib import KIND_ELAST_PSTRESS, KIND_ELAST_PSTRAIN, KIND_MAGNETIC, KIND_ELAST_TET
        if self._flatten()['magnetic']:
            return KIND_MAGNETIC
        fe_elements.check_plane_flags(self.plane_strain, self.plane_stress)
        if self.dimension == 3:   # tetrahedra: the flags are checked but do not enter (elements.py:753-771)
            return KIND_ELAST_TET
        return KIND_ELAST_PSTRAIN if self.plane_strain else KIND_ELAST_PSTRESS

    def _mass_kind(self):
        from ._lib import KIND_MASS, KIND_MASS_TET
        return KIND_MASS_TET if self.dimension == 3 else KIND_MASS

    def _dm(self):
        if self._device_mesh is None:
            from .device import DeviceMesh, DeviceMesh3D
            flat = self._flatten()
            if self.dimension == 3:
                self._device_mesh = DeviceMesh3D(flat['coords'], flat['conn'], flat['mat_id'], device=self.device)
                return self._device_mesh
            self._device_mesh = DeviceMesh(flat['coords'], flat['conn'], flat['mat_id'], dim=self.dimension,
                                           device=self.device)
        return self._device_mesh

    def _element_flat_index(self, element):
        flat = self._flatten()
        if flat['element_index'] is None:
            return int(element)  # ArrayMesh: elements are addressed by index
        return flat['element_index'][id(element)]

    def _factors_of(self, elements):
        """element_to_node_factors of the listed elements, evaluated on the device in one call."""
        flat = self._flatten()
        if flat['element_index'] is None or all(id(e) in flat['element_index'] for e in elements):
            idx = [self._element_flat_index(e) for e in elements]
            missing = [i for i in idx if i not in self._factor_cache]
            if missing:
                fac, _ = self._dm().source_factors(np.array(missing, dtype=np.int32))
                for i, row in zip(missing, fac.cpu().numpy()):
                    self._factor_cache[i] = tuple(row)
            return [self._factor_cache[i] for i in idx]
        return [e.element_to_node_factors() for e in elements]  # element not part of the mesh

    # ------------------------------------------------------------------ DOF numbering
    @property
    def positions(self):
        if not self._positions:
            self._positions = core.global_matrix_positions(dimension=self.dimension,
                                                           nodes_number=len(self.mesh.nodes))
        return self._positions

    def get_row_col_indices(self, element):
        """Rows / columns of the element's (3 dim)^2 triplets in row-major Ke order
        (analysis.py:714-735)."""
        positions = self.positions
        dofs = [positions[(self.mesh.node_to_index[point], d + 1)]
                for point in element.points for d in range(element.dimension)]
        row_ind = [dof for dof in dofs for _ in dofs]
        col_ind = dofs * len(dofs)
        return row_ind, col_ind

    # ------------------------------------------------------------------ triplet views
    def _matrix_data(self, kind):
        dm = self._dm()
        flat = self._flatten()
        data = dm.element_matrices(kind, flat['mat']).cpu().numpy().reshape(-1)
        dim = self.dimension
        npe = flat['conn'].shape[1]   # 3 (triangles) or 4 (tetrahedra) nodes per element
        dofs = (flat['conn'].astype(np.int64)[:, :, None] * dim + np.arange(dim)[None, None, :]).reshape(-1, npe * dim)
        row_ind = np.repeat(dofs, npe * dim, axis=1).reshape(-1)
        col_ind = np.tile(dofs, (1, npe * dim)).reshape(-1)
        return list(data), list(row_ind), list(col_ind)

    def k_matrix_data(self):
        """(data, row_ind, col_ind) triplets of all elements (analysis.py:324-339)."""
        return self._matrix_data(self._kind())

    def m_matrix_data(self):
        return self._matrix_data(self._mass_kind())

    def _assembled(self, kind):
        dm = self._dm()
        return dm.to_scipy(dm.assemble(kind, self._flatten()['mat'], variant=self.assembly_variant))

    def k_matrix(self, method_name):
        if method_name == 'dense':
            return self.k_matrix_dense()
        if method_name == 'sparse':
            return self.k_matrix_sparse()
        raise NotImplementedError(f'Class {self.__class__.__name__} does not implement {method_name} k matrix')

    def m_matrix(self, method_name):
        if method_name == 'dense':
            return self.m_matrix_dense()
        if method_name == 'sparse':
            return self.m_matrix_sparse()
        raise NotImplementedError(f'Class {self.__class__.__name__} does not implement {method_name} m matrix')

    def k_matrix_dense(self):
        return self._assembled(self._kind()).toarray()

    def k_matrix_sparse(self):
        """ndof x ndof K without condition rows, CSC like the reference (analysis.py:345-347)."""
        return self._assembled(self._kind()).tocsc()

    def m_matrix_dense(self):
        return self._assembled(self._mass_kind()).toarray()

    def m_matrix_sparse(self):
        return self._assembled(self._mass_kind()).tocsc()

    def matrix_dense(self, method_name):
        if method_name == 'k_matrix_data':
            return self.k_matrix_dense()
        if method_name == 'm_matrix_data':
            return self.m_matrix_dense()
        raise NotImplementedError(f'Class {self.__class__.__name__} does not implement {method_name}')

    def matrix_sparse(self, method_name):
        if method_name == 'k_matrix_data':
            return self.k_matrix_sparse()
        if method_name == 'm_matrix_data':
            return self.m_matrix_sparse()
        raise NotImplementedError(f'Class {self.__class__.__name__} does not implement {method_name}')

    # ------------------------------------------------------------------ conditions -> node records
    def boundary_conditions_element_to_node(self):
        out = []
        applications = [c.application for c in self.element_boundary_conditions]
        for condition, factors in zip(self.element_boundary_conditions, self._factors_of(applications)):
            for p_index, point in enumerate(self._points_of(condition.application)):
                out.append(fe_conditions.NodeBoundaryCondition(application=point,
                                                               value=condition.value * factors[p_index],
                                                               dimension=condition.dimension))
        return out

    def boundary_conditions_edge_to_node(self):
        out = []
        for condition in self.edge_boundary_conditions:
            for point in (condition.application.start, condition.application.end):
                out.append(fe_conditions.NodeBoundaryCondition(application=point, value=condition.value * 0.5,
                                                               dimension=condition.dimension))
        return out

    def _points_of(self, element):
        if self._flatten()['element_index'] is None:  # ArrayMesh: element index -> node indices
            return [int(n) for n in self._flatten()['conn'][int(element)]]
        return element.points

    def _node_key(self, node):
        """Records are de-duplicated on the node INDEX (equivalent to the reference's
        node-object key whenever node_to_index is consistent with node equality)."""
        return self.mesh.node_to_index[node]

    def _resolved_boundary_conditions(self):
        if self._boundary_conditions:
            return self._boundary_conditions
        records = list(self.node_boundary_conditions)
        records.extend(self.boundary_conditions_element_to_node())
        records.extend(self.boundary_conditions_edge_to_node())
        merged = {}
        for condition in records:
            merged[(self._node_key(condition.application), condition.dimension)] = (condition.application,
                                                                                      condition.value)
        self._boundary_conditions = [
            fe_conditions.NodeBoundaryCondition(application=app, value=value, dimension=key[1])
            for key, (app, value) in merged.items()]
        return self._boundary_conditions

    def _bc_arrays(self):
        conditions = self._resolved_boundary_conditions()
        positions = self.positions
        dofs = np.array([positions[(self._node_key(c.application), c.dimension)] for c in conditions],
                        dtype=np.int64)
        vals = np.array([c.value for c in conditions], dtype=np.float64)
        return dofs, vals

    def c_matrix_boundary_conditions(self):
        """(data, row_ind, col_ind) of the unit Lagrange rows (analysis.py:241-279)."""
        dofs, _ = self._bc_arrays()
        ndof = len(self.mesh.nodes) * self.dimension
        data, row_ind, col_ind = [], [], []
        for i, pos in enumerate(dofs.tolist()):
            data.extend((1, 1))
            row_ind.extend((ndof + i, pos))
            col_ind.extend((pos, ndof + i))
        return data, row_ind, col_ind

    def c_matrix_continuity_conditions(self):
        if self.continuity_conditions:
            raise NotImplementedError(
                'ContinuityCondition is out of scope: the reference indexes its rows at '
                'len(nodes) + len(node_loads) + i (analysis.py:190), which collides with the '
                'boundary-condition rows at len(nodes) * dimension + i (analysis.py:276)')
        return [], [], []

    # ------------------------------------------------------------------ loads -> node records
    def loads_element_to_node(self):
        out = []
        for elements_load in self.element_loads:
            factors = self._factors_of(elements_load.elements)
            for j, element in enumerate(elements_load.elements):
                for p_index, point in enumerate(self._points_of(element)):
                    out.append(NodeLoad(node=point, value=elements_load.value_per_element[j] * factors[j][p_index],
                                        dimension=elements_load.dimension))
        return out

    def loads_edge_to_node(self):
        out = []
        for edge_load in self.edge_loads:
            for point in (edge_load.edge.start, edge_load.edge.end):
                out.append(NodeLoad(node=point, value=edge_load.value * 0.5, dimension=edge_load.dimension))
        return out

    def source_c_matrix_loads(self):
        """(data, row_ind) of the nodal loads after the last-wins merge (analysis.py:457-489)."""
        records = list(self.node_loads)
        records.extend(self.loads_element_to_node())
        records.extend(self.loads_edge_to_node())
        merged = {}
        for load in records:
            merged[(self._node_key(load.node), load.dimension)] = (load.node, load.value)
        node_loads = [NodeLoad(node=node, value=value, dimension=key[1]) for key, (node, value) in merged.items()]
        if not self._node_loads:
            self._node_loads = node_loads
        positions = self.positions
        data = [load.source_c_matrix() for load in node_loads]
        row_ind = [positions[(self._node_key(load.node), load.dimension)] for load in node_loads]
        return data, row_ind

    def source_c_matrix_boundary_conditions(self):
        """(data, row_ind): condition values at rows ndof + i (analysis.py:509-543)."""
        conditions = self._resolved_boundary_conditions()
        ndof = len(self.mesh.nodes) * self.dimension
        return [c.source_c_matrix() for c in conditions], [ndof + i for i in range(len(conditions))]

    def source_c_matrix_magnet_loads(self):
        """Magnetisation contribution along magnet contours (analysis.py:556-577); host side,
        O(contour edges)."""
        data, row_ind = [], []
        for magnet_load in self.magnet_loads:
            for edge in magnet_load.contour_linear_elements():
                length = edge.length()
                normal = edge.interior_normal
                tangent = (-normal[1], normal[0])
                m = magnet_load.magnetization_vector
                share = (m[0] * tangent[0] + m[1] * tangent[1]) * length / 2
                for point in edge.points[:2]:
                    data.append(share)
                    row_ind.append(self.mesh.node_to_index[point])
        return data, row_ind


class FiniteElementAnalysis(FiniteElements):
    """create_matrix / create_source_matrix / solve / modal_analysis (analysis.py:580-830)."""

    def create_matrix(self):
        """Augmented system [[K, C^T], [C, 0]] as a canonical scipy CSR matrix, identical in
        pattern to the reference's csr_matrix((data, (row, col))) (analysis.py:617-663)."""
        import scipy.sparse as sp
        self.c_matrix_continuity_conditions()
        k = self._assembled(self._kind())
        dofs, _ = self._bc_arrays()
        n_bc = len(dofs)
        if n_bc == 0:
            return k
        ndof = k.shape[0]
        c = sp.csr_matrix((np.ones(n_bc), (np.arange(n_bc), dofs)), shape=(n_bc, ndof))
        matrix = sp.bmat([[k, c.T], [c, None]], format='csr')
        matrix.sum_duplicates()
        matrix.sort_indices()
        return matrix

    def get_source_matrix_length(self):
        return len(self.mesh.nodes) * self.dimension + len(self.continuity_conditions) \
            + len(self._resolved_boundary_conditions())

    def create_source_matrix(self):
        """Right-hand side of the augmented system, shape (ndof + n_bc, 1) (analysis.py:665-708)."""
        matrix = np.zeros((self.get_source_matrix_length(), 1))
        for method in (self.source_c_matrix_loads, self.source_c_matrix_magnet_loads,
                       self.source_c_matrix_boundary_conditions):
            data, row_ind = method()
            for value, row in zip(data, row_ind):
                matrix[row][0] += value
        return matrix

    def solve_arrays(self):
        """Device solve; returns the full result vector (solution, then multipliers) as numpy."""
        import torch
        self.c_matrix_continuity_conditions()
        dm = self._dm()
        kind = self._kind()
        flat = self._flatten()
        ndof = dm.n_rows
        source = self.create_source_matrix()[:, 0]
        bc_dofs, bc_vals = self._bc_arrays()
        dev = dm.ctx.device
        vals = dm.assemble(kind, flat['mat'], variant=self.assembly_variant)
        f = torch.as_tensor(source[:ndof].copy()).to(dev)
        rhs = f.clone()
        dm.dirichlet(vals, rhs, bc_dofs, bc_vals)
        u, iters, relres = dm.pcg(vals, rhs, rtol=self.solver_rtol, maxit=self.solver_maxit)
        self.last_solve_info = dict(iterations=iters, relative_residual=relres, ndof=ndof, nnz=dm.nnz)
        if relres > self.solver_rtol:   # accepted at attainable accuracy (fe_b200.h: <= max(100 rtol, 1e-8))
            import warnings
            warnings.warn(f"solve: PCG stopped at relative residual {relres:.2e} > solver_rtol "
                          f"{self.solver_rtol:.1e} (attainable accuracy)", RuntimeWarning, stacklevel=3)
        # multipliers: rows c of  K u + lambda = f  (the eliminated copy was overwritten in place)
        dm.assemble(kind, flat['mat'], out=vals, variant=self.assembly_variant)
        residual = f - dm.spmv(vals, u)
        lam = residual[torch.as_tensor(bc_dofs).to(dev)]
        return np.concatenate([u.cpu().numpy(), lam.cpu().numpy()])

    def solve(self):
        """F = K X -> Result(mesh, list(X)) of length ndof + n_bc (analysis.py:798-830).  A
        singular / non-SPD system raises NotImplementedError as the reference does."""
        return Result(self.mesh, list(self.solve_arrays()))

    def modal_analysis(self, order, k, constrained=False, tol=1e-9, maxit=5000, cheb_degree=None):
        """(eigvals, eigvecs.T) of  K x = lambda M x  (analysis.py:741-796): the k 'largest'
        (reference: eigsh(A=K, M=M, which='LM'), :779-782) or 'smallest' eigenvalues, ascending,
        eigenvectors M-orthonormal in the rows of the second array.

        Like the reference, the raw K and M are used -- boundary conditions are ignored -- unless
        `constrained=True` (extension), which restricts the pencil to the DOFs without a Dirichlet
        condition.  The reference's 'smallest' branch inverts the dense K (:790), singular for an
        unconstrained mesh; here 'smallest' is the mathematically intended lower end of the same
        pencil (rigid-body modes first when unconstrained), computed by LOBPCG on the device
        (modal.py).  cheb_degree > 1 turns the Jacobi preconditioner of the 'smallest' branch into a
        Chebyshev polynomial of that degree (fewer outer iterations on large meshes; None = automatic)."""
        import torch
        from .modal import modal_solve
        if order not in ('largest', 'smallest'):
            raise ValueError("Order parameter should be either 'largest' or 'smallest'")
        if self._flatten()['magnetic']:
            raise NotImplementedError('modal_analysis needs elements with a mass matrix (elements.py:513-536)')
        dm = self._dm()
        flat = self._flatten()
        k_vals = dm.assemble(self._kind(), flat['mat'], variant=self.assembly_variant)
        m_vals = dm.assemble(self._mass_kind(), flat['mat'], variant=self.assembly_variant)
        mask = None
        if constrained:
            bc_dofs, _ = self._bc_arrays()
            mask = torch.ones(dm.n_rows, dtype=torch.float64, device=k_vals.device)
            if len(bc_dofs):
                mask[torch.as_tensor(np.asarray(bc_dofs, dtype=np.int64)).to(k_vals.device)] = 0.0
        lam, vec, info = modal_solve(dm, k_vals, m_vals, int(k), order, mask=mask, tol=tol, maxit=maxit,
                                     cheb_degree=cheb_degree)
        self.last_modal_info = info
        if not info.converged:
            # the reference's eigsh raises ArpackNoConvergence here; the partial pairs stay available
            # through last_modal_info, but a caller must not mistake them for converged ones
            import warnings
            worst = float(np.max(info.residual_norms)) if info.residual_norms is not None else float('nan')
            warnings.warn(f"modal_analysis: LOBPCG stopped after {info.iterations} iterations, worst residual "
                          f"{worst:.2e} x the tolerance {tol:.1e}", RuntimeWarning, stacklevel=2)
        return lam.cpu().numpy(), vec.T.contiguous().cpu().numpy()
