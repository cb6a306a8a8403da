"""Boundary-condition records (reference: finite_elements/conditions.py)."""
from .core import DessiaObject


class BoundaryCondition(DessiaObject):
    """Imposes `value` on DOF `dimension` (1-based) of `application` (conditions.py:11-58).
    In the reference each condition is one Lagrange row with entries (1, 1)."""

    def __init__(self, application, value, dimension, name=''):
        self.application = application
        self.value = value
        self.dimension = dimension
        DessiaObject.__init__(self, name=name)

    def c_matrix(self):
        return (1, 1)

    def source_c_matrix(self):
        return self.value


class NodeBoundaryCondition(BoundaryCondition):
    """application = a node (conditions.py:61-107)."""


class EdgeBoundaryCondition(BoundaryCondition):
    """application = an edge with .start / .end; each end node receives value * 0.5 when the
    analysis converts it (analysis.py:222-239)."""

    def to_node_boundary_condition(self):
        return [NodeBoundaryCondition(point, self.value, self.dimension)
                for point in (self.application.start, self.application.end)]


class ElementBoundaryCondition(BoundaryCondition):
    """application = an element; node p receives value * element_to_node_factors()[p]
    (analysis.py:201-220)."""

    def to_node_boundary_condition(self):
        return [NodeBoundaryCondition(point, self.value, self.dimension)
                for point in self.application.points]


class ContinuityCondition(DessiaObject):
    """A(node1) = value * A(node2) (conditions.py:196-241).  Record only: the reference's
    row indexing for these collides with the boundary-condition rows (analysis.py:190 vs
    :276) and no shipped script uses them, so FiniteElementAnalysis rejects a non-empty list
    (SURVEY §2 #5)."""

    def __init__(self, node1, node2, value):
        self.node1 = node1
        self.node2 = node2
        self.value = value
        DessiaObject.__init__(self, name='')

    def c_matrix(self):
        return (1, 1, -self.value, -self.value)

    def source_c_matrix(self):
        return ()
