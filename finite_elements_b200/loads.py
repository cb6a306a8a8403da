"""Load records (reference: finite_elements/loads.py).  Plain attribute bags; the numerics
they imply (area weighting, element->node factors) are evaluated in analysis.py."""
from .core import DessiaObject


class ElementsLoad(DessiaObject):
    """A source `value` shared by `elements` in proportion to their areas
    (loads.py:24-35: value_per_element[j] = value * A_j / sum A)."""

    def __init__(self, elements, value, dimension):
        self.elements = elements
        self.value = value
        self.dimension = dimension
        total_area = sum(element.area for element in elements)
        self.value_per_element = [value * element.area / total_area for element in elements]
        DessiaObject.__init__(self, name='')


class ElementLoad(DessiaObject):
    def __init__(self, element, value, dimension):
        self.element = element
        self.value = value
        self.dimension = dimension
        DessiaObject.__init__(self, name='')


class EdgeLoad(DessiaObject):
    """`value` on an edge (an object with .start and .end); half goes to each end node
    (analysis.py:431-447)."""

    def __init__(self, edge, value, dimension):
        self.edge = edge
        self.value = value
        self.dimension = dimension
        DessiaObject.__init__(self, name='')


class NodeLoad(DessiaObject):
    """`value` added to the right-hand side at (node, dimension) (loads.py:83-102)."""

    def __init__(self, node, value, dimension):
        self.node = node
        self.value = value
        self.dimension = dimension
        DessiaObject.__init__(self, name='')

    def c_matrix(self):
        return ()

    def source_c_matrix(self):
        return self.value


class MagnetLoad(DessiaObject):
    """Magnetisation of a set of elements; contributes along the contour edges
    (loads.py:105-147, analysis.py:556-577).  Host-side only: O(boundary) work."""

    def __init__(self, elements, non_contour_nodes, magnetization_vector):
        self.elements = elements
        self.non_contour_nodes = non_contour_nodes
        self.magnetization_vector = magnetization_vector
        self.element_magnetization_vector = magnetization_vector / len(elements)
        DessiaObject.__init__(self, name='')

    def contour_linear_elements(self):
        """Edges that belong to exactly one element of the magnet and have at least one end
        outside `non_contour_nodes`."""
        seen = {}
        for element in self.elements:
            for edge in element.linear_elements:
                seen[edge] = seen.get(edge, 0) + 1
        return [edge for edge, count in seen.items()
                if count == 1 and not (edge.points[0] in self.non_contour_nodes
                                       and edge.points[1] in self.non_contour_nodes)]
