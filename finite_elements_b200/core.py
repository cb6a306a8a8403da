"""Constants, DOF numbering and the Material record (reference: finite_elements/core.py).

Only what the hot path needs is mirrored: MU (core.py:31), global_matrix_positions
(core.py:89-108) and Material (core.py:139-171).  The colour-map / matplotlib helpers of
the reference's core.py are plotting code and out of scope.
"""
import math

MU = 4 * math.pi * 1e-7


class DessiaObject:
    """Minimal stand-in for dessia_common.core.DessiaObject (a name-carrying base class);
    the real one is used instead when dessia_common is importable."""

    def __init__(self, name='', **kwargs):
        self.name = name


try:  # pragma: no cover - dessia_common is not in this image
    from dessia_common.core import DessiaObject  # noqa: F811,F401
except Exception:  # noqa: BLE001
    pass


class _Positions:
    """Mapping (node_index, dimension_1_based) -> node_index * dim + dimension - 1.

    The reference materialises this as a dict with nodes_number * dimension entries
    (core.py:102-106); here it is the closed form with the same look-up surface."""

    def __init__(self, dimension, nodes_number):
        self.dimension = dimension
        self.nodes_number = nodes_number

    def __getitem__(self, key):
        node, d = key
        if not (0 <= node < self.nodes_number and 1 <= d <= self.dimension):
            raise KeyError(key)
        return node * self.dimension + (d - 1)

    def __len__(self):
        return self.dimension * self.nodes_number

    def __contains__(self, key):
        try:
            self[key]
            return True
        except (KeyError, TypeError, ValueError):
            return False

    def __bool__(self):
        return True

    def keys(self):
        return ((i, j + 1) for i in range(self.nodes_number) for j in range(self.dimension))

    def items(self):
        return ((k, self[k]) for k in self.keys())


def global_matrix_positions(dimension, nodes_number):
    return _Positions(dimension, nodes_number)


class Material(DessiaObject):
    """elasticity_modulus, poisson_ratio, mass_density (core.py:139-171)."""

    def __init__(self, elasticity_modulus, poisson_ratio, mass_density, name=''):
        self.elasticity_modulus = elasticity_modulus
        self.poisson_ratio = poisson_ratio
        self.mass_density = mass_density
        DessiaObject.__init__(self, name=name)
