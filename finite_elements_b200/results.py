"""Result containers (reference: finite_elements/results.py:23-54).

Only `Result` -- the return type of the solve boundary -- is on the hot path.  The
post-processing classes of the reference (MagneticResults, ElasticityResults*) are plotting /
per-element Python dictionaries and are listed as "next" in SURVEY §8f.
"""
from .core import DessiaObject


class Result(DessiaObject):
    """mesh + result_vector (length ndof + n_bc: solution, then the multipliers of the
    boundary conditions, exactly as the reference's augmented solve returns them)."""

    def __init__(self, mesh, result_vector):
        self.mesh = mesh
        self.result_vector = result_vector
        DessiaObject.__init__(self, name='')

    @property
    def dimension(self):
        return self.mesh.elements_groups[0].elements[0].dimension
