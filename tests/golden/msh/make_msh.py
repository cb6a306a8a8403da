"""TEST INFRASTRUCTURE.  Re-serialises the gmsh meshes of the golden fixtures (coords / conn read from
the reference's scripts/InputFiles/{2D,3D}/*.msh by oracle/make_golden.py) as gmsh 4.1 ASCII files
with finite_elements_b200.mesh.write_gmsh41: several node entity blocks, non-contiguous tags
(tag = 3 i + 7), a block of line elements in front of the triangles -- what a reader of real gmsh
output has to cope with.  The GPU box has no /root/reference; these files let the tests drive
scripts/Elasticity/beam2d_example_3.py:52-106 and beam3d_example_2.py:52-98 from a .msh file.

    python tests/golden/msh/make_msh.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(HERE))))

from finite_elements_b200.mesh import read_gmsh41, write_gmsh41  # noqa: E402

NAMES = ("gmsh_beam_0.8", "gmsh_beam_0.5", "gmsh_beam_0.3", "gmsh_beam_0.18", "gmsh_beam_0.1",
         "gmsh_beam3d_1", "gmsh_beam3d_0.5")

if __name__ == "__main__":
    for name in NAMES:
        z = np.load(os.path.join(os.path.dirname(HERE), name + ".npz"))
        c, t = z["coords"], z["conn"]
        path = os.path.join(HERE, name + ".msh")
        write_gmsh41(path, c, t, node_blocks=9 if c.shape[1] == 2 else 4,
                     boundary_edges=[(0, 1), (1, 2)] if c.shape[1] == 2 else [], tag_of=lambda i: 3 * i + 7)
        c2, t2 = read_gmsh41(path, tetrahedra=c.shape[1] == 3)
        assert np.array_equal(c, c2) and np.array_equal(t, t2), name
        print(name, c.shape, t.shape, os.path.getsize(path))
