#!/usr/bin/env python
"""Small assemblies of every kind for `compute-sanitizer --tool memcheck python scripts/sanitize_small.py`."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finite_elements_b200 import mesh  # noqa: E402
from finite_elements_b200.device import (DeviceMesh, DeviceMesh3D, KIND_ELAST_PSTRESS, KIND_MAGNETIC, KIND_MASS,  # noqa: E402
                                         KIND_ELAST_TET, KIND_MASS_TET)

coords, conn = mesh.structured_mesh(67, 33, jitter=0.1, seed=1)
mid = (np.arange(len(conn)) >= len(conn) // 2).astype(np.int32)
mat = np.array([[1.0, 0.3, 1.0, 1.0], [2.0, 0.25, 0.5, 2.0]])
dm = DeviceMesh(coords, conn, mid, dim=2)
for v in (0, 4, 2, 1):
    k = dm.assemble(KIND_ELAST_PSTRESS, mat, variant=v)
    m = dm.assemble(KIND_MASS, mat, variant=v)
dm1 = DeviceMesh(coords, conn, mid, dim=1)
for v in (0, 4):
    dm1.assemble(KIND_MAGNETIC, np.array([[1.0, 0, 0, 0], [30.0, 0, 0, 0]]), variant=v)
c3, t3 = mesh.structured_tet_mesh(19, 7, 5, h=0.5, jitter=0.15, seed=4)   # 960 nodes: 60 tiles, several per CTA
d3 = DeviceMesh3D(c3, t3, (np.arange(len(t3)) % 2).astype(np.int32))
m3 = np.array([[210e9, 0.25, 1.0, 7860.0], [70e9, 0.3, 1.0, 2700.0]])
for v in (0, 6, 5, 4, 3, 2, 1):
    d3.assemble(KIND_ELAST_TET, m3, variant=v)
    d3.assemble(KIND_MASS_TET, m3, variant=v)
# a few PCG iterations on the tetrahedral matrix: the 3x3-block streamed SpMV (k_block_pattern3, k_spmv_stream3)
kv = d3.assemble(KIND_ELAST_TET, m3)
n3 = d3.n_rows
left = np.nonzero(c3[:, 0] == 0)[0]
bc3 = (3 * left[:, None] + np.arange(3)[None, :]).reshape(-1)
rhs3 = torch.ones(n3, dtype=torch.float64, device="cuda")
d3.dirichlet(kv, rhs3, bc3, np.zeros(len(bc3)))
x3 = torch.zeros(n3, dtype=torch.float64, device="cuda")
d3.pcg_fixed(kv, rhs3, x3, 10, work=d3.pcg_workspace())
# and on the 2-DOF matrix (k_spmv_stream)
kk = dm.assemble(KIND_ELAST_PSTRESS, mat)
bc2 = np.arange(0, 2 * 34, dtype=np.int64)
rhs2 = torch.ones(dm.n_rows, dtype=torch.float64, device="cuda")
dm.dirichlet(kk, rhs2, bc2, np.zeros(len(bc2)))
x2 = torch.zeros(dm.n_rows, dtype=torch.float64, device="cuda")
dm.pcg_fixed(kk, rhs2, x2, 10, work=dm.pcg_workspace())
torch.cuda.synchronize()
print("sanitize_small: done", float(k.abs().sum()), dm.fan_record_bytes)
