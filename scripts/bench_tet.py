#!/usr/bin/env python
"""Timing of the tetrahedral path on one GPU: numeric assembly of K on a structured
Kuhn mesh + Jacobi-PCG iterations, CUDA events, with the algorithmic-byte roofline fraction.
    python scripts/bench_tet.py [nx ny nz]"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finite_elements_b200.device import DeviceMesh3D, KIND_ELAST_TET  # noqa: E402
from finite_elements_b200.mesh import structured_tet_mesh  # noqa: E402

VARIANT = int(os.environ.get('FE_TET_VARIANT', '0'))
nx, ny, nz = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (96, 48, 48)
coords, conn = structured_tet_mesh(nx, ny, nz, h=1.0 / ny)
mat = torch.as_tensor(np.array([[210e9, 0.25, 1.0, 7860.0]])).cuda()
torch.cuda.set_stream(torch.cuda.Stream())
dm = DeviceMesh3D(coords, conn, None)
rowptr, colidx = dm.csr_pattern()
vals = torch.empty(dm.nnz, dtype=torch.float64, device="cuda")
ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
for _ in range(3):
    dm.assemble(KIND_ELAST_TET, mat, out=vals, variant=VARIANT)
torch.cuda.synchronize()
t = []
for _ in range(5):
    torch.cuda._sleep(200_000)
    a, b = ev(), ev()
    a.record()
    dm.assemble(KIND_ELAST_TET, mat, out=vals, variant=VARIANT)
    b.record()
    torch.cuda.synchronize()
    t.append(a.elapsed_time(b) * 1e-3)
t_asm = float(np.mean(t))
n = dm.n_rows
left = np.nonzero(coords[:, 0] == 0)[0]
bc = (3 * left[:, None] + np.arange(3)[None, :]).reshape(-1)
f = torch.zeros(n, dtype=torch.float64, device="cuda")
f[torch.as_tensor(3 * np.nonzero(coords[:, 0] == coords[:, 0].max())[0] + 2).cuda()] = -1000.0 / (ny * nz)
rhs = f.clone()
dm.dirichlet(vals, rhs, bc, np.zeros(len(bc)))
x = torch.zeros(n, dtype=torch.float64, device="cuda")
work = dm.pcg_workspace()
dm.pcg_fixed(vals, rhs, x, 50, work=work)
tp = []
for _ in range(3):
    x.zero_()
    a, b = ev(), ev()
    a.record()
    dm.pcg_fixed(vals, rhs, x, 50, work=work)
    b.record()
    torch.cuda.synchronize()
    tp.append(a.elapsed_time(b) * 1e-3 / 50)
t_it = float(np.mean(tp))
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
asm_bytes = 16.0 * len(conn) + 24.0 * len(coords) + 8.0 * dm.nnz
pcg_bytes = 12.0 * dm.nnz + 108.0 * n
print(json.dumps({"workload": f"{nx}x{ny}x{nz}-cell Kuhn mesh: {len(conn)} tetrahedra, {len(coords)} nodes, {n} DOF, nnz {dm.nnz}",
                  "assembly_ms": 1e3 * t_asm, "melem_per_s": len(conn) / t_asm / 1e6,
                  "assembly_roofline_frac": asm_bytes / t_asm / 1e9 / peak,
                  "pcg_ms_per_iter": 1e3 * t_it, "pcg_dof_iters_per_s": n / t_it,
                  "pcg_roofline_frac": pcg_bytes / t_it / 1e9 / peak, "max_degree": dm.max_degree}))
