"""Persistent PCG kernel against the three-kernel path: same problem, fixed iteration counts."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS
from finite_elements_b200.mesh import structured_mesh_torch

torch.cuda.set_stream(torch.cuda.Stream())
MAT = np.array([[210e9, 0.25, 1.0, 7860.0]])
for nx, ny in ((64, 32), (200, 100), (512, 256)):
    coords, conn = structured_mesh_torch(nx, ny, torch.device("cuda", 0))
    dm = DeviceMesh(coords, conn, None, dim=2)
    left = torch.arange(ny + 1, device="cuda") * (nx + 1)
    bc = torch.stack([2 * left, 2 * left + 1], dim=1).reshape(-1).int()
    f = torch.zeros(dm.n_rows, dtype=torch.float64, device="cuda")
    f[2 * (left + nx) + 1] = -1000.0 / ny
    vals = dm.assemble(KIND_ELAST_PSTRESS, MAT)
    rhs = f.clone()
    dm.dirichlet(vals, rhs, bc, torch.zeros(bc.numel(), dtype=torch.float64, device="cuda"))
    work = dm.pcg_workspace()
    for iters in (1, 2, 3, 10, 100):
        os.environ["FE_B200_PERSIST"] = "0"
        xl = dm.pcg_fixed(vals, rhs, torch.zeros_like(rhs), iters, work=work).clone()
        os.environ["FE_B200_PERSIST"] = "1"
        for g in (1, 3, 37, 148, 296):
            os.environ["FE_B200_PERSIST_GRID"] = str(g)
            xp = dm.pcg_fixed(vals, rhs, torch.zeros_like(rhs), iters, work=work)
            err = float(torch.linalg.norm(xp - xl) / torch.linalg.norm(xl))
            print(f"mesh {nx}x{ny} tiles {(dm.n_rows // 2 + 119) // 120:5d} iters {iters:3d} grid<= {g:3d}: rel diff to legacy {err:.3e}", flush=True)
        del os.environ["FE_B200_PERSIST_GRID"]
