"""Section timing of the LOBPCG path on the 1 M-triangle mesh (BASELINE configs[4]) for several polynomial degrees."""
import os, sys, time, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS, KIND_MASS
from finite_elements_b200.mesh import structured_mesh_torch
from finite_elements_b200.modal import modal_solve
torch.cuda.set_stream(torch.cuda.Stream())
MAT = np.array([[210e9, 0.25, 1.0, 7860.0]])
dev = torch.device("cuda", 0)
coords, conn = structured_mesh_torch(1024, 512, dev)
dm = DeviceMesh(coords, conn, None, dim=2)
kv = dm.assemble(KIND_ELAST_PSTRESS, MAT); mv = dm.assemble(KIND_MASS, MAT)
for deg in [int(a) if a != "auto" else None for a in sys.argv[1:]] or [None]:
    for prof in (True, False):
        if prof: os.environ["FE_B200_MODAL_PROF"] = "1"
        else: os.environ.pop("FE_B200_MODAL_PROF", None)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        lam, vec, info = modal_solve(dm, kv, mv, 10, "smallest", tol=1e-8, cheb_degree=deg)
        torch.cuda.synchronize(); t = time.perf_counter() - t0
        print(json.dumps({"degree": deg, "profiled": prof, "seconds": round(t, 3), "iterations": info.iterations,
                          "products": info.products, "converged": info.converged, "lam3": float(lam[3]), "prof": info.get("prof")}), flush=True)
