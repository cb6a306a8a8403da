#!/usr/bin/env python
"""Single-GPU diagnostic: numeric assembly of ONE rank's block of the S16M mesh (owned-first / ghosts-after layout,
exactly what bench.py --gpus N gives rank r) against a stand-alone mesh of the same size."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from finite_elements_b200.device import DeviceMesh, KIND_ELAST_PSTRESS  # noqa: E402
from finite_elements_b200.dist import structured_rank_problem  # noqa: E402
from finite_elements_b200.mesh import structured_mesh_torch  # noqa: E402

dev = torch.device("cuda", 0)
mat = torch.as_tensor(np.array([[210e9, 0.3, 1.0, 7860.0]])).to(dev)
torch.cuda.set_stream(torch.cuda.Stream())


def time_asm(dm, variant=0):
    vals = torch.empty(dm.nnz, dtype=torch.float64, device=dev)
    for _ in range(3):
        dm.assemble(KIND_ELAST_PSTRESS, mat, out=vals, variant=variant)
    torch.cuda.synchronize()
    ts = []
    for _ in range(7):
        torch.cuda._sleep(300_000)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dm.assemble(KIND_ELAST_PSTRESS, mat, out=vals, variant=variant)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


nx, ny = 4096, 2048
for world in (2, 8):
    for rank in sorted({0, world // 2, world - 1}):
        lp, coords_local, bounds = structured_rank_problem(nx, ny, rank, world, dev)
        dm = DeviceMesh(coords_local, lp.conn_local, None, dim=2, device=0, n_owned=lp.n_owned)
        print(f"world {world} rank {rank}: owned {lp.n_owned} of {dm.n_nodes} nodes, {dm.n_elems} elements, records "
              f"{dm.fan_record_bytes} B, max degree {dm.max_degree}: assemble {time_asm(dm):.4f} ms "
              f"(8-byte forced {time_asm(dm, 4):.4f})", flush=True)
        del dm, lp, coords_local
    coords, conn = structured_mesh_torch(nx, ny // world, dev)
    dm = DeviceMesh(coords, conn, None, dim=2, device=0)
    print(f"stand-alone {nx}x{ny // world}: {dm.n_nodes} nodes, records {dm.fan_record_bytes} B: assemble {time_asm(dm):.4f} ms "
          f"(8-byte forced {time_asm(dm, 4):.4f})", flush=True)
    del dm
