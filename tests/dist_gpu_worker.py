"""Run under torchrun (one rank per GPU): the row-block distributed path (fe_dist_pcg: NCCL halo
send/recv + dot all-reduce) must reproduce the single-GPU solve of the same problem.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29533 tests/dist_gpu_worker.py [nx ny]

Checks, per rank: local CSR rows == the owned rows of the global CSR (bit-exact values, columns
mapped back to global ids); distributed solution == single-GPU solution (1e-9 relative);
iteration counts within a few of each other (same algorithm, different reduction order).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    nx = int(sys.argv[1]) if len(sys.argv) > 1 else 96
    ny = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    magnetic = len(sys.argv) > 3 and sys.argv[3] == "mag"   # 1 DOF per node: scalar-CSR kernels
    tet = len(sys.argv) > 3 and sys.argv[3] == "tet"        # tetrahedra, 3 DOF per node (nx x ny x 4 cells)
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_stream(torch.cuda.Stream(device=dev))
    dist.init_process_group("nccl", device_id=dev)
    rank, world = dist.get_rank(), dist.get_world_size()

    from finite_elements_b200.device import DeviceMesh, DeviceMesh3D, KIND_ELAST_PSTRESS, KIND_MAGNETIC, KIND_ELAST_TET
    from finite_elements_b200.dist import DistributedMesh, partition_bounds, local_problem, localize_dofs
    from finite_elements_b200.mesh import structured_mesh

    if tet:
        from finite_elements_b200.mesh import structured_tet_mesh
        coords, conn = structured_tet_mesh(nx, ny, 4, h=1.0 / ny, jitter=0.2, seed=7)
    else:
        coords, conn = structured_mesh(nx, ny, jitter=0.2, seed=7)
    n_nodes = len(coords)
    if tet:
        kind, dim = KIND_ELAST_TET, 3
        mat = np.array([[210e9, 0.25, 1.0, 7860.0], [70e9, 0.33, 1.0, 2700.0]])
        mat_id = (np.arange(len(conn)) % 2).astype(np.int32)
    elif magnetic:
        kind, dim = KIND_MAGNETIC, 1
        mu0 = 4e-7 * np.pi
        mat = np.array([[mu0 * 100.0, 0, 0, 0], [mu0, 0, 0, 0], [mu0 * 30.0, 0, 0, 0]])
        mat_id = (np.arange(len(conn)) % 3).astype(np.int32)
    else:
        kind, dim = KIND_ELAST_PSTRESS, 2
        mat = np.array([[210e9, 0.25, 1.0, 7860.0], [70e9, 0.33, 0.5, 2700.0]])
        mat_id = (np.arange(len(conn)) % 2).astype(np.int32)
    # deliberately NOT aligned to grid lines: ranks cut through the middle of a line
    bounds = partition_bounds(n_nodes, world, align=1)
    lo, hi = bounds[rank], bounds[rank + 1]
    lp = local_problem(torch.as_tensor(conn).to(dev), lo, hi, bounds, dim)
    gid = lp.node_gid
    dmesh = DistributedMesh(torch.as_tensor(coords).to(dev)[gid], lp,
                            torch.as_tensor(mat_id).to(dev)[lp.elem_sel], device=local_rank)
    dm = dmesh.dm
    vals = dm.assemble(kind, mat)

    # global problem on every rank's own GPU (small) as the single-GPU reference
    gm = DeviceMesh3D(coords, conn, mat_id, device=local_rank) if tet else \
        DeviceMesh(coords, conn, mat_id, dim=dim, device=local_rank)
    gvals = gm.assemble(kind, mat)
    k_glob = gm.to_scipy(gvals)
    k_loc = dm.to_scipy(vals)
    gd = (gid.cpu().numpy()[:, None] * dim + np.arange(dim)[None, :]).reshape(-1)
    sub = k_glob[gd[:dm.n_rows]][:, gd]
    diff = abs(k_loc - sub)
    scale = abs(k_glob).max()
    assert diff.max() <= 1e-14 * scale, f"rank {rank}: local rows differ from global rows ({diff.max() / scale:.2e})"

    lines = np.nonzero(coords[:, 0] == 0)[0] if tet else np.arange(ny + 1) * (nx + 1)
    if tet:
        bc_g = (3 * lines[:, None] + np.arange(3)[None, :]).reshape(-1)
        f_g = np.zeros(3 * n_nodes)
        f_g[3 * (lines + nx) + 2] = -1000.0 / ny
    elif magnetic:
        bc_g = lines + nx                      # A = 0 on the right edge
        f_g = np.zeros(n_nodes)
        f_g[conn[:6].reshape(-1)] = 2.5e3      # source on the first few elements
    else:
        bc_g = np.stack([2 * lines, 2 * lines + 1], axis=1).reshape(-1)
        f_g = np.zeros(2 * n_nodes)
        f_g[2 * (lines + nx) + 1] = -1000.0 / ny
    bc_l, _ = localize_dofs(lp, bc_g)
    f = torch.as_tensor(f_g[gd[:dm.n_rows]].copy()).to(dev)
    rhs = f.clone()
    dm.dirichlet(vals, rhs, bc_l, torch.zeros(bc_l.numel(), dtype=torch.float64, device=dev))
    x, iters, relres = dmesh.pcg(vals, rhs, rtol=1e-11, raise_on_maxit=False)

    rhs_g = torch.as_tensor(f_g.copy()).to(dev)
    gm.dirichlet(gvals, rhs_g, bc_g, np.zeros(len(bc_g)))
    u, iters_g, relres_g = gm.pcg(gvals, rhs_g, rtol=1e-11, raise_on_maxit=False)
    u_own = u[torch.as_tensor(gd[:dm.n_rows]).to(dev)]
    err = float(torch.linalg.norm(x - u_own) / torch.linalg.norm(u))
    # rtol 1e-11 can be below the attainable FP64 residual of a slender beam: fe_pcg then stops at
    # stagnation (FE_OK, relres > rtol); what must hold is agreement with the single-GPU solve
    assert relres <= 1e-8 and err <= 1e-9, f"rank {rank}: err {err:.2e} relres {relres:.2e}"
    if relres <= 1e-11 and relres_g <= 1e-11:  # (restarts near the attainable residual are chaotic)
        assert abs(iters - iters_g) <= max(5, iters_g // 50), (iters, iters_g)
    # iteration counts at a tolerance both runs reach cleanly: same algorithm, different reduction order
    _, it8, rr8 = dmesh.pcg(vals, rhs, rtol=1e-8)
    _, it8_g, rr8_g = gm.pcg(gvals, rhs_g, rtol=1e-8)
    assert rr8 <= 1e-8 and rr8_g <= 1e-8 and abs(it8 - it8_g) <= max(3, it8_g // 100), (it8, it8_g, rr8, rr8_g)
    # fixed-iteration mode runs and keeps ranks in lock-step
    x2 = torch.zeros_like(x)
    dmesh.pcg(vals, rhs, x=x2, fixed_iters=7)
    t = torch.tensor([err], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"DIST-OK world={world} {'magnetic ' if magnetic else ('tetrahedral ' if tet else '')}mesh={nx}x{ny} iters={iters} (single GPU {iters_g}) max_err={t.item():.2e} "
              f"neighbours={len(lp.nbr_rank)}")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
