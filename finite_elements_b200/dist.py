"""Multi-GPU host logic: one process per GPU, contiguous row (node) blocks.

Rank r owns the nodes [bounds[r], bounds[r+1]) and therefore the CSR rows of their DOFs.
It assembles those rows from EVERY element incident to an owned node (one layer of ghost
elements, no communication), numbers its local columns owned-first / ghosts-after, and during
PCG exchanges the interface values of the search direction with the ranks that own its ghosts
(NCCL send/recv over NVLink) and all-reduces the dot products.

Everything in `local_problem` is plain index arithmetic on torch tensors, so it runs on CPU
tensors too -- tests/test_dist_cpu.py drives it with world_size 2 over gloo.

The halo lists need no negotiation: both sides derive them from the elements they hold.
Rank s needs node a of rank r  <=>  some element contains a (owned by r) and a node owned by
s; rank r holds all elements incident to its own nodes, hence sees every such pair, and both
ranks order the list by global node id.
"""
import ctypes as C
import os
import time

import numpy as np
import torch

from . import _lib
from ._lib import check, lib


def partition_bounds(n_nodes, world, align=1):
    """Contiguous, balanced node ranges; interior bounds are multiples of `align` (use the grid
    line length nx+1 on structured meshes so that a rank owns whole lines)."""
    units = (n_nodes + align - 1) // align
    bounds = [min(n_nodes, ((units * r) // world) * align) for r in range(world)] + [n_nodes]
    return [int(b) for b in bounds]


class LocalProblem:
    """Index data of one rank (all torch tensors on the device of `conn`)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


def local_problem(conn, lo, hi, bounds, dim, elem_offset=0):
    """conn: int tensor [E_sub, 3] (triangles) or [E_sub, 4] (tetrahedra) of GLOBAL node ids containing at least every element
    incident to a node in [lo, hi).  Returns a LocalProblem with
      elem_sel   indices (into conn) of the local elements, + elem_offset = global element ids
      conn_local int32 [E_loc, 3] in local numbering (owned: g - lo; ghosts: n_owned + k)
      node_gid   int64 [n_local] global id of every local node (owned range, then sorted ghosts)
      nbr_rank, send_ptr, recv_ptr (host int32 numpy), send_idx (int32 tensor, local DOFs)."""
    conn = conn.long()
    dev = conn.device
    own = (conn >= lo) & (conn < hi)
    sel = own.any(dim=1)
    lc = conn[sel]
    lown = own[sel]
    n_owned = hi - lo
    nodes = torch.unique(lc)
    ghosts = nodes[(nodes < lo) | (nodes >= hi)]
    bnd = torch.as_tensor(bounds, device=dev, dtype=torch.long)
    g_owner = torch.searchsorted(bnd, ghosts, right=True) - 1
    # local numbering
    gpos = torch.searchsorted(ghosts, lc.reshape(-1)).reshape(lc.shape)
    conn_local = torch.where(lown, lc - lo, n_owned + gpos).to(torch.int32).contiguous()
    node_gid = torch.cat([torch.arange(lo, hi, device=dev), ghosts])
    # receive side: ghosts are sorted by global id, hence grouped by (ascending) owner
    nbr_rank, counts = torch.unique_consecutive(g_owner, return_counts=True)
    recv_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(counts, 0)])
    # send side: every (owned a, foreign b) pair inside an element -> rank(b) needs a
    keys = []
    npe = conn.shape[1]
    for v in range(npe):
        for w in range(npe):
            if v == w:
                continue
            m = lown[:, v] & ~lown[:, w]
            if bool(m.any()):
                a = lc[m, v]
                s = torch.searchsorted(bnd, lc[m, w], right=True) - 1
                keys.append(s * (hi - lo + 1) + (a - lo))
    if keys:
        key = torch.unique(torch.cat(keys))
        s_rank = key // (hi - lo + 1)
        s_node = key % (hi - lo + 1)
        s_nbr, s_counts = torch.unique_consecutive(s_rank, return_counts=True)
    else:
        s_node = torch.zeros(0, dtype=torch.long, device=dev)
        s_nbr = torch.zeros(0, dtype=torch.long, device=dev)
        s_counts = torch.zeros(0, dtype=torch.long, device=dev)
    if not torch.equal(s_nbr, nbr_rank):
        raise RuntimeError("partition: send and receive neighbour sets differ (element set incomplete?)")
    send_ptr = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), torch.cumsum(s_counts, 0)])
    send_idx = (s_node[:, None] * dim + torch.arange(dim, device=dev)[None, :]).reshape(-1).to(torch.int32)
    return LocalProblem(
        lo=lo, hi=hi, n_owned=n_owned, n_local=int(node_gid.numel()), dim=dim,
        elem_sel=torch.nonzero(sel).reshape(-1) + elem_offset, conn_local=conn_local, node_gid=node_gid,
        nbr_rank=nbr_rank.cpu().numpy().astype(np.int32),
        send_ptr=(send_ptr * dim).cpu().numpy().astype(np.int32),
        recv_ptr=(recv_ptr * dim).cpu().numpy().astype(np.int32),
        send_idx=send_idx.contiguous())


def localize_dofs(lp, global_dofs, values=None):
    """Global DOF list -> the entries present on this rank (owned or ghost), local numbering."""
    g = torch.as_tensor(global_dofs, device=lp.node_gid.device).long()
    node, d = g // lp.dim, g % lp.dim
    owned = (node >= lp.lo) & (node < lp.hi)
    ghosts = lp.node_gid[lp.n_owned:]
    if ghosts.numel():
        gp = torch.searchsorted(ghosts, node).clamp(max=int(ghosts.numel()) - 1)
        is_ghost = ~owned & (ghosts[gp] == node)
    else:
        gp = torch.zeros_like(node)
        is_ghost = torch.zeros_like(owned)
    local_node = torch.where(owned, node - lp.lo, lp.n_owned + gp)
    keep = owned | is_ghost
    out = (local_node * lp.dim + d)[keep].to(torch.int32)
    if values is None:
        return out, keep
    return out, torch.as_tensor(values, device=g.device, dtype=torch.float64)[keep]


# ---------------------------------------------------------------------------------------
# device side
# ---------------------------------------------------------------------------------------
def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def init_nccl(ctx):
    """fe_dist_init with a unique id broadcast over the existing torch.distributed group."""
    import torch.distributed as dist
    if getattr(ctx, "_dist_ready", False):
        return
    rank, world = dist.get_rank(), dist.get_world_size()
    buf = (C.c_ubyte * 128)()
    if rank == 0:
        check(lib.fe_dist_unique_id(C.byref(buf)))
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device=ctx.device)
    dist.broadcast(t, src=0)
    raw = bytes(t.cpu().tolist())
    with torch.cuda.device(ctx.device):
        check(lib.fe_dist_init(ctx.handle, C.c_char_p(raw), rank, world))
    ctx._dist_ready = True


def init_peer_memory(ctx, lp):
    """Peer-memory transport (fe_dist_p2p_export / _import): exchange the CUDA IPC handles of the
    ranks' communication blocks and work out where this rank's interface values land inside each
    neighbour's ghost block.  Returns peer_dst_off (int32[n_nbr]) or None when unavailable
    (more ranks than the library maps, or FE_B200_NO_P2P set)."""
    import torch.distributed as dist
    rank, world = dist.get_rank(), dist.get_world_size()
    if world < 2 or world > 16 or os.environ.get("FE_B200_NO_P2P"):
        return None
    dev = ctx.device
    n_ghost = int(lp.recv_ptr[-1]) if len(lp.recv_ptr) else 0
    handle = (C.c_ubyte * 64)()
    with torch.cuda.device(dev):
        check(lib.fe_dist_p2p_export(ctx.handle, n_ghost, C.byref(handle)))
    mine = torch.tensor(list(bytes(handle)), dtype=torch.uint8, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    raw = b"".join(bytes(t.cpu().tolist()) for t in gathered)
    with torch.cuda.device(dev):
        check(lib.fe_dist_p2p_import(ctx.handle, C.c_char_p(raw)))
    # table[r][s] = offset (DOFs) inside rank r's ghost block of the values owned by rank s
    row = torch.full((world,), -1, dtype=torch.int64, device=dev)
    for k, s in enumerate(lp.nbr_rank.tolist()):
        row[s] = int(lp.recv_ptr[k])
    table = [torch.empty_like(row) for _ in range(world)]
    dist.all_gather(table, row)
    off = np.array([int(table[s][rank].item()) for s in lp.nbr_rank.tolist()] or [0], dtype=np.int32)
    if len(lp.nbr_rank) and (off < 0).any():
        raise RuntimeError("partition: a neighbour does not list this rank as its neighbour")
    return off


class DistributedMesh:
    """The rank-local DeviceMesh + halo description; pcg() runs fe_dist_pcg."""

    def __init__(self, coords_local, lp, mat_id_local=None, device=0, ctx=None):
        from .device import DeviceMesh, DeviceMesh3D, Context
        self.ctx = ctx or Context.get(device)
        self.lp = lp
        mesh_cls = DeviceMesh3D if lp.dim == 3 else DeviceMesh   # tetrahedra / triangles
        self.dm = mesh_cls(coords_local, lp.conn_local, mat_id_local, dim=lp.dim, device=device,
                           n_owned=lp.n_owned, ctx=self.ctx)
        self.send_idx = lp.send_idx.to(self.ctx.device)
        self._nbr = np.ascontiguousarray(lp.nbr_rank, dtype=np.int32)
        self._sp = np.ascontiguousarray(lp.send_ptr, dtype=np.int32)
        self._rp = np.ascontiguousarray(lp.recv_ptr, dtype=np.int32)
        t0 = time.perf_counter()
        init_nccl(self.ctx)
        self._dst_off = init_peer_memory(self.ctx, lp)  # None -> NCCL transport inside the loop
        torch.cuda.synchronize(self.ctx.device)
        self.peer_setup_ms = 1e3 * (time.perf_counter() - t0)

    def pcg(self, vals, b, x=None, rtol=1e-8, maxit=None, fixed_iters=0, work=None, raise_on_maxit=True):
        dm = self.dm
        rowptr, colidx = dm._vouch_pattern()
        if x is None:
            x = torch.zeros(dm.n_rows, dtype=torch.float64, device=self.ctx.device)
        if work is None:
            work = dm.pcg_workspace()
        if maxit is None:
            maxit = 10 ** 7
        iters, relres = C.c_int32(0), C.c_double(0.0)
        ip = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
        with torch.cuda.device(self.ctx.device):
            rc = lib.fe_dist_pcg(self.ctx.handle, C.c_void_p(torch.cuda.current_stream().cuda_stream), dm.n_rows,
                                 dm.n_cols, _ptr(rowptr), _ptr(colidx), _ptr(vals), _ptr(b), _ptr(x), _ptr(work),
                                 len(self._nbr), ip(self._nbr), ip(self._sp), _ptr(self.send_idx), ip(self._rp),
                                 ip(self._dst_off) if self._dst_off is not None else C.c_void_p(0),
                                 dm.block_dim, float(rtol), int(maxit), int(fixed_iters), C.byref(iters), C.byref(relres))
        if rc == _lib.FE_ERR_NOT_CONVERGED and not raise_on_maxit:
            return x, iters.value, relres.value
        check(rc)
        return x, iters.value, relres.value


def _gpu_spin(torch_mod, cycles=600_000):
    spin = getattr(torch_mod.cuda, "_sleep", None)
    if spin is not None:
        spin(cycles)


def structured_rank_problem(nx, ny, rank, world, dev, dim=2):
    """Rank-local piece of the SURVEY §8d structured mesh, built on the device without ever
    materialising the global connectivity: node lines [j0, j1) are owned, cell rows
    [j0-1, j1) hold every element incident to them."""
    from .mesh import structured_mesh_torch
    n_nodes = (nx + 1) * (ny + 1)
    bounds = partition_bounds(n_nodes, world, align=nx + 1)
    lo, hi = bounds[rank], bounds[rank + 1]
    j0, j1 = lo // (nx + 1), hi // (nx + 1)
    r_lo, r_hi = max(j0 - 1, 0), min(j1, ny)
    coords, conn = structured_mesh_torch(nx, ny, dev, row_lo=r_lo, row_hi=r_hi)
    lp = local_problem(conn, lo, hi, bounds, dim, elem_offset=2 * nx * r_lo)
    return lp, coords[lp.node_gid].contiguous(), bounds


def solution_check(dmesh, lp, vals, rhs, x, nx, ny):
    """Parity evidence carried by every multi-GPU bench line, independent of the transport the solver used:
    the TRUE residual ||rhs - A x|| / ||rhs|| recomputed with the ghost values of x fetched over NCCL
    (torch.distributed all_gather of the owned blocks), and fingerprints of x that must agree between runs
    at different GPU counts: sum, sum of squares and the tip displacement u_y(node (nx, ny))."""
    import torch.distributed as dist
    dm = dmesh.dm
    dev = x.device
    world = dist.get_world_size()
    n_own = torch.tensor([x.numel()], dtype=torch.int64, device=dev)
    sizes = [torch.zeros_like(n_own) for _ in range(world)]
    dist.all_gather(sizes, n_own)
    sizes = [int(t.item()) for t in sizes]
    parts = [torch.empty(sz, dtype=torch.float64, device=dev) for sz in sizes]
    dist.all_gather(parts, x.contiguous())                 # NCCL
    xg = torch.cat(parts)                                  # global solution in global DOF order (row blocks)
    gd = (lp.node_gid[:, None] * lp.dim + torch.arange(lp.dim, device=dev)[None, :]).reshape(-1)
    x_local = xg[gd].contiguous()                          # owned values, then ghost values
    res = rhs - dm.spmv(vals, x_local)
    acc = torch.stack([torch.dot(res, res), torch.dot(rhs, rhs)])
    dist.all_reduce(acc)                                   # NCCL
    tip = 2 * ((ny + 1) * (nx + 1) - 1) + 1
    return {"true_relres_nccl": float(torch.sqrt(acc[0] / acc[1])), "x_sum": float(xg.sum()),
            "x_sumsq": float(torch.dot(xg, xg)), "tip_uy": float(xg[tip])}


def bench_distributed(args, metric, mat, measured_peak_hbm, ClockSampler, asm_bytes, pcg_bytes_per_iter):
    """N > 1 arm of bench.py: the S16M mesh split into N row blocks (strong scaling)."""
    import json
    import torch.distributed as dist
    from .device import Context, KIND_ELAST_PSTRESS
    rank, world = dist.get_rank(), dist.get_world_size()
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    nx, ny = args.nx, args.ny
    ctx = Context.get(local_rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    # communicator set-up (ncclCommInitRank, IPC handle exchange) is timed apart from the symbolic phase
    t0 = time.perf_counter()
    init_nccl(ctx)
    torch.cuda.synchronize()
    comm_init_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    lp, coords_local, bounds = structured_rank_problem(nx, ny, rank, world, dev)
    torch.cuda.synchronize()
    partition_ms = 1e3 * (time.perf_counter() - t0)
    t0 = time.perf_counter()
    dmesh = DistributedMesh(coords_local, lp, None, device=local_rank, ctx=ctx)
    dm = dmesh.dm
    dm.csr_pattern()
    torch.cuda.synchronize()
    plan_ms = 1e3 * (time.perf_counter() - t0) - dmesh.peer_setup_ms
    comm_init_ms += dmesh.peer_setup_ms

    n_el_total, n_nodes_total = 2 * nx * ny, (nx + 1) * (ny + 1)
    n_total = 2 * n_nodes_total
    # clamp the left edge (i = 0), load the right edge (i = nx): local DOF lists incl. ghosts
    lines = torch.arange(ny + 1, device=dev) * (nx + 1)
    bc_g = torch.stack([2 * lines, 2 * lines + 1], dim=1).reshape(-1)
    bc, _ = localize_dofs(lp, bc_g)
    bc_val = torch.zeros(bc.numel(), dtype=torch.float64, device=dev)
    ld, keep = localize_dofs(lp, 2 * (lines + nx) + 1)
    f = torch.zeros(dm.n_rows, dtype=torch.float64, device=dev)
    ld = ld[ld < dm.n_rows].long()
    f[ld] = -1000.0 / ny
    mat_dev = torch.as_tensor(mat).to(dev)
    vals = torch.empty(dm.nnz, dtype=torch.float64, device=dev)
    rhs, x, work = torch.empty_like(f), torch.zeros_like(f), dm.pcg_workspace()

    def step(timers=None):
        a0, a1, p0, p1 = ev(), ev(), ev(), ev()
        # the previous step ended with a host synchronisation: keep the GPU busy for ~0.3 ms so the
        # launches below are queued before it gets to them (the events then bracket device time,
        # not the CPU's launch latency)
        _gpu_spin(torch)
        a0.record()
        dm.assemble(KIND_ELAST_PSTRESS, mat_dev, out=vals, variant=args.variant)
        a1.record()
        rhs.copy_(f)
        dm.dirichlet(vals, rhs, bc, bc_val)
        x.zero_()
        p0.record()
        dmesh.pcg(vals, rhs, x=x, fixed_iters=args.pcg_iters, work=work)
        p1.record()
        if timers is not None:
            timers.append((a0, a1, p0, p1))

    warm = max(args.warmup, 3)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    dist.barrier()
    launches0 = ctx.launches
    timers = []
    s0, s1 = ev(), ev()
    torch.cuda.synchronize()
    tw0 = time.perf_counter()
    s0.record()
    for _ in range(args.steps):
        step(timers)
    s1.record()
    torch.cuda.synchronize()
    tw1 = time.perf_counter()
    dist.barrier()
    launches = ctx.launches - launches0
    clocks = sampler.stop(tw0, tw1) if rank == 0 else None
    loc = torch.tensor([np.mean([a0.elapsed_time(a1) for a0, a1, _, _ in timers]),
                        np.mean([p0.elapsed_time(p1) for _, _, p0, p1 in timers]),
                        s0.elapsed_time(s1) / args.steps, float(dm.nnz), float(launches)],
                       dtype=torch.float64, device=dev)
    mx = loc.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    sm = loc.clone()
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    t_asm, t_pcg, ms_per_step = mx[0].item() * 1e-3, mx[1].item() * 1e-3, mx[2].item()
    nnz_total = sm[3].item()

    # end to end with host buffers: H2D local coords -> assemble -> D2H local vals
    h_coords = dm.coords.cpu().pin_memory()
    h_vals = torch.empty(dm.nnz, dtype=torch.float64).pin_memory()
    h_x = torch.empty(dm.n_rows, dtype=torch.float64).pin_memory()
    h_rhs = f.cpu().pin_memory()
    e2e_a, e2e_p = [], []
    for it in range(3):
        torch.cuda.synchronize()
        dist.barrier()
        t0 = time.perf_counter()
        dm.coords.copy_(h_coords, non_blocking=True)
        dm.assemble(KIND_ELAST_PSTRESS, mat_dev, out=vals, variant=args.variant)
        h_vals.copy_(vals, non_blocking=True)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        rhs.copy_(h_rhs, non_blocking=True)
        dm.dirichlet(vals, rhs, bc, bc_val)
        x.zero_()
        dmesh.pcg(vals, rhs, x=x, fixed_iters=args.pcg_iters, work=work)
        h_x.copy_(x, non_blocking=True)
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if it > 0:
            e2e_a.append(t1 - t0)
            e2e_p.append(t2 - t1)
    e2 = torch.tensor([np.mean(e2e_a), np.mean(e2e_p)], dtype=torch.float64, device=dev)
    dist.all_reduce(e2, op=dist.ReduceOp.MAX)

    solve = None
    if args.full_solve:
        dm.assemble(KIND_ELAST_PSTRESS, mat_dev, out=vals, variant=args.variant)
        rhs.copy_(f)
        dm.dirichlet(vals, rhs, bc, bc_val)
        x.zero_()
        q0, q1 = ev(), ev()
        q0.record()
        _, iters, relres = dmesh.pcg(vals, rhs, x=x, rtol=1e-8, work=work)
        q1.record()
        torch.cuda.synchronize()
        ts = torch.tensor([q0.elapsed_time(q1) * 1e-3], dtype=torch.float64, device=dev)
        dist.all_reduce(ts, op=dist.ReduceOp.MAX)
        solve = {"rtol": 1e-8, "iters": iters, "relres": relres, "seconds": ts.item(),
                 "dof_iters_per_s": n_total * iters / ts.item()}
        solve.update(solution_check(dmesh, lp, vals, rhs, x, nx, ny))

    if rank == 0:
        peak, peak_kind = measured_peak_hbm()
        a_bytes = asm_bytes(n_el_total, n_nodes_total, nnz_total)
        p_bytes = pcg_bytes_per_iter(n_total, nnz_total)
        asm_gbs = a_bytes / t_asm / 1e9
        pcg_gbs = p_bytes * args.pcg_iters / t_pcg / 1e9
        line = {
            "metric": metric, "value": n_el_total / t_asm / 1e6, "unit": "Melem/s", "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"S16M-family structured plane-stress mesh {nx}x{ny} cells "
                                   f"({n_el_total} triangles, {n_total} DOF) in {world} row blocks; step = numeric "
                                   f"assembly (no communication) + Dirichlet + {args.pcg_iters} Jacobi-PCG iterations "
                                   "(interface halo + dot-product all-reduce per iteration)",
                       "nx": nx, "ny": ny, "pcg_iters_per_step": args.pcg_iters, "l2": "inputs_larger_than_l2",
                       "parallelism": f"row-block x{world}", "pattern_build_ms": plan_ms,
                       "comm_init_ms": comm_init_ms, "partition_ms": partition_ms,
                       "pcg_kernel": "k_pcg_persist<MULTI> (one cooperative launch, single-reduction CG)"
                       if (world >= 4 or os.environ.get("FE_B200_PERSIST") == "1") and dmesh._dst_off is not None
                       and os.environ.get("FE_B200_PERSIST") != "0" and not os.environ.get("FE_B200_NO_PERSIST")
                       else "k_spmv_stream + k_pcg_update + k_pcg_pupdate",
                       "transport": "peer-memory halo stores + all-reduce fused into the PCG kernels (NVLink)"
                       if dmesh._dst_off is not None else "NCCL send/recv + all-reduce"},
            "assembly": {"ms": 1e3 * t_asm, "melem_per_s": n_el_total / t_asm / 1e6, "algorithmic_bytes": a_bytes},
            "pcg": {"dof_iters_per_s": n_total * args.pcg_iters / t_pcg, "ms_per_iter": 1e3 * t_pcg / args.pcg_iters,
                    "iters": args.pcg_iters,
                    "roofline": {"bound": "hbm", "achieved": pcg_gbs, "peak": peak * world, "unit": "GB/s",
                                 "frac": pcg_gbs / (peak * world), "traffic": None, "peak_kind": peak_kind}},
            "roofline": {"bound": "hbm", "achieved": asm_gbs, "peak": peak * world, "unit": "GB/s",
                         "frac": asm_gbs / (peak * world), "traffic": None, "peak_kind": peak_kind,
                         "note": "aggregate over all ranks"},
            "e2e": {"value": n_el_total / e2[0].item() / 1e6, "unit": "Melem/s",
                    "h2d_bytes_per_step": int((h_coords.numel() + h_rhs.numel()) * 8),
                    "d2h_bytes_per_step": int((h_vals.numel() + h_x.numel()) * 8),
                    "pcg_dof_iters_per_s": n_total * args.pcg_iters / e2[1].item(), "note": "bytes are per rank"},
            "gpu_launches": int(sm[4].item()), "clocks": clocks, "solve": solve,
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()
