"""CPU, world_size 2 and 3 over gloo: the N>1 host logic (row-block partition, ghost layer,
halo send/receive lists derived without negotiation, localisation of condition DOFs).

Each rank builds its LocalProblem exactly as the GPU path does (finite_elements_b200/dist.py
on CPU tensors), assembles its rows with the numpy oracle, and runs a Jacobi-PCG whose only
communication is the halo exchange described by (nbr_rank, send_ptr, send_idx, recv_ptr) and
a scalar all-reduce -- the same schedule fe_dist_pcg executes with NCCL.  The result must
equal the single-domain solve.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _halo_exchange(lp, vec, n_rows):
    reqs, bufs = [], []
    for k, r in enumerate(lp.nbr_rank.tolist()):
        send = torch.as_tensor(vec[lp.send_idx.numpy()[lp.send_ptr[k]:lp.send_ptr[k + 1]]].copy())
        recv = torch.empty(int(lp.recv_ptr[k + 1] - lp.recv_ptr[k]), dtype=torch.float64)
        reqs.append(dist.isend(send, dst=r))
        reqs.append(dist.irecv(recv, src=r))
        bufs.append((k, recv))
    for q in reqs:
        q.wait()
    for k, recv in bufs:
        vec[n_rows + lp.recv_ptr[k]:n_rows + lp.recv_ptr[k + 1]] = recv.numpy()


def _allsum(*vals):
    t = torch.tensor(vals, dtype=torch.float64)
    dist.all_reduce(t)
    return t.tolist()


def _worker(rank, world, port, nx, ny, kind_name, shuffle, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from finite_elements_b200.dist import partition_bounds, local_problem, localize_dofs
        from oracle import numpy_oracle as no
        tet = kind_name == "tet"
        if tet:   # slab partition of a Kuhn box, nx x ny x 2 cells (node planes are contiguous in k)
            coords, conn = no.structured_tet_mesh(nx, ny, 2, h=0.5, jitter=0.2, seed=2)
        else:
            coords, conn = no.structured_mesh(nx, ny, jitter=0.2, seed=2)
        if shuffle:  # general numbering: ranks get more than two neighbours' worth of ghosts
            rng = np.random.default_rng(9)
            conn = conn[rng.permutation(len(conn))]
        kind = no.KIND_MAGNETIC if kind_name == "mag" else (no.KIND_ELAST_TET if tet else no.KIND_ELAST_PSTRESS)
        dim = no.kind_dim(kind)
        mat = np.array([[4e-7 * np.pi, 0, 0, 0]]) if kind_name == "mag" else np.array([[210e9, 0.25, 1.0, 7860]])
        n_nodes = len(coords)
        bounds = partition_bounds(n_nodes, world, align=1 if shuffle else nx + 1)
        lo, hi = bounds[rank], bounds[rank + 1]
        lp = local_problem(torch.as_tensor(conn), lo, hi, bounds, dim)
        n_rows, n_cols = lp.n_owned * dim, lp.n_local * dim
        gid = lp.node_gid.numpy()
        assert np.array_equal(gid[:lp.n_owned], np.arange(lo, hi)) and np.all(np.diff(gid[lp.n_owned:]) > 0)
        cl = lp.conn_local.numpy()
        assert np.array_equal(gid[cl], conn[lp.elem_sel.numpy()])
        # local rows from the oracle (owned rows are complete: all incident elements are present)
        k_loc = no.assemble_k(kind, coords[gid], cl, np.zeros(len(cl), np.int32), mat)[:n_rows]
        k_glob = no.assemble_k(kind, coords, conn, np.zeros(len(conn), np.int32), mat)
        gd = (gid[:, None] * dim + np.arange(dim)[None, :]).reshape(-1)   # local dof -> global dof
        assert abs(k_loc - k_glob[gd[:n_rows]][:, gd]).max() <= 1e-9 * abs(k_glob).max()
        if tet:   # the device path's symbolic phase on the owned/ghost layout = the oracle's local rows
            from finite_elements_b200.device import tet_symbolic, tet_csr
            cp, ce, ap, adj, deg = tet_symbolic(torch.as_tensor(cl).long(), lp.n_local, lp.n_owned)
            rowptr, colidx = tet_csr(ap, adj, deg)
            assert np.array_equal(rowptr.numpy(), k_loc.indptr) and np.array_equal(colidx.numpy(), k_loc.indices)
            for i in (0, lp.n_owned - 1):
                assert np.array_equal(ce.numpy()[cp[i]:cp[i + 1]], np.nonzero((cl == i).any(axis=1))[0])
        # problem: clamp i = 0, load i = nx
        lines = np.nonzero(coords[:, 0] == 0)[0] if tet else np.arange(ny + 1) * (nx + 1)
        bc_g = (lines[:, None] * dim + np.arange(dim)[None, :]).reshape(-1)
        f_g = np.zeros(n_nodes * dim)
        f_g[(lines + nx) * dim + dim - 1] = -1000.0 / ny    # (nodes with i = nx in both mesh families)
        bc_l, _ = localize_dofs(lp, bc_g)
        bc_l = bc_l.numpy()
        assert np.array_equal(np.sort(gd[bc_l]), np.sort(np.intersect1d(bc_g, gd)))
        # eliminate on the local rectangular block (what fe_dirichlet_apply does, g = 0)
        k_loc = k_loc.tolil()
        is_bc = np.zeros(n_cols, bool)
        is_bc[bc_l] = True
        k_loc = k_loc.tocsr()
        rows = np.repeat(np.arange(n_rows), np.diff(k_loc.indptr))
        kill = is_bc[rows] | is_bc[k_loc.indices]
        k_loc.data[kill] = 0.0
        k_loc.data[(rows == k_loc.indices) & is_bc[rows]] = 1.0
        b = f_g[gd[:n_rows]].copy()
        b[is_bc[:n_rows]] = 0.0
        # distributed Jacobi-PCG: halo exchange of p, all-reduced dots
        dinv = 1.0 / k_loc.diagonal()
        x = np.zeros(n_rows)
        r = b.copy()
        p = np.zeros(n_cols)
        p[:n_rows] = dinv * r
        rz, bb = _allsum(float(r @ (dinv * r)), float(b @ b))
        it = 0
        while it < 5000:
            _halo_exchange(lp, p, n_rows)
            q = k_loc @ p
            (pq,) = _allsum(float(p[:n_rows] @ q))
            alpha = rz / pq
            x += alpha * p[:n_rows]
            r -= alpha * q
            rz_new, rr = _allsum(float(r @ (dinv * r)), float(r @ r))
            it += 1
            if rr <= 1e-26 * bb:
                break
            p[:n_rows] = dinv * r + (rz_new / rz) * p[:n_rows]
            rz = rz_new
        # single-domain reference
        u = no.solve_reduced_direct(k_glob, f_g, bc_g, np.zeros(len(bc_g)), permc_spec='COLAMD')
        err = np.linalg.norm(x - u[gd[:n_rows]]) / np.linalg.norm(u)
        out.put((rank, float(err), it, len(lp.nbr_rank)))
    finally:
        dist.destroy_process_group()


def _run(world, nx, ny, kind_name, shuffle=False):
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, kind_name, shuffle, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    res = sorted(out.get(timeout=5) for _ in range(world))
    for rank, err, it, n_nbr in res:
        assert err <= 1e-8, (rank, err, it)
    return res


@pytest.mark.timeout(300)
def test_two_ranks_elasticity_row_blocks():
    res = _run(2, 12, 9, "stress")
    assert [r[3] for r in res] == [1, 1]


@pytest.mark.timeout(300)
def test_three_ranks_magnetic_row_blocks():
    res = _run(3, 10, 11, "mag")
    assert [r[3] for r in res] == [1, 2, 1]


@pytest.mark.timeout(300)
def test_two_ranks_tetrahedra_slabs():
    """3 DOF per node, 4 nodes per element: same partition / halo logic, plus the owned-row symbolic
    phase of DeviceMesh3D against the oracle's local rows."""
    res = _run(2, 4, 3, "tet")
    assert [r[3] for r in res] == [1, 1]


@pytest.mark.timeout(300)
def test_two_ranks_shuffled_elements_unaligned_bounds():
    _run(2, 9, 7, "stress", shuffle=True)


def test_partition_bounds():
    from finite_elements_b200.dist import partition_bounds
    assert partition_bounds(100, 4) == [0, 25, 50, 75, 100]
    b = partition_bounds(4097 * 2049, 8, align=4097)
    assert b[0] == 0 and b[-1] == 4097 * 2049 and all(x % 4097 == 0 for x in b)
    assert max(np.diff(b)) - min(np.diff(b)) <= 4097
    assert partition_bounds(5, 8)[-1] == 5 and all(np.diff(partition_bounds(5, 8)) >= 0)
