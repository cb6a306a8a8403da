"""Array-level device API: flat mesh arrays in, CSR / solution tensors out.

This is the layer `analysis.FiniteElementAnalysis` (the drop-in class surface) and
bench.py sit on.  torch tensors are used only as device buffers; every computation is
a call into libfe_b200.so through ctypes (finite_elements_b200/_lib.py).
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (check, lib, KIND_ELAST_PSTRESS, KIND_ELAST_PSTRAIN, KIND_MAGNETIC, KIND_MASS,  # noqa: F401
                   KIND_ELAST_TET, KIND_MASS_TET)


def kind_dim(kind):
    if kind in (KIND_ELAST_TET, KIND_MASS_TET):
        return 3
    return 1 if kind == KIND_MAGNETIC else 2


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Context:
    """One fe_ctx per (thread, device)."""

    _cache = {}

    def __init__(self, device=0):
        if not torch.cuda.is_available():
            raise RuntimeError("finite_elements_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", device)
        h = C.c_void_p()
        check(lib.fe_ctx_create(device, C.byref(h)))
        self.handle = h

    @classmethod
    def get(cls, device=0):
        """The calling thread's context for `device` (fe_b200.h: one ctx per (thread, device) -- a ctx
        owns scratch buffers and a cached CUDA graph that two threads must not share)."""
        import threading
        key = (threading.get_ident(), int(device))
        if key not in cls._cache:
            cls._cache[key] = cls(device)
        return cls._cache[key]

    @property
    def launches(self):
        return int(lib.fe_ctx_launch_count(self.handle))


class DeviceMesh:
    """Flat mesh on the device + the per-mesh plan (CSR pattern, corner->slot map).

    coords f64[N,2], conn i32[E,3], mat_id i32[E] or None.  `n_owned` < N selects the
    multi-GPU layout (owned nodes first, ghosts after; rows only for owned nodes)."""

    def __init__(self, coords, conn, mat_id=None, dim=2, device=0, n_owned=None, ctx=None):
        self.ctx = ctx or Context.get(device)
        dev = self.ctx.device
        self.coords = torch.as_tensor(np.ascontiguousarray(coords, dtype=np.float64)).to(dev) \
            if not torch.is_tensor(coords) else coords.to(dev, torch.float64).contiguous()
        self.conn = torch.as_tensor(np.ascontiguousarray(conn, dtype=np.int32)).to(dev) \
            if not torch.is_tensor(conn) else conn.to(dev, torch.int32).contiguous()
        if mat_id is None:
            self.mat_id = None
        else:
            self.mat_id = torch.as_tensor(np.ascontiguousarray(mat_id, dtype=np.int32)).to(dev) \
                if not torch.is_tensor(mat_id) else mat_id.to(dev, torch.int32).contiguous()
        self.n_nodes = int(self.coords.shape[0])
        self.n_elems = int(self.conn.shape[0])
        self.n_owned = self.n_nodes if n_owned is None else int(n_owned)
        self.dim = int(dim)
        if self.coords.ndim != 2 or self.coords.shape[1] != 2:
            raise ValueError("coords must have shape (N, 2)")
        if self.n_elems and (self.conn.ndim != 2 or self.conn.shape[1] != 3):
            raise ValueError("conn must have shape (E, 3)")
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.fe_plan_create(self.ctx.handle, _stream(), self.n_nodes, self.n_owned, self.n_elems, self.dim,
                                     _ptr(self.conn), _ptr(self.mat_id), C.byref(h)))
        self.plan = h
        self.nnz = int(lib.fe_plan_nnz(h))
        self.n_rows = int(lib.fe_plan_n_rows(h))
        self.n_cols = self.n_nodes * self.dim
        self.max_degree = int(lib.fe_plan_max_degree(h))
        self.plan_bytes = int(lib.fe_plan_bytes(h))
        self.fan_record_bytes = int(lib.fe_plan_fan_record_bytes(h))
        self._csr = None
        DeviceMesh._tokens += 1
        self._token = DeviceMesh._tokens   # identifies this mesh's immutable CSR pattern to the solver

    _tokens = 0

    @property
    def block_dim(self):
        """What the solve entry points may assume about the CSR (fe_b200.h): 2 = rows come in
        (2i, 2i+1) pairs sharing one column list of (2m, 2m+1) pairs; 3 = triples likewise; 1 = nothing."""
        return self.dim if self.dim in (1, 2, 3) else 1

    def _vouch_pattern(self):
        """fe_pcg_cache_pattern: the CSR tensors of csr_pattern() live as long as this object and
        are never written again, so the solver may keep what it derives from them."""
        rowptr, colidx = self.csr_pattern()
        check(lib.fe_pcg_cache_pattern(self.ctx.handle, _ptr(rowptr), _ptr(colidx), self._token))
        return rowptr, colidx

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                lib.fe_plan_destroy(self.plan)
                self.plan = None
        except Exception:  # interpreter shutdown
            pass

    # ---- pattern ---------------------------------------------------------------------
    def csr_pattern(self):
        """(rowptr i32[n_rows+1], colidx i32[nnz]) device tensors; canonical CSR."""
        if self._csr is None:
            dev = self.ctx.device
            rowptr = torch.empty(self.n_rows + 1, dtype=torch.int32, device=dev)
            colidx = torch.empty(max(self.nnz, 1), dtype=torch.int32, device=dev)[:self.nnz]
            with torch.cuda.device(dev):
                check(lib.fe_plan_csr(self.plan, _stream(), _ptr(rowptr), _ptr(colidx)))
            self._csr = (rowptr, colidx)
        return self._csr

    # ---- numeric ---------------------------------------------------------------------
    def _mat(self, mat):
        m = torch.as_tensor(np.ascontiguousarray(mat, dtype=np.float64)) if not torch.is_tensor(mat) else mat
        m = m.to(self.ctx.device, torch.float64).contiguous()
        if m.ndim != 2 or m.shape[1] != 4:
            raise ValueError("mat must have shape (G, 4)")
        return m

    def assemble(self, kind, mat, out=None, variant=0):
        """Global matrix values (CSR order of csr_pattern()).  fe_assemble."""
        m = self._mat(mat)
        if out is None:
            out = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=self.ctx.device)[:self.nnz]
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_assemble(self.ctx.handle, _stream(), self.plan, int(kind), _ptr(self.coords), _ptr(m),
                                  int(m.shape[0]), _ptr(out), int(variant)))
        return out

    def element_matrices(self, kind, mat):
        """Per-element matrices f64[E, (3 dim)^2] (row-major Ke).  fe_elem_matrices."""
        m = self._mat(mat)
        nd = 3 * kind_dim(kind)
        out = torch.empty((self.n_elems, nd * nd), dtype=torch.float64, device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_elem_matrices(self.ctx.handle, _stream(), int(kind), self.n_elems, _ptr(self.coords),
                                       _ptr(self.conn), _ptr(self.mat_id), _ptr(m), int(m.shape[0]), _ptr(out)))
        return out

    def element_post(self, kind, mat, u):
        """Per-element post-processing of the solution u.  fe_elem_post.
        elasticity: f64[E,7] = (eps_xx, eps_yy, gamma_xy, sig_xx, sig_yy, tau_xy, energy);
        magnetic: f64[E,2] = (B_x, B_y)."""
        m = self._mat(mat)
        u = (torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)) if not torch.is_tensor(u) else u)
        u = u.to(self.ctx.device, torch.float64).contiguous()
        if u.numel() < self.n_nodes * self.dim:
            raise ValueError("solution vector is shorter than n_nodes * dim")
        out = torch.empty((self.n_elems, 2 if kind == KIND_MAGNETIC else 7), dtype=torch.float64,
                          device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_elem_post(self.ctx.handle, _stream(), int(kind), self.n_elems, _ptr(self.coords),
                                   _ptr(self.conn), _ptr(self.mat_id), _ptr(m), int(m.shape[0]), _ptr(u), _ptr(out)))
        return out

    def source_factors(self, elem_sel=None):
        """(factors f64[n,3], area f64[n]) for the selected elements.  fe_source_factors."""
        dev = self.ctx.device
        if elem_sel is None:
            sel, n = None, self.n_elems
        else:
            sel = torch.as_tensor(np.ascontiguousarray(elem_sel, dtype=np.int32)).to(dev)
            n = int(sel.numel())
        fac = torch.empty((n, 3), dtype=torch.float64, device=dev)
        area = torch.empty(n, dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            check(lib.fe_source_factors(self.ctx.handle, _stream(), n, _ptr(sel), _ptr(self.coords), _ptr(self.conn),
                                        _ptr(fac), _ptr(area)))
        return fac, area

    def dirichlet(self, vals, rhs, bc_dof, bc_val):
        """In-place symmetric elimination on (vals, rhs).  fe_dirichlet_apply."""
        dev = self.ctx.device
        rowptr, colidx = self.csr_pattern()
        d = torch.as_tensor(np.ascontiguousarray(bc_dof, dtype=np.int32)).to(dev) if not torch.is_tensor(bc_dof) \
            else bc_dof.to(dev, torch.int32)
        v = torch.as_tensor(np.ascontiguousarray(bc_val, dtype=np.float64)).to(dev) if not torch.is_tensor(bc_val) \
            else bc_val.to(dev, torch.float64)
        with torch.cuda.device(dev):
            check(lib.fe_dirichlet_apply(self.ctx.handle, _stream(), self.n_rows, self.n_cols, _ptr(rowptr),
                                         _ptr(colidx), _ptr(vals), _ptr(rhs), int(d.numel()), _ptr(d), _ptr(v)))

    def scatter_add(self, rhs, dof, val):
        dev = self.ctx.device
        d = torch.as_tensor(np.ascontiguousarray(dof, dtype=np.int32)).to(dev)
        v = torch.as_tensor(np.ascontiguousarray(val, dtype=np.float64)).to(dev)
        with torch.cuda.device(dev):
            check(lib.fe_scatter_add(self.ctx.handle, _stream(), int(d.numel()), _ptr(d), _ptr(v), _ptr(rhs)))

    def spmv(self, vals, x, y=None):
        rowptr, colidx = self.csr_pattern()
        if y is None:
            y = torch.empty(self.n_rows, dtype=torch.float64, device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_spmv(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals),
                              _ptr(x), _ptr(y), self.block_dim))
        return y

    def spmm_pair(self, vals_a, vals_b, x, out_a=None, out_b=None):
        """(A X, B X) for a row-major (n_cols, m) block X, one pass over the shared pattern
        (vals_b None: A X only).  fe_spmm_pair."""
        rowptr, colidx = self.csr_pattern()
        if x.ndim != 2 or x.shape[0] != self.n_cols or not x.is_contiguous() or x.dtype != torch.float64:
            raise ValueError("spmm_pair: X must be a contiguous float64 (n_cols, m) tensor")
        m = int(x.shape[1])
        ya = out_a if out_a is not None else torch.empty(self.n_rows, m, dtype=torch.float64, device=x.device)
        yb = None
        if vals_b is not None:
            yb = out_b if out_b is not None else torch.empty(self.n_rows, m, dtype=torch.float64, device=x.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_spmm_pair(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals_a),
                                   _ptr(vals_b), _ptr(x), _ptr(ya), _ptr(yb), m, self.block_dim))
        return ya, yb

    def cheb_step(self, vals, dinv, d_in, d_out, r, z, c1, c2):
        """z += d_in; r -= A d_in; d_out = c1 d_in + c2 dinv r on (n, m) blocks, one pass.  fe_cheb_step."""
        rowptr, colidx = self.csr_pattern()
        for t in (d_in, d_out, r, z):
            if t.shape != d_in.shape or not t.is_contiguous() or t.dtype != torch.float64:
                raise ValueError("cheb_step: blocks must be contiguous float64 tensors of one shape")
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_cheb_step(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals),
                                   _ptr(dinv), _ptr(d_in), _ptr(d_out), _ptr(r), _ptr(z), float(c1), float(c2),
                                   int(d_in.shape[1]), self.block_dim))

    def csr_diagonal(self, vals):
        rowptr, colidx = self.csr_pattern()
        d = torch.empty(self.n_rows, dtype=torch.float64, device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_csr_diagonal(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals),
                                      _ptr(d)))
        return d

    def pcg_workspace(self):
        n = int(lib.fe_pcg_work_len(self.n_rows, self.n_cols))
        return torch.empty(n, dtype=torch.float64, device=self.ctx.device)

    def pcg(self, vals, b, x=None, rtol=1e-8, maxit=None, work=None, raise_on_maxit=True):
        """Jacobi-PCG on the (eliminated, SPD) system.  Returns (x, iters, relres)."""
        rowptr, colidx = self._vouch_pattern()
        if x is None:
            x = torch.zeros(self.n_rows, dtype=torch.float64, device=self.ctx.device)
        if work is None:
            work = self.pcg_workspace()
        if maxit is None:
            maxit = max(1000, 10 * self.n_rows)
        iters, relres = C.c_int32(0), C.c_double(0.0)
        with torch.cuda.device(self.ctx.device):
            rc = lib.fe_pcg(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals), _ptr(b),
                            _ptr(x), _ptr(work), self.block_dim, float(rtol), int(min(maxit, 2 ** 31 - 1)), C.byref(iters),
                            C.byref(relres))
        if rc == _lib.FE_ERR_NOT_CONVERGED and not raise_on_maxit:
            return x, iters.value, relres.value
        check(rc)
        return x, iters.value, relres.value

    def pcg_fixed(self, vals, b, x, iters, work=None):
        """Exactly `iters` PCG iterations, no convergence test (throughput runs)."""
        rowptr, colidx = self._vouch_pattern()
        if work is None:
            work = self.pcg_workspace()
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_pcg_fixed(self.ctx.handle, _stream(), self.n_rows, _ptr(rowptr), _ptr(colidx), _ptr(vals),
                                   _ptr(b), _ptr(x), _ptr(work), self.block_dim, int(iters)))
        return x

    # ---- convenience ------------------------------------------------------------------
    def to_scipy(self, vals):
        import scipy.sparse as sp
        rowptr, colidx = self.csr_pattern()
        return sp.csr_matrix((vals.cpu().numpy(), colidx.cpu().numpy(), rowptr.cpu().numpy()),
                             shape=(self.n_rows, self.n_cols))


def element_post_arrays(kind, coords, conn, mat_id, mat, u, device=0, ctx=None):
    """fe_elem_post / fe_tet_elem_post on flat arrays, WITHOUT a plan: post-processing needs the
    connectivity and the solution only, not the CSR pattern (results.py never assembles).
    Returns f64[E,7] (plane elasticity), f64[E,2] (magnetic) or f64[E,13] (tetrahedra) on the device."""
    ctx = ctx or Context.get(device)
    dev = ctx.device

    def up(a, dt):
        t = torch.as_tensor(np.ascontiguousarray(a, dtype=dt)) if not torch.is_tensor(a) else a
        return t.to(dev).contiguous()

    coords, conn = up(coords, np.float64), up(conn, np.int32)
    mat_id = None if mat_id is None else up(mat_id, np.int32)
    mat, u = up(mat, np.float64), up(u, np.float64)
    n_el, npe = int(conn.shape[0]), (int(conn.shape[1]) if conn.ndim == 2 else 0)
    sdim = kind_dim(kind)
    if coords.ndim != 2 or (n_el and npe != (4 if sdim == 3 else 3)) or mat.ndim != 2 or mat.shape[1] != 4:
        raise ValueError("element_post_arrays: inconsistent array shapes")
    if u.numel() < coords.shape[0] * sdim:
        raise ValueError("solution vector is shorter than n_nodes * dim")
    width = 13 if sdim == 3 else (2 if kind == KIND_MAGNETIC else 7)
    out = torch.empty((n_el, width), dtype=torch.float64, device=dev)
    with torch.cuda.device(dev):
        if sdim == 3:
            check(lib.fe_tet_elem_post(ctx.handle, _stream(), n_el, _ptr(coords), _ptr(conn), _ptr(mat_id), _ptr(mat),
                                       int(mat.shape[0]), _ptr(u), _ptr(out)))
        else:
            check(lib.fe_elem_post(ctx.handle, _stream(), int(kind), n_el, _ptr(coords), _ptr(conn), _ptr(mat_id),
                                   _ptr(mat), int(mat.shape[0]), _ptr(u), _ptr(out)))
    return out


def tet_symbolic(conn64, n, n_owned=None):
    """Symbolic phase of the tetrahedral path on conn64 (E,4) int64, any torch device.
    Returns (corner_ptr i64[n_owned+1], corner_elem i32, adj_ptr i64[n_owned+1], adj i64[sum deg],
    deg i64[n_owned]): per-node element lists in ascending element order, and the sorted node adjacency
    (diagonal included) -- for the owned nodes [0, n_owned) only (multi-GPU layout: owned nodes first,
    ghosts after; rows exist for owned nodes, columns for all n local nodes)."""
    dev = conn64.device
    n_owned = n if n_owned is None else n_owned
    e = conn64.shape[0]
    corner_node = conn64.reshape(-1)
    order = torch.sort(corner_node, stable=True).indices
    counts = torch.bincount(corner_node, minlength=n)[:n_owned]
    corner_ptr = torch.zeros(n_owned + 1, dtype=torch.int64, device=dev)
    corner_ptr[1:] = torch.cumsum(counts, 0)
    corner_elem = (order[:int(corner_ptr[-1].item())] // 4).to(torch.int32).contiguous()   # owned nodes sort first
    a = conn64[:, :, None].expand(e, 4, 4).reshape(-1)
    b = conn64[:, None, :].expand(e, 4, 4).reshape(-1)
    keep = a < n_owned
    diag = torch.arange(n_owned, device=dev, dtype=torch.int64)
    keys = torch.unique(torch.cat([a[keep] * n + b[keep], diag * n + diag]))      # sorted
    deg = torch.bincount(keys // n, minlength=n_owned)
    adj_ptr = torch.zeros(n_owned + 1, dtype=torch.int64, device=dev)
    adj_ptr[1:] = torch.cumsum(deg, 0)
    return corner_ptr, corner_elem, adj_ptr, keys % n, deg


def tet_csr(adj_ptr, adj, deg):
    """(rowptr i32[3n+1], colidx i32[nnz]) of the 3x3-block expansion: row 3i+r starts at
    9 adj_ptr[i] + r * 3 deg_i and holds columns 3 adj[k] + c, k-major -- sorted, scipy-canonical."""
    dev = adj.device
    n = deg.numel()
    rowptr = torch.empty(3 * n + 1, dtype=torch.int64, device=dev)
    base = 9 * adj_ptr[:-1]
    for r in range(3):
        rowptr[r:3 * n:3] = base + r * 3 * deg
    rowptr[3 * n] = 9 * adj_ptr[-1]
    node_of_block = torch.repeat_interleave(torch.arange(n, device=dev), deg)       # per adjacency entry
    k_in_node = torch.arange(adj.numel(), device=dev) - adj_ptr[:-1][node_of_block]
    colidx = torch.empty(int(9 * adj_ptr[-1].item()), dtype=torch.int32, device=dev)
    for r in range(3):
        pos = base[node_of_block] + r * 3 * deg[node_of_block] + 3 * k_in_node
        for c in range(3):
            colidx[pos + c] = (3 * adj + c).to(torch.int32)
    return rowptr.to(torch.int32).contiguous(), colidx


class DeviceMesh3D(DeviceMesh):
    """Linear tetrahedra, 3 DOF per node: coords f64[N,3], conn i32[E,4].

    The symbolic phase (per-node element lists, node adjacency, per-block element lists, the scipy-canonical
    CSR of the 3x3 blocks) is fe_tet_plan_create -- hand-written kernels like the triangles' fe_plan_create;
    the numeric phase is fe_tet_assemble / fe_tet_elem_matrices.  Everything CSR-level (Dirichlet
    elimination, SpMV, PCG, block products) is inherited unchanged."""

    def __init__(self, coords, conn, mat_id=None, device=0, ctx=None, n_owned=None, dim=3):
        if dim != 3:
            raise ValueError("DeviceMesh3D is the 3 DOF per node path")
        self.ctx = ctx or Context.get(device)
        dev = self.ctx.device
        self.coords = torch.as_tensor(np.ascontiguousarray(coords, dtype=np.float64)).to(dev) \
            if not torch.is_tensor(coords) else coords.to(dev, torch.float64).contiguous()
        self.conn = torch.as_tensor(np.ascontiguousarray(conn, dtype=np.int32)).to(dev) \
            if not torch.is_tensor(conn) else conn.to(dev, torch.int32).contiguous()
        if mat_id is None:
            self.mat_id = None
        else:
            self.mat_id = torch.as_tensor(np.ascontiguousarray(mat_id, dtype=np.int32)).to(dev) \
                if not torch.is_tensor(mat_id) else mat_id.to(dev, torch.int32).contiguous()
        if self.coords.ndim != 2 or self.coords.shape[1] != 3:
            raise ValueError("coords must have shape (N, 3)")
        self.n_nodes = int(self.coords.shape[0])
        self.n_elems = int(self.conn.shape[0])
        if self.n_elems and (self.conn.ndim != 2 or self.conn.shape[1] != 4):
            raise ValueError("conn must have shape (E, 4)")
        self.n_owned = self.n_nodes if n_owned is None else int(n_owned)   # multi-GPU: owned first, ghosts after
        self.dim = 3
        h = C.c_void_p()
        with torch.cuda.device(dev):
            check(lib.fe_tet_plan_create(self.ctx.handle, _stream(), self.n_nodes, self.n_owned, self.n_elems,
                                         _ptr(self.conn), C.byref(h)))
        self.plan = h
        self.nnz = int(lib.fe_plan_nnz(h))
        self.n_rows = int(lib.fe_plan_n_rows(h))
        self.n_cols = 3 * self.n_nodes
        self.max_degree = int(lib.fe_plan_max_degree(h))
        self.plan_bytes = int(lib.fe_plan_bytes(h))
        self._csr = None
        DeviceMesh._tokens += 1
        self._token = DeviceMesh._tokens

    def assemble(self, kind, mat, out=None, variant=0):
        """Global K (KIND_ELAST_TET) or M (KIND_MASS_TET) values in csr_pattern() order.  fe_tet_assemble."""
        m = self._mat(mat)
        if out is None:
            out = torch.empty(max(self.nnz, 1), dtype=torch.float64, device=self.ctx.device)[:self.nnz]
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_tet_assemble(self.ctx.handle, _stream(), self.plan, int(kind), _ptr(self.coords),
                                      _ptr(self.conn), _ptr(self.mat_id), _ptr(m), int(m.shape[0]), _ptr(out),
                                      int(variant)))
        return out

    def element_matrices(self, kind, mat):
        """Per-element matrices f64[E, 144] (row-major 12x12).  fe_tet_elem_matrices."""
        m = self._mat(mat)
        out = torch.empty((self.n_elems, 144), dtype=torch.float64, device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_tet_elem_matrices(self.ctx.handle, _stream(), int(kind), self.n_elems, _ptr(self.coords),
                                           _ptr(self.conn), _ptr(self.mat_id), _ptr(m), int(m.shape[0]), _ptr(out)))
        return out

    def element_post(self, kind, mat, u):
        """f64[E,13] = (6 strains, 6 stresses, energy) of the solution u.  fe_tet_elem_post."""
        m = self._mat(mat)
        u = (torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)) if not torch.is_tensor(u) else u)
        u = u.to(self.ctx.device, torch.float64).contiguous()
        if u.numel() < self.n_nodes * 3:
            raise ValueError("solution vector is shorter than n_nodes * dim")
        out = torch.empty((self.n_elems, 13), dtype=torch.float64, device=self.ctx.device)
        with torch.cuda.device(self.ctx.device):
            check(lib.fe_tet_elem_post(self.ctx.handle, _stream(), self.n_elems, _ptr(self.coords), _ptr(self.conn),
                                       _ptr(self.mat_id), _ptr(m), int(m.shape[0]), _ptr(u), _ptr(out)))
        return out

    def source_factors(self, elem_sel=None):
        raise NotImplementedError("the reference defines no element_to_node_factors for tetrahedra")


def solve_dirichlet_system(dm, kind, mat, load_dof, load_val, bc_dof, bc_val, rtol=1e-12, maxit=None,
                           variant=0, with_multipliers=True):
    """assemble -> rhs -> eliminate -> PCG -> (u, lambda, iters, relres) on the device.

    lambda = f_c - (K u)_c reproduces the Lagrange-multiplier tail of the reference's
    augmented solve (analysis.py:272-277, :539-541)."""
    dev = dm.ctx.device
    vals = dm.assemble(kind, mat, variant=variant)
    f = torch.zeros(dm.n_rows, dtype=torch.float64, device=dev)
    if len(load_dof):
        dm.scatter_add(f, load_dof, load_val)
    rhs = f.clone()
    dm.dirichlet(vals, rhs, bc_dof, bc_val)
    u, iters, relres = dm.pcg(vals, rhs, rtol=rtol, maxit=maxit)
    lam = None
    if with_multipliers:
        dm.assemble(kind, mat, out=vals, variant=variant)  # K again (the eliminated copy was overwritten in place)
        ku = dm.spmv(vals, u)
        idx = torch.as_tensor(np.asarray(bc_dof, dtype=np.int64)).to(dev)
        lam = (f - ku)[idx]
    return u, lam, iters, relres
