"""Modal analysis  K x = lambda M x  on the device (BASELINE configs[4]).

Replaces `FiniteElementAnalysis.modal_analysis` (analysis.py:741-796).  The reference builds K and
M as scipy matrices (m_matrix_sparse / k_matrix_sparse, :773-774) and calls ARPACK
(`eigsh(A=K, M=M, which='LM', k=k)`, :779-782) for the largest eigenvalues; its 'smallest' branch
inverts the DENSE K (:790), which cannot scale and is singular for the free-free beams of
scripts/ModalAnalysis.  Here both ends of the spectrum come from one block method that only needs
block products with K and M:

  LOBPCG (Knyazev 2001): Rayleigh-Ritz on span[X, W, P] with X the current block, W the
  preconditioned residuals and P the previous search directions.  Per iteration: ONE pass over the
  sparsity pattern K and M share (fe_spmm_pair: K W and M W together), a Jacobi or
  Chebyshev-polynomial preconditioner built from further block products, and (3m x 3m) dense work.

Everything O(n) lives in torch tensors on the device; the matrices never leave HBM.  The dense
(3m x 3m) Rayleigh-Ritz problems are solved on the host (scipy.linalg.eigh, microseconds).
"""
import numpy as np
import torch


class ModalInfo(dict):
    """iterations, residual norms, block size ... of the last call (attribute access for convenience)."""
    __getattr__ = dict.get


def _sym(g):
    return 0.5 * (g + g.T)


def _b_orthonormalize(v, bv, av=None):
    """V <- V T with T = D L^-T, (D V^T B V D) = L L^T, so that V^T B V = I.  Returns None when the
    block is numerically rank deficient (caller drops it)."""
    g = _sym(v.T @ bv)
    d = torch.diagonal(g)
    if not bool((d > 0).all()):
        return None
    d = d.rsqrt()
    g = d[:, None] * g * d[None, :]
    chol, info = torch.linalg.cholesky_ex(g)
    if int(info) != 0 or not bool(torch.isfinite(chol).all()):
        return None
    # T = D L^-T  <=>  T^T = L^-1 D
    t = torch.linalg.solve_triangular(chol, torch.diag(d), upper=False).T.contiguous()
    return v @ t, bv @ t, (av @ t if av is not None else None)


def chebyshev_preconditioner(apply_a, dinv, lmax, degree, ratio=30.0, fused_step=None):
    """T r ~ A^-1 r: `degree` steps of the Chebyshev iteration for D^-1 A on [lmax/ratio, lmax]
    (Saad, Iterative Methods, alg. 12.1), started from zero.  A fixed polynomial in A, hence a
    symmetric positive definite operator as LOBPCG requires; each step is one block product.
    fused_step(d_in, d_out, r, z, c1, c2) -- DeviceMesh.cheb_step -- does a whole step in one kernel."""
    lmin = lmax / ratio
    theta, delta = 0.5 * (lmax + lmin), 0.5 * (lmax - lmin)
    sigma1 = theta / delta

    def apply(r):
        r = r.clone() if fused_step is None else r.contiguous().clone()
        rho = 1.0 / sigma1
        d = (dinv[:, None] * r) / theta
        z = torch.zeros_like(r)
        d_next = torch.empty_like(d) if fused_step is not None else None
        for i in range(degree - 1):
            rho_new = 1.0 / (2.0 * sigma1 - rho)
            c1, c2 = rho_new * rho, 2.0 * rho_new / delta
            if fused_step is not None:
                fused_step(d, d_next, r, z, c1, c2)
                d, d_next = d_next, d
            else:
                z += d
                r -= apply_a(d)
                d = c1 * d + c2 * (dinv[:, None] * r)
            rho = rho_new
        z += d
        return z
    return apply


def gershgorin_lmax(abs_rowsum, dinv):
    """Upper bound of the spectrum of D^-1 A: max_i sum_j |a_ij| / a_ii.  The Chebyshev polynomial is
    only positive definite if the whole spectrum lies below its upper end, so a guaranteed bound is
    used rather than a power-iteration estimate (which approaches lambda_max from below)."""
    return float((abs_rowsum * dinv).max())


def _orthonormalize_against(w, basis, b_basis, op):
    """W <- B-orthonormal block spanning W minus its components along `basis` (itself B-orthonormal):
    two rounds of block Gram-Schmidt + Cholesky QR (the second round removes what the first one's
    rounding left).  Returns (W, A W, B W) with the products computed from the final W, or None."""
    aw = bw = None
    for _ in range(2):
        w = w - basis @ (b_basis.T @ w)
        aw, bw = op(w)
        out = _b_orthonormalize(w, bw, aw)
        if out is None:
            return None
        w, bw, aw = out
    return w, aw, bw


def lobpcg(apply_pair, n, k, device, *, largest=False, precond=None, mask=None, tol=1e-9, maxit=5000,
           guard=None, seed=0, anorm=1.0, refresh=30, floor=2e-13):
    """k extreme eigenpairs of A x = lambda B x (A symmetric, B symmetric positive definite).

    apply_pair(V) -> (A V, B V) for a (n, j) block.  precond(R) -> T R with T ~ A^-1 (smallest) or
    ~ B^-1 (largest); None = identity.  mask: optional (n,) 0/1 vector restricting the problem to the
    DOFs with mask 1 (Dirichlet-constrained modal analysis).  Returns (eigenvalues ascending,
    (n, k) B-orthonormal eigenvectors, ModalInfo).

    The basis S = [X, W, P] is kept B-orthonormal (W is orthogonalised against X and P in the big
    space; the new P comes from an orthogonalisation of the Ritz coefficients in the small space --
    Hetmaniuk & Lehoucq 2006), so the Rayleigh-Ritz problem is a standard symmetric one and every
    implicit update X <- S C is an orthogonal combination: nothing amplifies rounding.

    Convergence of pair i:  ||A x - lambda B x||_2 <= tol (||A x|| + |lambda| ||B x||) + floor * anorm * ||x||
    (the second term is the accuracy any method that multiplies by A can attain).
    """
    import scipy.linalg as sla
    if guard is None:
        guard = max(2, k // 4)
    m = k + guard
    if 3 * m > n // 2:
        raise ValueError(f"lobpcg: block of {m} vectors is too large for n = {n}; use the dense path")
    sign = -1.0 if largest else 1.0

    def op(v):
        av, bv = apply_pair(v.contiguous())
        if sign < 0:
            av = -av
        if mask is not None:
            av = av * mask[:, None]
            bv = bv * mask[:, None]
        return av, bv

    # optional section timers (FE_B200_MODAL_PROF=1: device-synchronised wall clock per section -> info["prof"])
    import os
    import time
    prof = {} if os.environ.get("FE_B200_MODAL_PROF") else None
    t_last = [0.0]

    def lap(name):
        if prof is not None:
            torch.cuda.synchronize()
            t = time.perf_counter()
            if name:
                prof[name] = prof.get(name, 0.0) + (t - t_last[0])
            t_last[0] = t

    gen = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(n, m, generator=gen, dtype=torch.float64).to(device)
    if mask is not None:
        x = x * mask[:, None]
    info = ModalInfo(iterations=0, block=m, restarts=0, products=0)

    def rayleigh_ritz_x(x):
        for _ in range(2):
            ax, bx = op(x)
            info["products"] += 1
            out = _b_orthonormalize(x, bx, ax)
            if out is None:
                raise np.linalg.LinAlgError("lobpcg: the block lost rank")
            x, bx, ax = out
        ax, bx = op(x)
        wv, c = sla.eigh(_sym(x.T @ ax).cpu().numpy())
        c = torch.as_tensor(c, device=device)
        return x @ c, ax @ c, bx @ c, torch.as_tensor(wv, device=device)

    x, ax, bx, theta = rayleigh_ritz_x(x)
    p = ap = bp = None
    rn = den = None
    for it in range(maxit):
        lap(None)
        r = ax - bx * theta[None, :]
        rn = torch.linalg.norm(r, dim=0)
        den = (tol * (torch.linalg.norm(ax, dim=0) + theta.abs() * torch.linalg.norm(bx, dim=0))
               + floor * anorm * torch.linalg.norm(x, dim=0))
        conv = rn <= den
        info["iterations"] = it
        if bool(conv[:k].all()):
            break
        act = torch.nonzero(~conv).flatten()           # soft locking: converged pairs get no new direction
        w = r[:, act]
        lap("residual+test")
        if precond is not None:
            w = precond(w)
        lap("preconditioner")
        if mask is not None:
            w = w * mask[:, None]
        basis, b_basis = (x, bx) if p is None else (torch.cat([x, p], 1), torch.cat([bx, bp], 1))
        out = _orthonormalize_against(w, basis, b_basis, op)
        lap("orthonormalise W")
        info["products"] += 2
        if out is None:                                 # residual block collapsed: refresh and retry without P
            x, ax, bx, theta = rayleigh_ritz_x(x)
            p = ap = bp = None
            info["restarts"] += 1
            if info["restarts"] > 20:
                break
            continue
        w, aw, bw = out
        if p is None:
            s, a_s, b_s = torch.cat([x, w], 1), torch.cat([ax, aw], 1), torch.cat([bx, bw], 1)
        else:
            s, a_s, b_s = torch.cat([x, w, p], 1), torch.cat([ax, aw, ap], 1), torch.cat([bx, bw, bp], 1)
        gram = _sym(s.T @ a_s)
        lap("gram")
        wv, c = sla.eigh(gram.cpu().numpy())  # S is B-orthonormal: standard problem
        c = c[:, :m]
        # new directions: the part of the Ritz vectors outside the old X, orthogonalised against the
        # new X in coefficient space (S orthonormal => B-inner products are Euclidean ones there)
        y = c.copy()
        y[:m] = 0.0
        y -= c @ (c.T @ y)
        y -= c @ (c.T @ y)
        q, rr, _ = sla.qr(y, mode="economic", pivoting=True)
        dg = np.abs(np.diag(rr))
        rank = int((dg > 1e-10 * max(dg[0], 1e-300)).sum()) if dg.size else 0
        c_t = torch.as_tensor(c, device=device)
        lap("ritz (host)")
        if rank > 0:
            q_t = torch.as_tensor(np.ascontiguousarray(q[:, :rank]), device=device)
            p, ap, bp = s @ q_t, a_s @ q_t, b_s @ q_t
        else:
            p = ap = bp = None
        x, ax, bx = s @ c_t, a_s @ c_t, b_s @ c_t
        theta = torch.as_tensor(wv[:m].copy(), device=device)
        lap("update X, P")
        if (it + 1) % refresh == 0:                     # recompute A X, B X from X: no drift
            x, ax, bx, theta = rayleigh_ritz_x(x)
            if p is not None:
                outp = _orthonormalize_against(p, x, bx, op)
                p, ap, bp = (None, None, None) if outp is None else outp
            lap("refresh")
    else:
        info["iterations"] = maxit
    if prof is not None:
        info["prof"] = {k_: round(v_, 4) for k_, v_ in prof.items()}
    lam = sign * theta[:k]
    vec = x[:, :k]
    info["residual_norms"] = (rn[:k] / den[:k]).cpu().numpy()   # <= 1 means converged
    info["converged"] = bool((rn[:k] <= den[:k]).all())
    if largest:                                          # eigsh(which='LM') returns ascending order
        lam = torch.flip(lam, dims=[0])
        vec = torch.flip(vec, dims=[1])
        info["residual_norms"] = info["residual_norms"][::-1].copy()
    return lam, vec, info


def dense_pencil_eigh(dm, k_vals, m_vals, k, largest, mask=None):
    """Tiny problems (n < 6 block widths): the whole pencil, on the device, by Cholesky reduction
    M = L L^T, C = L^-1 K L^-T, eigh(C) -- the same fallback scipy's lobpcg takes."""
    rowptr, colidx = dm.csr_pattern()
    n = dm.n_rows
    dev = k_vals.device
    rows = torch.repeat_interleave(torch.arange(n, device=dev), (rowptr[1:] - rowptr[:-1]).long())
    kd = torch.zeros(n, n, dtype=torch.float64, device=dev)
    md = torch.zeros(n, n, dtype=torch.float64, device=dev)
    kd.index_put_((rows, colidx.long()), k_vals, accumulate=True)
    md.index_put_((rows, colidx.long()), m_vals, accumulate=True)
    keep = torch.arange(n, device=dev) if mask is None else torch.nonzero(mask > 0).flatten()
    kd, md = kd[keep][:, keep], md[keep][:, keep]
    chol = torch.linalg.cholesky(_sym(md))
    c = torch.linalg.solve_triangular(chol, kd, upper=False)
    c = torch.linalg.solve_triangular(chol, c.T.contiguous(), upper=False)
    w, y = torch.linalg.eigh(_sym(c))
    v = torch.linalg.solve_triangular(chol.T.contiguous(), y, upper=True)
    sel = slice(len(w) - k, len(w)) if largest else slice(0, k)
    vec = torch.zeros(n, k, dtype=torch.float64, device=dev)
    vec[keep] = v[:, sel]
    return w[sel], vec


def modal_solve(dm, k_vals, m_vals, k, order, *, mask=None, tol=1e-9, maxit=5000, cheb_degree=None, cheb_ratio=None,
                guard=None, seed=0):
    """k 'largest' or 'smallest' eigenpairs of (K, M) given as value arrays on dm's CSR pattern.
    Returns (eigenvalues ascending (k,), eigenvectors (n, k) M-orthonormal, ModalInfo).

    cheb_degree: degree of the Chebyshev polynomial preconditioner of the 'smallest' branch (0 / 1 =
    plain Jacobi).  The outer iteration count of the lowest modes grows like 1/h with Jacobi and a
    degree-d polynomial divides it by about d/1.5, while a Chebyshev step (one fused kernel) is far cheaper
    than an outer iteration (orthogonalisations, Gram products, a host Rayleigh-Ritz): measured on the
    1 M-triangle mesh, degree 24 / 48 / 96 = 454 / 199 / 93 outer iterations = 5.3 / 3.4 / 2.6 s.  None
    therefore scales the degree with the mesh, d ~ 0.13 sqrt(nodes) in [24, 128] (96 at 1 M triangles),
    on [lmax / (0.45 d^2), lmax]; below 20 000 DOFs plain Jacobi is used."""
    if cheb_degree is None:
        nodes = dm.n_rows / max(1, getattr(dm, "dim", 2))
        cheb_degree = int(min(128, max(24, 8 * round(0.13 * nodes ** 0.5 / 8)))) if dm.n_rows > 20000 else 0
    if cheb_ratio is None:
        cheb_ratio = max(30.0, 0.45 * cheb_degree * cheb_degree)
    if order not in ("largest", "smallest"):
        raise ValueError("Order parameter should be either 'largest' or 'smallest'")  # analysis.py:796
    n, dev = dm.n_rows, k_vals.device
    largest = order == "largest"
    g = max(2, k // 4) if guard is None else guard
    n_free = n if mask is None else int(mask.sum().item())
    if k < 1 or k >= n_free:
        raise ValueError(f"modal analysis: k = {k} must be in [1, {n_free})")
    if 6 * (k + g) > n_free:
        lam, vec = dense_pencil_eigh(dm, k_vals, m_vals, k, largest, mask)
        return lam, vec, ModalInfo(iterations=0, block=0, dense=True, converged=True)

    def apply_pair(v):
        return dm.spmm_pair(k_vals, m_vals, v.contiguous())

    kdiag = dm.csr_diagonal(k_vals)
    mdiag = dm.csr_diagonal(m_vals)
    anorm = float(kdiag.abs().max())
    if largest:
        dinv = 1.0 / mdiag
        precond = lambda r: dinv[:, None] * r  # noqa: E731
    else:
        dinv = 1.0 / kdiag
        if cheb_degree and cheb_degree > 1:
            apply_k = lambda v: dm.spmm_pair(k_vals, None, v.contiguous())[0]  # noqa: E731
            ones = torch.ones(n, 1, dtype=torch.float64, device=dev)
            lmax = gershgorin_lmax(dm.spmm_pair(k_vals.abs(), None, ones)[0][:, 0], dinv)
            fused = lambda d, dn, r, z, c1, c2: dm.cheb_step(k_vals, dinv, d, dn, r, z, c1, c2)  # noqa: E731
            precond = chebyshev_preconditioner(apply_k, dinv, lmax, cheb_degree, cheb_ratio, fused_step=fused)
        else:
            precond = lambda r: dinv[:, None] * r  # noqa: E731
    lam, vec, info = lobpcg(apply_pair, n, k, dev, largest=largest, precond=precond, mask=mask, tol=tol,
                            maxit=maxit, guard=g, seed=seed, anorm=anorm)
    info["cheb_degree"] = 0 if largest else int(cheb_degree or 0)
    return lam, vec, info
