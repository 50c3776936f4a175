"""
Command-line flags owned by the reference's linsolver module (src/odil/linsolver.py:90-131; note it
also defines --lr and --nlvl) and the host sparse solvers used by Newton.  The Newton path
(SURVEY.md 8f-1): `solve` takes either a SciPy matrix (the reference's normal-equation path) or the
matrix-free `newton.StencilJacobian` that `Problem.linearize` returns (`--linsolver cg_b200`: CG on the
device; any other name: the SciPy solver on `matrix.tocsr()`).
"""
import numpy as np


def solve(matr, rhs, args, status=None, linsolver="direct"):
    """Solves min |M x - rhs| through the normal equations M^T M x = M^T rhs (linsolver.py:4-87)."""
    import scipy.sparse
    import scipy.sparse.linalg

    status = status if status is not None else dict()
    from .newton import StencilJacobian, cg_normal

    if hasattr(matr, "rmatvec") and hasattr(matr, "tocsr") and not scipy.sparse.issparse(matr):
        import torch

        if hasattr(rhs, "full"):  # Known (what Problem.linearize returns)
            rhs = rhs.full()

        if linsolver in ("cg_b200", "cg"):
            # matrix-free CG on the device; everything stays in HBM
            return cg_normal(matr, rhs, tol=getattr(args, "linsolver_tol", 1e-6),
                             maxiter=getattr(args, "linsolver_maxiter", None),
                             damp=getattr(args, "linsolver_damp", 0) or 0.0, status=status)
        # the reference's SciPy solvers on the assembled matrix (small grids)
        sol = solve(matr.tocsr(), rhs.detach().cpu().numpy().astype(np.float64), args, status, linsolver)
        return torch.as_tensor(np.asarray(sol), dtype=matr.dtype, device=matr.device)
    if getattr(args, "linsolver_maxiter", None) is None:
        args.linsolver_maxiter = 1000 if linsolver == "lsqr" else 50
    normal = matr.T.dot(matr).tocsr()
    if getattr(args, "linsolver_damp", 0):
        normal += args.linsolver_damp ** 2 * scipy.sparse.identity(matr.shape[1], format="csr")
    if getattr(args, "linsolver_dampdiag", 0):
        normal += args.linsolver_dampdiag ** 2 * scipy.sparse.diags(normal.diagonal())
    rhs_n = matr.T.dot(rhs)
    if linsolver == "direct":
        return scipy.sparse.linalg.spsolve(normal, rhs_n, permc_spec="MMD_ATA")
    if linsolver == "directsq":
        return scipy.sparse.linalg.spsolve(matr, rhs, permc_spec="MMD_ATA")
    if linsolver == "lsqr":
        res = scipy.sparse.linalg.lsqr(matr, rhs, damp=args.linsolver_damp, atol=args.linsolver_tol,
                                       btol=args.linsolver_tol, iter_lim=args.linsolver_maxiter)
        status.update(residual=res[7], anorm=res[5], acond=res[6], niter=res[2])
        return res[0]
    if linsolver == "bicgstab":
        residuals = []
        sol, _ = scipy.sparse.linalg.bicgstab(
            normal, rhs_n, rtol=0, atol=args.linsolver_tol, maxiter=args.linsolver_maxiter,
            callback=lambda x: residuals.append(np.mean((normal.dot(x) - rhs_n) ** 2) ** 0.5))
        status.update(residual=residuals[-1] if residuals else 0.0, niter=len(residuals))
        return sol
    raise ValueError("Unknown linsolver=" + linsolver)


# (flag, type, default, help) -- names, types and defaults of linsolver.py:90-131; `cg_b200` is the one addition
_FLAGS = [
    ("linsolver", str, "direct", "solver of the Newton step; cg_b200 = matrix-free conjugate gradients on the GPU"),
    ("linsolver_maxiter", int, None, "iteration limit of iterative solvers"),
    ("linsolver_tol", float, 1e-6, "tolerance of iterative solvers"),
    ("linsolver_damp", float, 0, "Tikhonov damping d: solves (J^T J + d^2 I) x = J^T r"),
    ("linsolver_dampdiag", float, 0, "as linsolver_damp, scaled by the diagonal of J^T J"),
    ("linsolver_verbose", int, 0, "print the solver's status every step"),
    ("linsolver_history", int, 0, "record the solver's status in train.csv"),
    ("lr", float, 1e-3, "learning rate of the gradient optimizers"),
    ("nlvl", int, 100, "upper bound on multigrid levels"),
    ("smooth_pre", int, 2, "algebraic multigrid: smoothing steps before coarsening"),
    ("smooth_post", int, 2, "algebraic multigrid: smoothing steps after prolongation"),
    ("omega", float, 0.6, "algebraic multigrid: relaxation factor of the Jacobi smoother"),
    ("ndirect", int, 3, "algebraic multigrid: grids up to this size are solved directly"),
    ("restriction", str, "full", "algebraic multigrid: restriction operator"),
]
_CHOICES = {
    "linsolver": ["multigrid", "direct", "directsq", "direct_cu", "sparseqr", "lsqr", "lsqr_cu", "bicgstab",
                  "cg_b200"],
    "restriction": ("full", "half", "injection"),
}


def add_arguments(parser):
    for name, typ, default, text in _FLAGS:
        parser.add_argument("--" + name, type=typ, default=default, help=text,
                            **({"choices": _CHOICES[name]} if name in _CHOICES else {}))
