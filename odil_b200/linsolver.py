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

    if isinstance(matr, StencilJacobian):
        import torch

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


def add_arguments(parser):
    parser.add_argument("--linsolver", type=str, default="direct",
                        choices=["multigrid", "direct", "directsq", "direct_cu", "sparseqr", "lsqr", "lsqr_cu",
                                 "bicgstab", "cg_b200"], help="Linear solver to use (cg_b200: matrix-free CG on the GPU)")
    parser.add_argument("--linsolver_maxiter", type=int, default=None,
                        help="Maximum number of iterations of linear solver")
    parser.add_argument("--linsolver_tol", type=float, default=1e-6, help="Tolerance for linear solver")
    parser.add_argument("--linsolver_damp", type=float, default=0, help="Relaxation factor (0: no relaxation)")
    parser.add_argument("--linsolver_dampdiag", type=float, default=0,
                        help="Multiplier for diagonal (0: no relaxation)")
    parser.add_argument("--linsolver_verbose", type=int, default=0, help="Verbosity level for linsolver messages")
    parser.add_argument("--linsolver_history", type=int, default=0, help="Dump history from linsolver status")
    parser.add_argument("--lr", type=float, default=1e-3, help="Learning rate")
    parser.add_argument("--nlvl", type=int, default=100, help="Multigrid levels")
    parser.add_argument("--smooth_pre", type=int, default=2, help="Pre-smoothing steps")
    parser.add_argument("--smooth_post", type=int, default=2, help="Post-smoothing steps")
    parser.add_argument("--omega", type=float, default=0.6, help="Jacobi smoother relaxation factor")
    parser.add_argument("--ndirect", type=int, default=3, help="Systems on smaller grids are solved with direct solver")
    parser.add_argument("--restriction", type=str, choices=("full", "half", "injection"), default="full",
                        help="Multigrid restriction type")
