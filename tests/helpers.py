"""Shared helpers for the parity tests."""
import numpy as np


def is_approx(a, b, prec):
    """Eigen's isApprox: ||a-b||^2 <= prec^2 * min(||a||^2, ||b||^2)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.sum((a - b) ** 2) <= prec * prec * min(np.sum(a * a), np.sum(b * b))


# north_star: "same termination status, primal solution within 1e-6 relative"
X_REL_TOL = 1e-6


def assert_parity(got, ref, x_rel_tol=X_REL_TOL, y_abs_tol=1e-5, what=""):
    """Per-QP comparison of a CUDA result dict with the oracle's; failures are listed individually
    (SURVEY.md section 7 'decision parity': never average them away)."""
    B = ref["status"].shape[0]
    bad = []
    for i in range(B):
        msgs = []
        if int(got["status"][i]) != int(ref["status"][i]):
            msgs.append("status %d != %d" % (got["status"][i], ref["status"][i]))
        if int(got["iter"][i]) != int(ref["iter"][i]):
            msgs.append("iter %d != %d" % (got["iter"][i], ref["iter"][i]))
        if "rho_updates" in got and int(got["rho_updates"][i]) != int(ref["rho_updates"][i]):
            msgs.append("rho_updates %d != %d" % (got["rho_updates"][i], ref["rho_updates"][i]))
        nx = np.linalg.norm(ref["x"][i])
        dx = np.linalg.norm(got["x"][i] - ref["x"][i])
        if not (dx <= x_rel_tol * max(nx, 1e-300)) and not (nx == 0 and dx == 0):
            msgs.append("x rel err %.3e" % (dx / max(nx, 1e-300)))
        if "y" in got:
            dy = np.abs(got["y"][i] - ref["y"][i]).max() if ref["y"][i].size else 0.0
            sy = max(1.0, np.abs(ref["y"][i]).max() if ref["y"][i].size else 0.0)
            if not dy <= y_abs_tol * sy:
                msgs.append("y abs err %.3e" % dy)
        if msgs:
            bad.append("QP %d: %s" % (i, "; ".join(msgs)))
    assert not bad, "%s parity failures (%d of %d):\n%s" % (what, len(bad), B, "\n".join(bad[:20]))
    xr = [np.linalg.norm(got["x"][i] - ref["x"][i]) / max(np.linalg.norm(ref["x"][i]), 1e-300) for i in range(B)]
    return max(xr) if xr else 0.0


def oracle_settings_from(oracle, s):
    """Copy a sqp_solver_b200.api.Settings into the oracle's settings struct."""
    return oracle.default_settings(**{k: getattr(s, k) for k, _ in s._fields_})
