"""tools/precision_study.py -- why the packed-fp32 ADMM kernel is judged at the reference's own eps (1e-3) and the north-star
parity setting (eps 1e-5) is served by the fp64 kernel.  CPU only (numpy model of csrc/admm_pair.cuh, tools/admm_pcr_model.py).

 1. Spectrum of the reduced Hessian Z'PZ (Z = null space of the dynamics rows) of a reference-assembled QP: one exact zero
     and then ALL thirty curvature inputs at 1e-7 .. 5e-7 -- a dual tolerance of 1e-5 leaves them free by
    O(10), so "the solution at eps 1e-5" is a point on a 2000-pass trajectory, not a well-conditioned minimiser.
 2. The fp32 iteration against the fp64 oracle on the golden QPs, with individual pieces of state promoted to fp64
    (x, u = q + A'y, r = A x - z, v / z) and with the compensated curvature row the kernel uses (comp_kappa):
    at eps 1e-3 the curvature row's v / z is what matters (1.2e-3 -> 5e-4 = the floor set by fp32 data and fp32 linear solves);
    at eps 1e-5 even ALL state and the residual P x + u + A'(rho r) in fp64 leave O(1) errors as long as the scaled
    problem data and the linear solves are fp32, i.e. no mixed-precision variant short of fp64 follows that trajectory.
Prints one JSON line (profiles/r2_precision_study.json)."""
import json, os, sys, time
import numpy as np
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO); sys.path.insert(0, os.path.join(REPO, "tests")); sys.path.insert(0, os.path.join(REPO, "tools"))
from conftest import fixed_pattern, load_golden
from oracle import oracle as orc
from admm_pcr_model import *


def vform_mixed(N, Pd, q, Ax, l, u, hi64=(), rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
               eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25,
               adaptive_rho_interval=25, adaptive_rho_tolerance=5.0, comp=False):
    vl = 0; zl = 0
    cmask = np.array([0, 0, 0, 0, 1], np.float32)  # the curvature bound row only, as in the kernel
    dt = np.dtype(np.float32); T = dt.type
    s = from_reference_layout(N, Pd, q, Ax, l, u, dt)
    L = N + 1
    L2 = 32
    for k in ("a", "c", "e", "P", "q", "d", "lo", "hi"):
        s[k] = np.concatenate([s[k], np.zeros((L2 - L, s[k].shape[1]), dt)])
    L = L2
    mask = np.ones((L, 5), bool); mask[N, 3:] = False; mask[N + 1:] = False
    s["mask"] = mask
    s["e"][N, 3:] = 0; s["lo"][N, 3:] = 0; s["hi"][N, 3:] = 0
    ruiz(s, scaling, 5 * N + 3)
    D, Ed, Eb, cs = s["D"], s["Ed"], s["Eb"], s["cs"]
    lo, hi, dd, qq, P = s["lo"], s["hi"], s["d"], s["q"], s["P"]
    thr = T(OSQP_INFTY * MIN_SCALING)
    ctype = np.where((lo < -thr) & (hi > thr), -1, np.where(hi - lo < T(RHO_TOL), 1, 0))
    def rho_vec(r):
        return np.where(ctype == -1, T(RHO_MIN), np.where(ctype == 1, T(RHO_EQ_OVER_RHO_INEQ * r), T(r))).astype(dt)
    fac, sol = factorize_cr, solve_cr
    rho = T(rho); rb = rho_vec(rho)
    s["P"][N, 3:] = 1.0; fac(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
    al = T(alpha); status = 0
    def st(name): return np.float64 if name in hi64 else np.float32
    x = np.zeros((L, 5), st("x"))
    vb = np.zeros((L, 5), st("v")); zb = np.zeros((L, 5), st("v")); rbd = np.zeros((L, 5), st("r")); rdy = np.zeros((L, 3), st("r"))
    uu = qq.astype(st("u"))
    f32 = lambda a: a.astype(np.float32)
    for it in range(1, max_iter + 1):
        if "g" in hi64:
            s64 = {k: (v.astype(np.float64) if isinstance(v, np.ndarray) else v) for k, v in s.items() if k in ("a","c","e")}
            g = np.where(mask, -((P.astype(np.float64) * x + uu) + At_apply(s64, np.float64(rd) * rdy, rb.astype(np.float64) * rbd)), 0).astype(dt)
        else:
            g = np.where(mask, -((P * f32(x) + f32(uu)) + At_apply(s, rd * f32(rdy), rb * f32(rbd))), 0).astype(dt)
        dl = (al * sol(s, g)).astype(dt)
        s1d, s1b = A_apply(s, dl)
        x = (x + dl).astype(x.dtype)
        wd = (al * f32(rdy) + s1d).astype(dt)
        rdy = (rdy + s1d).astype(rdy.dtype)
        dyd = (rd * wd).astype(dt)
        wb = (al * f32(rbd) + s1b).astype(dt)
        if comp:
            s_ = (vb + wb).astype(dt); bb = (s_ - vb).astype(dt)
            err = ((vb - (s_ - bb).astype(dt)).astype(dt) + (wb - bb).astype(dt)).astype(dt)
            vl = (vl + err * cmask).astype(dt); vb = s_
            zn = np.minimum(np.maximum(vb, lo), hi).astype(dt)
            zln = np.where(zn == vb, vl, T(0)).astype(dt)
            st_h = (zn - zb).astype(dt); st_l = (zln - zl).astype(dt)
            zb = zn; zl = zln
            rbd = (((rbd + s1b).astype(dt) - st_h).astype(dt) - st_l).astype(dt)
            dyb = (rb * ((wb - st_h).astype(dt) - st_l).astype(dt)).astype(dt)
        else:
            vb = (vb + wb).astype(vb.dtype)
            zn = np.minimum(np.maximum(vb, lo), hi).astype(vb.dtype)
            stepb = (zn - zb); zb = zn
            rbd = ((rbd + s1b) - stepb).astype(rbd.dtype)
            dyb = (rb * (wb - f32(stepb))).astype(dt)
        uu = (uu + At_apply(s, dyd, dyb)).astype(uu.dtype)
        if it == 1:
            rdy = (rdy - dd).astype(rdy.dtype)
            dyd = (dyd - rd * dd).astype(dt)
            uu = (uu + At_apply(s, -rd * dd, np.zeros_like(dyb))).astype(uu.dtype)
        if it % check_termination == 0:
            Axd, Axb = A_apply(s, f32(x))
            rpd, rpb = Axd - dd, Axb - f32(zb)
            Px = P * f32(x)
            Aty = f32(uu) - qq
            rdual = np.where(mask, Px + f32(uu), 0)
            pri_res = max(np.max(np.abs(rpd / Ed)), np.max(np.abs(rpb / Eb)))
            dua_res = np.max(np.abs(rdual / D)) / cs
            nz_ = max(np.max(np.abs(dd / Ed)), np.max(np.abs(zb / Eb)))
            nax = max(np.max(np.abs(Axd / Ed)), np.max(np.abs(Axb / Eb)))
            eps_prim = eps_abs + eps_rel * max(nz_, nax)
            nd = max(np.max(np.abs(qq / D)), np.max(np.abs(Aty / D)), np.max(np.abs(Px / D))) / cs
            eps_dual = eps_abs + eps_rel * nd
            if pri_res < eps_prim and dua_res < eps_dual: status = 1; break
            pn = max(np.abs(rpd).max(), np.abs(rpb).max())
            pn /= max(np.max(np.abs(dd)), np.max(np.abs(zb)), np.max(np.abs(Axd)), np.max(np.abs(Axb))) + 1e-10
            dn = np.max(np.abs(rdual)); dn /= max(np.max(np.abs(qq)), np.max(np.abs(Aty)), np.max(np.abs(Px))) + 1e-10
            rnew = float(rho) * np.sqrt(pn / (dn + 1e-10)); rnew = min(max(rnew, RHO_MIN), RHO_MAX)
            if rnew > float(rho) * adaptive_rho_tolerance or rnew < float(rho) / adaptive_rho_tolerance:
                rb_old = rb
                rho = T(rnew); rb = rho_vec(rho); rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
                if comp:
                    vb = (zb + (((vb - zb).astype(dt) + (vl - zl).astype(dt)) * (rb_old / rb))).astype(dt); vl = zl
                else:
                    vb = (zb + (vb - zb) * (rb_old / rb)).astype(vb.dtype)
                s["P"][N, 3:] = 1.0; fac(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    if status == 0: status = -2
    xs = (D * x).astype(np.float64)
    return dict(x=np.concatenate([xs[:N + 1, :3].ravel(), xs[:N, 3:].ravel()]), iter=it, status=status)



def spectrum(N, Pd, Ax):
    from scipy import sparse
    from scipy.linalg import null_space
    Ap, Ai = fixed_pattern(N)
    A = sparse.csc_matrix((Ax, Ai, Ap), shape=(8 * N + 6, 5 * N + 3)).toarray()
    Z = null_space(A[:3 * (N + 1)])
    w = np.linalg.eigvalsh(Z.T @ np.diag(Pd) @ Z)
    return [float(v) for v in w]


def main(n_tight=10):
    TF, C1 = load_golden("teacher_forced.npz"), load_golden("c1_lap.npz")
    Pd, q, Ax, l, u = (np.concatenate([TF["qp_" + k], C1["qp_" + k]]) for k in ("Pd", "q", "Ax", "l", "u"))
    Ap, Ai = fixed_pattern(30)
    out = {"reduced_hessian_eigenvalues_qp0": spectrum(30, Pd[0], Ax[0])}
    variants = [("fp32", (), False), ("fp32 + compensated kappa row (the kernel)", (), True), ("x fp64", ("x",), False),
                ("u fp64", ("u",), False), ("r fp64", ("r",), False), ("v, z fp64", ("v",), False),
                ("x, u, r, v, z fp64", ("u", "r", "v", "x"), False),
                ("all state + the residual P x + u + A'(rho r) in fp64; data and linear solves fp32", ("u", "r", "v", "x", "g"), False)]
    for eps in (1e-3, 1e-5):
        xo, ito, sto = orc.batch_qp_solve(30, Pd, q, Ap, Ai, Ax, l, u, eps_abs=eps, eps_rel=eps)
        sel = [b for b in range(len(sto)) if sto[b] == 1]
        if eps < 1e-4:
            sel = sel[:n_tight]
        rows = []
        for name, hi64, comp in variants:
            errs, dit = [], 0
            for b in sel:
                r = vform_mixed(30, Pd[b], q[b], Ax[b], l[b], u[b], hi64=hi64, eps_abs=eps, eps_rel=eps, comp=comp)
                errs.append(float(np.abs(r["x"] - xo[b]).max())); dit += int(r["iter"] != ito[b])
            rows.append(dict(variant=name, max_err=max(errs), median_err=float(np.median(errs)), other_iteration_count=dit))
        out["eps_%g" % eps] = dict(n_qps=len(sel), oracle_mean_iters=float(np.mean(ito[sel])), rows=rows)
    print(json.dumps(out))


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
