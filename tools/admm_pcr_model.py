"""tools/admm_pcr_model.py -- executable design model of the K2 ADMM kernel (development aid).

Lane-per-stage formulation of the OSQP iteration for the MPC QP (reference layout: MPC.py:128-155),
with the reduced KKT system solved by eliminating the two inputs of every stage and running
parallel cyclic reduction (PCR) over the resulting block-tridiagonal system of 3x3 blocks.
Every array is indexed [lane, ...]; a "shuffle" is an index shift.  Written in numpy so that the
same code runs in float32 and float64 on the CPU; csrc/admm_kernel.cuh is the CUDA transcription.

Not part of the product path and not the oracle: tests use it only to check the algorithm
(not-gpu tier) against oracle/osqp_oracle.c.
"""
import numpy as np

OSQP_INFTY = 1e30
MIN_SCALING, MAX_SCALING = 1e-4, 1e4
RHO_MIN, RHO_MAX = 1e-6, 1e6
RHO_EQ_OVER_RHO_INEQ = 1e3
RHO_TOL = 1e-4


def up(v, s=1, fill=0.0):
    """value held by lane j-s (shfl_up); lanes < s get fill"""
    out = np.full_like(v, fill)
    if s < v.shape[0]:
        out[s:] = v[:-s]
    return out


def down(v, s=1, fill=0.0):
    """value held by lane j+s (shfl_down)"""
    out = np.full_like(v, fill)
    if s < v.shape[0]:
        out[:-s] = v[s:]
    return out


def from_reference_layout(N, Pd, q, Ax, l, u, dtype):
    """Split the reference-layout QP (fixed CSC pattern of oracle orc_mpc_assemble) into per-stage data.
    lane j in 0..N:  a[j, 8] = nonzeros of [A_j B_j] (rows of dynamics block j+1):
        a0=A[0,0] a1=A[0,1] a2=A[1,0] a3=A[1,1] a4=A[2,0] a5=A[2,2] a6=B[1,1](kappa) a7=B[2,0](v)
      c[j,3] = the -1 entries of dynamics block j on x_j;  e[j,5] = bound-row entries (+1)
      P[j,5], q[j,5]; dyn rhs d[j,3] (l=u); bound lo[j,5], hi[j,5].  Stage N has no inputs (masked)."""
    nx, nu = 3, 2
    L = N + 1
    neq = nx * L
    a = np.zeros((L, 8)); c = np.zeros((L, 3)); e = np.zeros((L, 5))
    P = np.zeros((L, 5)); qq = np.zeros((L, 5)); d = np.zeros((L, 3))
    lo = np.zeros((L, 5)); hi = np.zeros((L, 5))
    # CSC walk identical to the oracle's emission order
    nz = 0
    for col in range(neq):
        k, j = divmod(col, nx)
        c[k, j] = Ax[nz]; nz += 1
        if k < N:
            if j == 0:
                a[k, 0], a[k, 2], a[k, 4] = Ax[nz], Ax[nz + 1], Ax[nz + 2]; nz += 3
            elif j == 1:
                a[k, 1], a[k, 3] = Ax[nz], Ax[nz + 1]; nz += 2
            else:
                a[k, 5] = Ax[nz]; nz += 1
        e[k, j] = Ax[nz]; nz += 1
    for col in range(neq, neq + nu * N):
        k, j = divmod(col - neq, nu)
        if j == 0:
            a[k, 7] = Ax[nz]; nz += 1
        else:
            a[k, 6] = Ax[nz]; nz += 1
        e[k, 3 + j] = Ax[nz]; nz += 1
    assert nz == 16 * N + 6
    for k in range(L):
        P[k, :3] = Pd[nx * k: nx * k + 3]; qq[k, :3] = q[nx * k: nx * k + 3]
        d[k] = l[nx * k: nx * k + 3]
        lo[k, :3] = l[neq + nx * k: neq + nx * k + 3]; hi[k, :3] = u[neq + nx * k: neq + nx * k + 3]
        if k < N:
            P[k, 3:] = Pd[neq + nu * k: neq + nu * k + 2]; qq[k, 3:] = q[neq + nu * k: neq + nu * k + 2]
            lo[k, 3:] = l[2 * neq + nu * k: 2 * neq + nu * k + 2]
            hi[k, 3:] = u[2 * neq + nu * k: 2 * neq + nu * k + 2]
    f = lambda v: v.astype(dtype)
    return dict(N=N, a=f(a), c=f(c), e=f(e), P=f(P), q=f(qq), d=f(d),
                lo=f(np.maximum(lo, -OSQP_INFTY)), hi=f(np.minimum(hi, OSQP_INFTY)))


def A_apply(s, w):
    """z = A w per lane: returns (zd[L,3] dynamics rows, zb[L,5] bound rows)."""
    a, c, e = s["a"], s["c"], s["e"]
    # outgoing = [A_j B_j] w_j  (rows of block j+1)
    o = np.stack([a[:, 0] * w[:, 0] + a[:, 1] * w[:, 1],
                  a[:, 2] * w[:, 0] + a[:, 3] * w[:, 1] + a[:, 6] * w[:, 4],
                  a[:, 4] * w[:, 0] + a[:, 5] * w[:, 2] + a[:, 7] * w[:, 3]], axis=1)
    zd = c * w[:, :3] + up(o)
    zb = e * w
    return zd, zb


def At_apply(s, yd, yb):
    """A' y per lane -> [L,5]"""
    a, c, e = s["a"], s["c"], s["e"]
    g = down(yd)  # dual of dynamics block j+1
    r = np.empty_like(yb)
    r[:, 0] = c[:, 0] * yd[:, 0] + a[:, 0] * g[:, 0] + a[:, 2] * g[:, 1] + a[:, 4] * g[:, 2]
    r[:, 1] = c[:, 1] * yd[:, 1] + a[:, 1] * g[:, 0] + a[:, 3] * g[:, 1]
    r[:, 2] = c[:, 2] * yd[:, 2] + a[:, 5] * g[:, 2]
    r[:, 3] = a[:, 7] * g[:, 2]
    r[:, 4] = a[:, 6] * g[:, 1]
    return r + e * yb


def limit_scaling(v):
    v = np.where(v < MIN_SCALING, 1.0, v)
    return np.where(v > MAX_SCALING, MAX_SCALING, v).astype(v.dtype)


def ruiz(s, iters, nvar):
    """OSQP scale_data on the stage layout.  Adds D[L,5], Ed[L,3], Eb[L,5], cscale."""
    dt = s["a"].dtype
    L = s["a"].shape[0]
    D = np.ones((L, 5), dt); Ed = np.ones((L, 3), dt); Eb = np.ones((L, 5), dt)
    cs = dt.type(1.0)
    a, c, e, P, q = s["a"], s["c"], s["e"], s["P"], s["q"]
    mask = s["mask"]
    for _ in range(iters):
        aa = np.abs(a)
        col = np.stack([np.maximum.reduce([aa[:, 0], aa[:, 2], aa[:, 4]]), np.maximum(aa[:, 1], aa[:, 3]),
                        aa[:, 5], aa[:, 7], aa[:, 6]], axis=1)
        col[:, :3] = np.maximum(col[:, :3], np.abs(c))
        col = np.maximum(col, np.abs(e))
        col = np.maximum(col, np.abs(P))
        rowo = np.stack([np.maximum(aa[:, 0], aa[:, 1]), np.maximum.reduce([aa[:, 2], aa[:, 3], aa[:, 6]]),
                         np.maximum.reduce([aa[:, 4], aa[:, 5], aa[:, 7]])], axis=1)
        rowd = np.maximum(np.abs(c), up(rowo))
        rowb = np.abs(e)
        Dt = (1.0 / np.sqrt(limit_scaling(col))).astype(dt)
        Edt = (1.0 / np.sqrt(limit_scaling(rowd))).astype(dt)
        Ebt = (1.0 / np.sqrt(limit_scaling(rowb))).astype(dt)
        Dt = np.where(mask, Dt, 1.0).astype(dt)
        Ebt = np.where(mask, Ebt, 1.0).astype(dt)
        P = P * Dt * Dt
        En = down(Edt)  # row scaling of dynamics block j+1
        a = np.stack([a[:, 0] * En[:, 0] * Dt[:, 0], a[:, 1] * En[:, 0] * Dt[:, 1],
                      a[:, 2] * En[:, 1] * Dt[:, 0], a[:, 3] * En[:, 1] * Dt[:, 1],
                      a[:, 4] * En[:, 2] * Dt[:, 0], a[:, 5] * En[:, 2] * Dt[:, 2],
                      a[:, 6] * En[:, 1] * Dt[:, 4], a[:, 7] * En[:, 2] * Dt[:, 3]], axis=1).astype(dt)
        c = c * Edt * Dt[:, :3]
        e = e * Ebt * Dt
        q = q * Dt
        D = D * Dt; Ed = Ed * Edt; Eb = Eb * Ebt
        # cost scaling
        ctemp = np.sum(np.abs(P), dtype=dt) / dt.type(nvar)
        nq = np.max(np.abs(q))
        nq = limit_scaling(np.array([nq], dt))[0]
        ctemp = max(ctemp, nq)
        ctemp = limit_scaling(np.array([ctemp], dt))[0]
        ctemp = dt.type(1.0) / ctemp
        P = P * ctemp; q = q * ctemp; cs = cs * ctemp
    s.update(a=a, c=c, e=e, P=P.astype(dt), q=q.astype(dt), D=D, Ed=Ed, Eb=Eb, cs=cs)
    s["d"] = s["d"] * Ed
    s["lo"] = s["lo"] * Eb
    s["hi"] = s["hi"] * Eb


def inv3(M):
    """batched inverse of symmetric positive definite 3x3 via adjugate"""
    a, b, c = M[:, 0, 0], M[:, 0, 1], M[:, 0, 2]
    d, e, f = M[:, 1, 1], M[:, 1, 2], M[:, 2, 2]
    A = d * f - e * e; B = c * e - b * f; C = b * e - c * d
    det = a * A + b * B + c * C
    r = 1.0 / det
    out = np.empty_like(M)
    out[:, 0, 0] = A * r; out[:, 0, 1] = out[:, 1, 0] = B * r; out[:, 0, 2] = out[:, 2, 0] = C * r
    out[:, 1, 1] = (a * f - c * c) * r; out[:, 1, 2] = out[:, 2, 1] = (b * c - a * e) * r
    out[:, 2, 2] = (a * d - b * b) * r
    return out.astype(M.dtype)


def factorize(s, sigma, rho, rho_b):
    """Build S = P + sigma I + A' R A per stage, eliminate inputs, PCR-factorise.  rho_b[L,5]."""
    dt = s["a"].dtype
    a, c, e, P = s["a"], s["c"], s["e"], s["P"]
    L = a.shape[0]
    rd = dt.type(RHO_EQ_OVER_RHO_INEQ * rho)  # all dynamics rows are equalities
    diag = P + dt.type(sigma) + rho_b * e * e
    cn = down(c)  # c of stage j+1
    # [A_j B_j]' R [A_j B_j] : rows r0=(a0,a1,0,0,0) r1=(a2,a3,0,0,a6) r2=(a4,0,a5,a7,0)
    Sxx = np.zeros((L, 3, 3), dt)
    Sxx[:, 0, 0] = diag[:, 0] + rd * (c[:, 0] ** 2 + a[:, 0] ** 2 + a[:, 2] ** 2 + a[:, 4] ** 2)
    Sxx[:, 1, 1] = diag[:, 1] + rd * (c[:, 1] ** 2 + a[:, 1] ** 2 + a[:, 3] ** 2)
    Sxx[:, 2, 2] = diag[:, 2] + rd * (c[:, 2] ** 2 + a[:, 5] ** 2)
    Sxx[:, 0, 1] = Sxx[:, 1, 0] = rd * (a[:, 0] * a[:, 1] + a[:, 2] * a[:, 3])
    Sxx[:, 0, 2] = Sxx[:, 2, 0] = rd * (a[:, 4] * a[:, 5])
    Sxx[:, 1, 2] = Sxx[:, 2, 1] = 0
    # inputs: v (col 3) only in row r2 via a7; kappa (col 4) only in row r1 via a6
    Svv = diag[:, 3] + rd * a[:, 7] ** 2
    Skk = diag[:, 4] + rd * a[:, 6] ** 2
    # S_xu columns
    Sxv = np.stack([rd * a[:, 4] * a[:, 7], np.zeros(L, dt), rd * a[:, 5] * a[:, 7]], axis=1)
    Sxk = np.stack([rd * a[:, 2] * a[:, 6], rd * a[:, 3] * a[:, 6], np.zeros(L, dt)], axis=1)
    # coupling to x_{j+1}: F_x[i, r] = rd * (coef of x_i in row r) * cn[r];  F_v -> t_{j+1}, F_k -> e_psi_{j+1}
    Fx = np.zeros((L, 3, 3), dt)
    Fx[:, 0, 0] = rd * a[:, 0] * cn[:, 0]; Fx[:, 1, 0] = rd * a[:, 1] * cn[:, 0]
    Fx[:, 0, 1] = rd * a[:, 2] * cn[:, 1]; Fx[:, 1, 1] = rd * a[:, 3] * cn[:, 1]
    Fx[:, 0, 2] = rd * a[:, 4] * cn[:, 2]; Fx[:, 2, 2] = rd * a[:, 5] * cn[:, 2]
    Fv = rd * a[:, 7] * cn[:, 2]
    Fk = rd * a[:, 6] * cn[:, 1]
    iv, ik = (1.0 / Svv).astype(dt), (1.0 / Skk).astype(dt)
    # Schur complement of the inputs
    Dm = Sxx - iv[:, None, None] * Sxv[:, :, None] * Sxv[:, None, :] - ik[:, None, None] * Sxk[:, :, None] * Sxk[:, None, :]
    U = Fx.copy()  # coupling block (row j, col j+1)
    U[:, :, 2] -= (iv * Fv)[:, None] * Sxv
    U[:, :, 1] -= (ik * Fk)[:, None] * Sxk
    addn = np.stack([np.zeros(L, dt), ik * Fk * Fk, iv * Fv * Fv], axis=1)
    addn = up(addn)
    for i in range(3):
        Dm[:, i, i] -= addn[:, i]
    Lo = np.transpose(up(U), (0, 2, 1))  # coupling block (row j, col j-1) = U_{j-1}'
    # zero couplings that leave the chain
    U[-1] = 0
    Lo[0] = 0
    levels = []
    sft = 1
    while sft < L:
        Dinv = inv3(Dm)
        al = np.einsum("lij,ljk->lik", Lo, up(Dinv, sft))     # L_j D_{j-s}^-1
        be = np.einsum("lij,ljk->lik", U, down(Dinv, sft))    # U_j D_{j+s}^-1
        Dm = Dm - np.einsum("lij,ljk->lik", al, up(U, sft)) - np.einsum("lij,ljk->lik", be, down(Lo, sft))
        Lo_n = -np.einsum("lij,ljk->lik", al, up(Lo, sft))
        U_n = -np.einsum("lij,ljk->lik", be, down(U, sft))
        Lo, U = Lo_n.astype(dt), U_n.astype(dt)
        Dm = Dm.astype(dt)
        levels.append((al.astype(dt), be.astype(dt), sft))
        sft *= 2
    s["fac"] = dict(levels=levels, Dinv=inv3(Dm), iv=iv, ik=ik, Sxv=Sxv, Sxk=Sxk, Fv=Fv, Fk=Fk)


def solve(s, b):
    """x = S^-1 b, b[L,5]"""
    f = s["fac"]
    bx = b[:, :3] - (f["iv"] * b[:, 3])[:, None] * f["Sxv"] - (f["ik"] * b[:, 4])[:, None] * f["Sxk"]
    # input elimination also touches x_{j+1}:  -F_u' S_uu^-1 b_u
    tn = np.stack([np.zeros_like(b[:, 0]), f["ik"] * f["Fk"] * b[:, 4], f["iv"] * f["Fv"] * b[:, 3]], axis=1)
    bx = bx - up(tn)
    for al, be, sft in f["levels"]:
        bx = bx - np.einsum("lij,lj->li", al, up(bx, sft)) - np.einsum("lij,lj->li", be, down(bx, sft))
    x = np.einsum("lij,lj->li", f["Dinv"], bx)
    xn = down(x)
    v = f["iv"] * (b[:, 3] - np.einsum("li,li->l", f["Sxv"], x) - f["Fv"] * xn[:, 2])
    k = f["ik"] * (b[:, 4] - np.einsum("li,li->l", f["Sxk"], x) - f["Fk"] * xn[:, 1])
    return np.concatenate([x, v[:, None], k[:, None]], axis=1).astype(b.dtype)


def admm(N, Pd, q, Ax, l, u, dtype=np.float64, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
         eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25,
         adaptive_rho_interval=25, adaptive_rho_tolerance=5.0, refine=0):
    """OSQP iteration in stage form.  Returns dict(x (reference layout), iter, status, ...)."""
    dt = np.dtype(dtype)
    T = dt.type
    s = from_reference_layout(N, Pd, q, Ax, l, u, dt)
    L = N + 1
    mask = np.ones((L, 5), bool); mask[N, 3:] = False
    s["mask"] = mask
    # the missing inputs of stage N: identity-like dummy (e=1, bounds 0) so S stays SPD and w stays 0
    s["e"][N, 3:] = 0; s["lo"][N, 3:] = 0; s["hi"][N, 3:] = 0
    nvar = 5 * N + 3
    ruiz(s, scaling, nvar)
    a, c, e = s["a"], s["c"], s["e"]
    D, Ed, Eb, cs = s["D"], s["Ed"], s["Eb"], s["cs"]
    lo, hi, dd, qq, P = s["lo"], s["hi"], s["d"], s["q"], s["P"]
    thr = T(OSQP_INFTY * MIN_SCALING)
    ctype = np.where((lo < -thr) & (hi > thr), -1, np.where(hi - lo < T(RHO_TOL), 1, 0))
    ctype[N, 3:] = 1

    def rho_vec(r):
        rb = np.where(ctype == -1, T(RHO_MIN), np.where(ctype == 1, T(RHO_EQ_OVER_RHO_INEQ * r), T(r))).astype(dt)
        return rb

    rho = T(rho)
    rb = rho_vec(rho)
    # dummy inputs of the last stage: give them a unit diagonal in S
    s["P"][N, 3:] = 1.0
    factorize(s, sigma, rho, rb)
    s["P"][N, 3:] = 0.0
    rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
    x = np.zeros((L, 5), dt); zd = np.zeros((L, 3), dt); zb = np.zeros((L, 5), dt)
    yd = np.zeros((L, 3), dt); yb = np.zeros((L, 5), dt)
    status, it, n_fac, rho_updates = 0, 0, 1, 0
    al = T(alpha)
    for it in range(1, max_iter + 1):
        rhs = T(sigma) * x - qq + At_apply(s, rd * zd - yd, rb * zb - yb)
        rhs = np.where(mask, rhs, 0).astype(dt)
        xt = solve(s, rhs)
        for _ in range(refine):
            zd_, zb_ = A_apply(s, xt)
            Sx = (P + T(sigma)) * xt + At_apply(s, rd * zd_, rb * zb_)
            res = np.where(mask, rhs - Sx, 0).astype(dt)
            xt = xt + solve(s, res)
        ztd, ztb = A_apply(s, xt)
        xn = al * xt + (1 - al) * x
        dx = xn - x
        vd = al * ztd + (1 - al) * zd
        vb = al * ztb + (1 - al) * zb
        zdn = dd.copy()  # equality rows: projection onto {d}
        zbn = np.minimum(np.maximum(vb + yb / rb, lo), hi)
        dyd = rd * (vd - zdn)
        dyb = rb * (vb - zbn)
        yd = yd + dyd; yb = yb + dyb
        x, zd, zb = xn.astype(dt), zdn.astype(dt), zbn.astype(dt)
        yd, yb = yd.astype(dt), yb.astype(dt)
        if it % check_termination == 0 or it % adaptive_rho_interval == 0:
            Axd, Axb = A_apply(s, x)
            rpd, rpb = Axd - zd, Axb - zb
            Px = P * x
            Aty = At_apply(s, yd, yb)
            rdual = np.where(mask, Px + qq + Aty, 0)
            pri_res = max(np.max(np.abs(rpd / Ed)), np.max(np.abs(rpb / Eb)))
            dua_res = np.max(np.abs(rdual / D)) / cs
            if it % check_termination == 0:
                nz_ = max(np.max(np.abs(zd / Ed)), np.max(np.abs(zb / Eb)))
                nax = max(np.max(np.abs(Axd / Ed)), np.max(np.abs(Axb / Eb)))
                eps_prim = eps_abs + eps_rel * max(nz_, nax)
                nd = max(np.max(np.abs(qq / D)), np.max(np.abs(Aty / D)), np.max(np.abs(Px / D))) / cs
                eps_dual = eps_abs + eps_rel * nd
                if pri_res < eps_prim and dua_res < eps_dual:
                    status = 1
                    break
                if not (pri_res < eps_prim):
                    # primal infeasibility certificate on delta_y
                    pyd = dyd  # equality rows: finite bounds, keep
                    pyb = np.where(hi > thr, np.where(lo < -thr, 0, np.minimum(dyb, 0)),
                                   np.where(lo < -thr, np.maximum(dyb, 0), dyb))
                    ndy = max(np.max(np.abs(Ed * pyd)), np.max(np.abs(Eb * pyb)))
                    if ndy > eps_prim_inf:
                        lhs = np.sum(dd * pyd) + np.sum(hi * np.maximum(pyb, 0) + lo * np.minimum(pyb, 0))
                        if lhs < -eps_prim_inf * ndy:
                            Atdy = np.where(mask, At_apply(s, pyd, pyb), 0)
                            if np.max(np.abs(Atdy / D)) < eps_prim_inf * ndy:
                                status = -3
                                break
                # (dual infeasibility check omitted in the model: P + A'A has full rank here; the
                #  CUDA kernel implements it)
            if it % adaptive_rho_interval == 0:
                pn = np.max(np.abs(np.concatenate([rpd.ravel(), rpb.ravel()])))
                pn /= max(np.max(np.abs(zd)), np.max(np.abs(zb)), np.max(np.abs(Axd)), np.max(np.abs(Axb))) + 1e-10
                dn = np.max(np.abs(rdual))
                dn /= max(np.max(np.abs(qq)), np.max(np.abs(Aty)), np.max(np.abs(Px))) + 1e-10
                rnew = float(rho) * np.sqrt(pn / (dn + 1e-10))
                rnew = min(max(rnew, RHO_MIN), RHO_MAX)
                if rnew > float(rho) * adaptive_rho_tolerance or rnew < float(rho) / adaptive_rho_tolerance:
                    rho = T(rnew)
                    rb = rho_vec(rho)
                    rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
                    s["P"][N, 3:] = 1.0
                    factorize(s, sigma, rho, rb)
                    s["P"][N, 3:] = 0.0
                    n_fac += 1; rho_updates += 1
    if status == 0:
        status = -2
    xs = (D * x).astype(np.float64)
    xout = np.concatenate([xs[:, :3].ravel(), xs[:N, 3:].ravel()])
    return dict(x=xout, iter=it, status=status, rho=float(rho), rho_updates=rho_updates, n_factor=n_fac)


# ------------------------------------------------------------------------------------------------------
# increment ("delta") form -- what csrc/admm.cuh::admm_solve implements
# ------------------------------------------------------------------------------------------------------
# The linear solve returns D = x~ - x from  S D = -(q + P x + A'y + A'(R r))  with tracked residuals r = A x - z and
# a tracked t = A'y (t += A'dy); rows update through  v - z = alpha (r + A D).  Same iterates as admm() in exact arithmetic; in fp32 every
# quantity multiplied by rho_eq = 1e3 rho is a small residual instead of a difference of O(1) numbers, which is
# what lets single precision reproduce OSQP's iteration counts and infeasibility certificates at eps = 1e-3.
def admm_delta(N, Pd, q, Ax, l, u, dtype=np.float32, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
         eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25,
         adaptive_rho_interval=25, adaptive_rho_tolerance=5.0, resync=False):
    dt = np.dtype(dtype); T = dt.type
    s = from_reference_layout(N, Pd, q, Ax, l, u, dt)
    L = N + 1
    mask = np.ones((L, 5), bool); mask[N, 3:] = False
    s["mask"] = mask
    s["e"][N, 3:] = 0; s["lo"][N, 3:] = 0; s["hi"][N, 3:] = 0
    ruiz(s, scaling, 5*N+3)
    D, Ed, Eb, cs = s["D"], s["Ed"], s["Eb"], s["cs"]
    lo, hi, dd, qq, P = s["lo"], s["hi"], s["d"], s["q"], s["P"]
    thr = T(OSQP_INFTY * MIN_SCALING)
    ctype = np.where((lo < -thr) & (hi > thr), -1, np.where(hi - lo < T(RHO_TOL), 1, 0))
    def rho_vec(r):
        return np.where(ctype == -1, T(RHO_MIN), np.where(ctype == 1, T(RHO_EQ_OVER_RHO_INEQ * r), T(r))).astype(dt)
    rho = T(rho); rb = rho_vec(rho)
    s["P"][N, 3:] = 1.0; factorize(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
    x = np.zeros((L, 5), dt); zb = np.zeros((L, 5), dt); yd = np.zeros((L, 3), dt); yb = np.zeros((L, 5), dt)
    # tracked residuals r = A x - z ; cold start x = z = 0 -> r = 0 ; the dynamics z jumps 0 -> d in iteration 1
    rdy = np.zeros((L, 3), dt); rbd = np.zeros((L, 5), dt)
    ty = np.zeros((L, 5), dt)  # tracked A'y: accumulated from the small dual steps, never recomputed from y
    zd = np.zeros((L, 3), dt)
    al = T(alpha); status = 0
    for it in range(1, max_iter + 1):
        g = -(qq + P * x + ty + At_apply(s, rd * rdy, rb * rbd))
        g = np.where(mask, g, 0).astype(dt)
        dl = solve(s, g)
        Add, Adb = A_apply(s, dl)
        x = (x + al * dl).astype(dt); dx = al*dl
        # dynamics rows: z_new = d
        wd = al * (rdy + Add)                 # v - z_prev
        zdn = dd
        stepd = zdn - zd                      # exactly 0 after the first iteration
        dyd = rd * (wd - stepd); yd = (yd + dyd).astype(dt)
        rdy = (rdy + al * Add - stepd).astype(dt); zd = zdn.copy()
        # bound rows
        wb = al * (rbd + Adb)
        zbn = np.minimum(np.maximum(zb + wb + yb / rb, lo), hi).astype(dt)
        stepb = (zbn - zb).astype(dt)
        dyb = rb * (wb - stepb); yb = (yb + dyb).astype(dt)
        rbd = (rbd + al * Adb - stepb).astype(dt); zb = zbn
        ty = (ty + At_apply(s, dyd, dyb)).astype(dt)
        if it % check_termination == 0 or it % adaptive_rho_interval == 0:
            Axd, Axb = A_apply(s, x)
            rpd, rpb = Axd - zd, Axb - zb
            drift = max(np.abs(rpd - rdy).max(), np.abs(rpb - rbd).max())
            if resync: rdy, rbd = rpd.astype(dt), rpb.astype(dt)
            Px = P * x; Aty = ty
            rdual = np.where(mask, Px + qq + Aty, 0)
            pri_res = max(np.max(np.abs(rpd / Ed)), np.max(np.abs(rpb / Eb)))
            dua_res = np.max(np.abs(rdual / D)) / cs
            if it % check_termination == 0:
                nz_ = max(np.max(np.abs(zd / Ed)), np.max(np.abs(zb / Eb)))
                nax = max(np.max(np.abs(Axd / Ed)), np.max(np.abs(Axb / Eb)))
                eps_prim = eps_abs + eps_rel * max(nz_, nax)
                nd = max(np.max(np.abs(qq / D)), np.max(np.abs(Aty / D)), np.max(np.abs(Px / D))) / cs
                eps_dual = eps_abs + eps_rel * nd
                if pri_res < eps_prim and dua_res < eps_dual: status = 1; break
                if not (pri_res < eps_prim):
                    pyd = dyd
                    pyb = np.where(hi > thr, np.where(lo < -thr, 0, np.minimum(dyb, 0)), np.where(lo < -thr, np.maximum(dyb, 0), dyb))
                    ndy = max(np.max(np.abs(Ed * pyd)), np.max(np.abs(Eb * pyb)))
                    if ndy > eps_prim_inf:
                        lhs = np.sum(dd * pyd) + np.sum(hi * np.maximum(pyb, 0) + lo * np.minimum(pyb, 0))
                        if lhs < -eps_prim_inf * ndy:
                            Atdy = np.where(mask, At_apply(s, pyd, pyb), 0)
                            if np.max(np.abs(Atdy / D)) < eps_prim_inf * ndy: status = -3; break
            if it % adaptive_rho_interval == 0:
                pn = max(np.abs(rpd).max(), np.abs(rpb).max())
                pn /= max(np.max(np.abs(zd)), np.max(np.abs(zb)), np.max(np.abs(Axd)), np.max(np.abs(Axb))) + 1e-10
                dn = np.max(np.abs(rdual)); dn /= max(np.max(np.abs(qq)), np.max(np.abs(Aty)), np.max(np.abs(Px))) + 1e-10
                rnew = float(rho) * np.sqrt(pn / (dn + 1e-10)); rnew = min(max(rnew, RHO_MIN), RHO_MAX)
                if rnew > float(rho) * adaptive_rho_tolerance or rnew < float(rho) / adaptive_rho_tolerance:
                    rho = T(rnew); rb = rho_vec(rho); rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
                    s["P"][N, 3:] = 1.0; factorize(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    if status == 0: status = -2
    xs = (D * x).astype(np.float64)
    return dict(x=np.concatenate([xs[:, :3].ravel(), xs[:N, 3:].ravel()]), iter=it, status=status, drift=float(drift))



# ------------------------------------------------------------------------------------------------------
# "v-form" -- what csrc/admm_pair.cuh implements (fewer operations per iteration than the increment form)
# ------------------------------------------------------------------------------------------------------
# OSQP's row update is  v+ = alpha z~ + (1 - alpha) z + y / rho,  z+ = clip(v+),  y+ = rho (v+ - z+).  Because the previous
# (z, y) came out of the same projection, y / rho = v - z and therefore  v+ = v + alpha (z~ - z) = v + w  with
# w = alpha (r + A D), r = A x - z tracked.  A bound row is the triple (v, z, r); y is never stored (dy = rho (w - (z+ - z))
# feeds the tracked u = q + A'y).  The dynamics rows are equalities whose z jumps from the cold start 0 to d in
# iteration 1 and stays there: they carry r only, and iteration 1 is patched up afterwards (r -= d, u -= A'(rho d)).
# When rho changes, v = z + (v - z) rho / rho+ (y itself is unchanged, as in OSQP).
# (Tracking T = A'(y + rho r) through its increments instead, to save the second A' product per iteration, was tried
#  and rejected: its rounding error accumulates and triples the fp32 error of the solution.)
def admm_vform(N, Pd, q, Ax, l, u, dtype=np.float32, rho=0.1, sigma=1e-6, alpha=1.6, eps_abs=1e-3, eps_rel=1e-3,
               eps_prim_inf=1e-4, eps_dual_inf=1e-4, max_iter=4000, scaling=10, check_termination=25,
               adaptive_rho_interval=25, adaptive_rho_tolerance=5.0, paired=False):
    dt = np.dtype(dtype); T = dt.type
    s = from_reference_layout(N, Pd, q, Ax, l, u, dt)
    L = N + 1
    if paired:  # pad to an even number of stages >= the lane-group size the kernel uses
        L2 = 16 if L <= 16 else (32 if L <= 32 else 64)
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
    fac = factorize_cr if paired else factorize
    sol = solve_cr if paired else solve
    rho = T(rho); rb = rho_vec(rho)
    s["P"][N, 3:] = 1.0; fac(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
    al = T(alpha); status = 0
    x = np.zeros((L, 5), dt)
    vb = np.zeros((L, 5), dt); zb = np.zeros((L, 5), dt); rbd = np.zeros((L, 5), dt); rdy = np.zeros((L, 3), dt)
    uu = qq.copy()  # u = q + A'y
    for it in range(1, max_iter + 1):
        g = np.where(mask, -((P * x + uu) + At_apply(s, rd * rdy, rb * rbd)), 0).astype(dt)
        dl = (al * sol(s, g)).astype(dt)
        s1d, s1b = A_apply(s, dl)
        x = (x + dl).astype(dt); dx = dl
        wd = (al * rdy + s1d).astype(dt)
        rdy = (rdy + s1d).astype(dt)
        dyd = (rd * wd).astype(dt)
        wb = (al * rbd + s1b).astype(dt)
        vb = (vb + wb).astype(dt)
        zn = np.minimum(np.maximum(vb, lo), hi).astype(dt)
        stepb = (zn - zb).astype(dt); zb = zn
        rbd = ((rbd + s1b) - stepb).astype(dt)
        dyb = (rb * (wb - stepb)).astype(dt)
        uu = (uu + At_apply(s, dyd, dyb)).astype(dt)
        if it == 1:  # the dynamics z jumped 0 -> d
            rdy = (rdy - dd).astype(dt)
            dyd = (dyd - rd * dd).astype(dt)
            uu = (uu + At_apply(s, -rd * dd, np.zeros_like(dyb))).astype(dt)
        if it % check_termination == 0 or it % adaptive_rho_interval == 0:
            Axd, Axb = A_apply(s, x)
            rpd, rpb = Axd - dd, Axb - zb
            Px = P * x
            Aty = uu - qq
            rdual = np.where(mask, Px + uu, 0)
            pri_res = max(np.max(np.abs(rpd / Ed)), np.max(np.abs(rpb / Eb)))
            dua_res = np.max(np.abs(rdual / D)) / cs
            if it % check_termination == 0:
                nz_ = max(np.max(np.abs(dd / Ed)), np.max(np.abs(zb / Eb)))
                nax = max(np.max(np.abs(Axd / Ed)), np.max(np.abs(Axb / Eb)))
                eps_prim = eps_abs + eps_rel * max(nz_, nax)
                nd = max(np.max(np.abs(qq / D)), np.max(np.abs(Aty / D)), np.max(np.abs(Px / D))) / cs
                eps_dual = eps_abs + eps_rel * nd
                if pri_res < eps_prim and dua_res < eps_dual: status = 1; break
                if not (pri_res < eps_prim):
                    pyd = dyd
                    pyb = np.where(hi > thr, np.where(lo < -thr, 0, np.minimum(dyb, 0)), np.where(lo < -thr, np.maximum(dyb, 0), dyb))
                    ndy = max(np.max(np.abs(Ed * pyd)), np.max(np.abs(Eb * pyb)))
                    if ndy > eps_prim_inf:
                        lhs = np.sum(dd * pyd) + np.sum(hi * np.maximum(pyb, 0) + lo * np.minimum(pyb, 0))
                        if lhs < -eps_prim_inf * ndy:
                            Atdy = np.where(mask, At_apply(s, pyd, pyb), 0)
                            if np.max(np.abs(Atdy / D)) < eps_prim_inf * ndy: status = -3; break
            if it % adaptive_rho_interval == 0:
                pn = max(np.abs(rpd).max(), np.abs(rpb).max())
                pn /= max(np.max(np.abs(dd)), np.max(np.abs(zb)), np.max(np.abs(Axd)), np.max(np.abs(Axb))) + 1e-10
                dn = np.max(np.abs(rdual)); dn /= max(np.max(np.abs(qq)), np.max(np.abs(Aty)), np.max(np.abs(Px))) + 1e-10
                rnew = float(rho) * np.sqrt(pn / (dn + 1e-10)); rnew = min(max(rnew, RHO_MIN), RHO_MAX)
                if rnew > float(rho) * adaptive_rho_tolerance or rnew < float(rho) / adaptive_rho_tolerance:
                    rb_old = rb
                    rho = T(rnew); rb = rho_vec(rho); rd = T(RHO_EQ_OVER_RHO_INEQ * rho)
                    vb = (zb + (vb - zb) * (rb_old / rb)).astype(dt)
                    s["P"][N, 3:] = 1.0; fac(s, sigma, rho, rb); s["P"][N, 3:] = 0.0
    if status == 0: status = -2
    xs = (D * x).astype(np.float64)
    return dict(x=np.concatenate([xs[:N + 1, :3].ravel(), xs[:N, 3:].ravel()]), iter=it, status=status)


# ------------------------------------------------------------------------------------------------------
# paired-stage factorisation: one level of cyclic reduction inside the lane, then PCR across lanes
# ------------------------------------------------------------------------------------------------------
# Lane l holds stages A = 2l and B = 2l + 1.  After the inputs are eliminated (as in factorize()), the even stages are
# eliminated locally:  x_A = DA^-1 b_A - G x_B - H x_B(l-1)  with  G = DA^-1 U_A,  H = DA^-1 Lo_A;  the odd stages then
# form a block-tridiagonal chain of half the length,
#     D_B' = D_B - U_A' G - U_B H(l+1),   U_B' = -U_B G(l+1),   b_B' = b_B - G' b_A - [H' b_A](l+1),
# which PCR solves in log2(lanes) levels.  csrc/admm_pair.cuh is the CUDA transcription.
def _stage_blocks(s, sigma, rho, rho_b):
    dt = s["a"].dtype
    a, c, e, P = s["a"], s["c"], s["e"], s["P"]
    L = a.shape[0]
    rd = dt.type(RHO_EQ_OVER_RHO_INEQ * rho)
    diag = P + dt.type(sigma) + rho_b * e * e
    cn = down(c)
    Sxx = np.zeros((L, 3, 3), dt)
    Sxx[:, 0, 0] = diag[:, 0] + rd * (c[:, 0] ** 2 + a[:, 0] ** 2 + a[:, 2] ** 2 + a[:, 4] ** 2)
    Sxx[:, 1, 1] = diag[:, 1] + rd * (c[:, 1] ** 2 + a[:, 1] ** 2 + a[:, 3] ** 2)
    Sxx[:, 2, 2] = diag[:, 2] + rd * (c[:, 2] ** 2 + a[:, 5] ** 2)
    Sxx[:, 0, 1] = Sxx[:, 1, 0] = rd * (a[:, 0] * a[:, 1] + a[:, 2] * a[:, 3])
    Sxx[:, 0, 2] = Sxx[:, 2, 0] = rd * (a[:, 4] * a[:, 5])
    Svv = diag[:, 3] + rd * a[:, 7] ** 2
    Skk = diag[:, 4] + rd * a[:, 6] ** 2
    Sxv = np.stack([rd * a[:, 4] * a[:, 7], np.zeros(L, dt), rd * a[:, 5] * a[:, 7]], axis=1)
    Sxk = np.stack([rd * a[:, 2] * a[:, 6], rd * a[:, 3] * a[:, 6], np.zeros(L, dt)], axis=1)
    Fx = np.zeros((L, 3, 3), dt)
    Fx[:, 0, 0] = rd * a[:, 0] * cn[:, 0]; Fx[:, 1, 0] = rd * a[:, 1] * cn[:, 0]
    Fx[:, 0, 1] = rd * a[:, 2] * cn[:, 1]; Fx[:, 1, 1] = rd * a[:, 3] * cn[:, 1]
    Fx[:, 0, 2] = rd * a[:, 4] * cn[:, 2]; Fx[:, 2, 2] = rd * a[:, 5] * cn[:, 2]
    Fv = rd * a[:, 7] * cn[:, 2]
    Fk = rd * a[:, 6] * cn[:, 1]
    iv, ik = (1.0 / Svv).astype(dt), (1.0 / Skk).astype(dt)
    Dm = Sxx - iv[:, None, None] * Sxv[:, :, None] * Sxv[:, None, :] - ik[:, None, None] * Sxk[:, :, None] * Sxk[:, None, :]
    U = Fx.copy()
    U[:, :, 2] -= (iv * Fv)[:, None] * Sxv
    U[:, :, 1] -= (ik * Fk)[:, None] * Sxk
    addn = up(np.stack([np.zeros(L, dt), ik * Fk * Fk, iv * Fv * Fv], axis=1))
    for i in range(3):
        Dm[:, i, i] -= addn[:, i]
    U[-1] = 0
    return Dm.astype(dt), U.astype(dt), dict(iv=iv, ik=ik, Sxv=Sxv, Sxk=Sxk, Fv=Fv, Fk=Fk)


def factorize_cr(s, sigma, rho, rho_b):
    dt = s["a"].dtype
    Dm, U, f = _stage_blocks(s, sigma, rho, rho_b)
    L = Dm.shape[0]
    assert L % 2 == 0
    DA, DB, UA, UB = Dm[0::2], Dm[1::2], U[0::2], U[1::2]
    LoA = np.transpose(up(UB), (0, 2, 1))          # couples A_l to B_(l-1)
    DAi = inv3(DA)
    G = np.einsum("lij,ljk->lik", DAi, UA).astype(dt)
    H = np.einsum("lij,ljk->lik", DAi, LoA).astype(dt)
    Dr = (DB - np.einsum("lji,ljk->lik", UA, G) - np.einsum("lij,ljk->lik", UB, down(H))).astype(dt)
    Ur = (-np.einsum("lij,ljk->lik", UB, down(G))).astype(dt)
    Ur[-1] = 0
    Lr = np.transpose(up(Ur), (0, 2, 1)).copy()
    levels = []
    sft = 1
    n = L // 2
    while sft < n:
        Dinv = inv3(Dr)
        al = np.einsum("lij,ljk->lik", Lr, up(Dinv, sft))
        be = np.einsum("lij,ljk->lik", Ur, down(Dinv, sft))
        Dr = (Dr - np.einsum("lij,ljk->lik", al, up(Ur, sft)) - np.einsum("lij,ljk->lik", be, down(Lr, sft))).astype(dt)
        Lr, Ur = (-np.einsum("lij,ljk->lik", al, up(Lr, sft))).astype(dt), (-np.einsum("lij,ljk->lik", be, down(Ur, sft))).astype(dt)
        levels.append((al.astype(dt), be.astype(dt), sft))
        sft *= 2
    f.update(levels=levels, Dinv=inv3(Dr), DAi=DAi, G=G, H=H)
    s["fac"] = f


def solve_cr(s, b):
    f = s["fac"]
    bx = b[:, :3] - (f["iv"] * b[:, 3])[:, None] * f["Sxv"] - (f["ik"] * b[:, 4])[:, None] * f["Sxk"]
    tn = np.stack([np.zeros_like(b[:, 0]), f["ik"] * f["Fk"] * b[:, 4], f["iv"] * f["Fv"] * b[:, 3]], axis=1)
    bx = bx - up(tn)
    bA, bB = bx[0::2], bx[1::2]
    tH = np.einsum("lji,lj->li", f["H"], bA)
    br = bB - np.einsum("lji,lj->li", f["G"], bA) - down(tH)
    for al, be, sft in f["levels"]:
        br = br - np.einsum("lij,lj->li", al, up(br, sft)) - np.einsum("lij,lj->li", be, down(br, sft))
    xB = np.einsum("lij,lj->li", f["Dinv"], br)
    xA = np.einsum("lij,lj->li", f["DAi"], bA) - np.einsum("lij,lj->li", f["G"], xB) - np.einsum("lij,lj->li", f["H"], up(xB))
    x = np.empty_like(bx)
    x[0::2], x[1::2] = xA, xB
    xn = down(x)
    v = f["iv"] * (b[:, 3] - np.einsum("li,li->l", f["Sxv"], x) - f["Fv"] * xn[:, 2])
    k = f["ik"] * (b[:, 4] - np.einsum("li,li->l", f["Sxk"], x) - f["Fk"] * xn[:, 1])
    return np.concatenate([x, v[:, None], k[:, None]], axis=1).astype(b.dtype)
