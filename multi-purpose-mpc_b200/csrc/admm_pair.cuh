// admm_pair.cuh -- K1 + K2, paired-stage fp32 production path for sm_100a (N + 1 <= 2 * LPS stages).
//
// A scenario occupies a group of LPS lanes (8 / 16 / 32); lane l keeps TWO consecutive stages of the QP that
// MPC._init_problem builds (reference: src/MPC.py:61-159), A = 2l and B = 2l + 1, as the halves of float2
// registers.  Everything that is element-wise per stage -- the OSQP row updates, A x, A' y, the input
// elimination -- is then one packed instruction for both stages (FFMA2 / FADD2 / FMUL2, sm_100a's
// fma.rn.f32x2 family): half the issue slots of the lane-per-stage kernel in admm.cuh, and half the
// neighbour shuffles, because A -> B traffic stays inside the lane.
// Linear system: after the two inputs of every stage are eliminated (as in admm.cuh), the even stages are
// eliminated inside the lane (one level of cyclic reduction),
//     x_A = DA^-1 b_A - G x_B - H x_B(l-1),    G = DA^-1 U_A,  H = DA^-1 Lo_A,
// and the odd stages form a block-tridiagonal chain of LPS 3x3 blocks that parallel cyclic reduction solves
// in log2(LPS) levels; a PCR level multiplies the up- and the down-neighbour at once with packed
// (alpha, beta) coefficients, and the last level (one partner per lane) is merged.
// Iteration ("v-form", executable model: tools/admm_pcr_model.py::admm_vform, same iterates as
// oracle/osqp_oracle.c in exact arithmetic):
//     OSQP:  v+ = alpha z~ + (1 - alpha) z + y / rho,  z+ = clip(v+),  y+ = rho (v+ - z+).
//     Since (z, y) came out of the same projection, y / rho = v - z, hence v+ = v + w with
//     w = alpha (r + A D), r = A x - z tracked, D = x~ - x the solution of S D = -(P x + u + A'(rho r)),
//     u = q + A'y tracked through dy = rho (w - (z+ - z)).  A bound row is the triple (v, z, r); y is never
//     stored.  Dynamics rows are equalities whose z jumps from the cold start 0 to d in iteration 1 and
//     stays: they carry r only and iteration 1 is patched afterwards.
// Everything multiplied by rho_eq = 1e3 rho is a small residual, which is what lets fp32 reproduce OSQP's
// iteration counts and infeasibility certificates at the reference's tolerance (as the increment form did).
// LOOSE = the bounds on e_psi and t are infinite for every stage (the reference's configuration): OSQP gives
// such rows rho = 1e-6 and their z follows A x exactly, so the loop skips them (their 1e-6 e^2 stays in S).
#pragma once
#include "admm.cuh"

namespace mpcb {

typedef float2 f2;
__constant__ float2 kNegOne2 = {-1.0f, -1.0f};
__device__ __forceinline__ f2 mk(float a, float b) { return make_float2(a, b); }
__device__ __forceinline__ f2 bc(float a) { return make_float2(a, a); }
__device__ __forceinline__ f2 pfma(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 padd(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 pmul(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 psub(f2 a, f2 b) { return __ffma2_rn(b, kNegOne2, a); }
__device__ __forceinline__ f2 pabs(f2 a) { return mk(fabsf(a.x), fabsf(a.y)); }
__device__ __forceinline__ f2 pmax(f2 a, f2 b) { return mk(fmaxf(a.x, b.x), fmaxf(a.y, b.y)); }
__device__ __forceinline__ f2 pmin(f2 a, f2 b) { return mk(fminf(a.x, b.x), fminf(a.y, b.y)); }
// volatile 64-bit shared-memory load: keeps per-iteration operands in shared memory (the compiler would otherwise hoist
// the loop-invariant loads out of the ADMM loop and hold them in registers, i.e. spill them)
__device__ __forceinline__ f2 ldsv(const f2* p) {
    const unsigned long long v = *reinterpret_cast<const volatile unsigned long long*>(p);
    return mk(__uint_as_float((unsigned)v), __uint_as_float((unsigned)(v >> 32)));
}
__device__ __forceinline__ void amax(float& m, f2 v) { m = fmaxf(m, fmaxf(fabsf(v.x), fabsf(v.y))); }

// A group of LPS lanes = one scenario.  prev()/next() are cyclic inside the group: the wrap-around always
// lands on a stage whose coupling coefficients are zero (stage 2 * LPS - 1 is either padding or stage N,
// which has no successor), so no edge selects are needed.
template <int LPS> struct GroupComm {
    unsigned gmask;
    int gl, src_prev, src_next;
    __device__ __forceinline__ GroupComm() {
        const int lane = threadIdx.x & 31;
        gl = lane & (LPS - 1);
        gmask = LPS == 32 ? kFull : (((1u << (LPS & 31)) - 1u) << (lane & ~(LPS - 1)));
        src_prev = (gl - 1) & (LPS - 1);
        src_next = (gl + 1) & (LPS - 1);
    }
    // Every collective uses the constant full-warp mask (a run-time mask makes nvcc wrap each shuffle in a
    // WARPSYNC.COLLECTIVE + convergence barrier); the width / xor distance keeps the exchange inside the group.
    // Consequently all groups of a warp execute every collective together: control flow around them is made
    // warp-uniform by voting (see admm_solve2).
    __device__ __forceinline__ float prev(float v) const { return __shfl_sync(kFull, v, src_prev, LPS); }
    __device__ __forceinline__ float next(float v) const { return __shfl_sync(kFull, v, src_next, LPS); }
    __device__ __forceinline__ float up(float v, int s) const { return __shfl_up_sync(kFull, v, s, LPS); }
    __device__ __forceinline__ float dn(float v, int s) const { return __shfl_down_sync(kFull, v, s, LPS); }
    __device__ __forceinline__ float bfly(float v, int s) const { return __shfl_xor_sync(kFull, v, s, LPS); }
    // the value the previous / next STAGE holds
    __device__ __forceinline__ f2 to_next(f2 p) const { return mk(prev(p.y), p.x); }
    __device__ __forceinline__ f2 from_next(f2 p) const { return mk(p.y, next(p.x)); }
    __device__ __forceinline__ float max(float v) const {  // v >= 0
        if constexpr (LPS == 32) {
            return __uint_as_float(__reduce_max_sync(kFull, __float_as_uint(v)));
        } else {
#pragma unroll
            for (int s = LPS / 2; s > 0; s >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, s, LPS));
            return v;
        }
    }
    __device__ __forceinline__ float sum(float v) const {
#pragma unroll
        for (int s = LPS / 2; s > 0; s >>= 1) v += __shfl_xor_sync(kFull, v, s, LPS);
        return v;
    }
    __device__ __forceinline__ bool any(bool p) const { return (__ballot_sync(kFull, p) & gmask) != 0u; }
    static __device__ __forceinline__ bool warp_any(bool p) { return __any_sync(kFull, p) != 0; }
    static __device__ __forceinline__ bool warp_all(bool p) { return __all_sync(kFull, p) != 0; }
};

struct Stage2 {
    f2 a[8], c[3], e[5];
    f2 P[5], q[5];
    f2 d[3];
    f2 lo[5], hi[5];
    f2 D[5], Ed[3], Eb[5];
    float cs;
};

__device__ __forceinline__ void pack_stages(Stage2& s, const Stage<float>& A, const Stage<float>& B) {
#pragma unroll
    for (int i = 0; i < 8; ++i) s.a[i] = mk(A.a[i], B.a[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i) { s.c[i] = mk(A.c[i], B.c[i]); s.d[i] = mk(A.d[i], B.d[i]); }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        s.e[i] = mk(A.e[i], B.e[i]); s.P[i] = mk(A.P[i], B.P[i]); s.q[i] = mk(A.q[i], B.q[i]);
        s.lo[i] = mk(A.lo[i], B.lo[i]); s.hi[i] = mk(A.hi[i], B.hi[i]);
    }
}

__device__ __forceinline__ float limit_scaling_f(float v) {
    v = v < (float)kMinScaling ? 1.0f : v;
    return fminf(v, (float)kMaxScaling);
}
// MUFU.RSQ alone: rsqrtf() wraps it in a denormal range fix-up (FSETP + two predicated FMUL) that can never trigger here,
// the argument is clamped to [1e-4, 1e4] -- bit-identical result, 3 instructions fewer per value (78 per Ruiz pass)
__device__ __forceinline__ float rsq_normal(float v) {
    float r;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}
__device__ __forceinline__ f2 prsqrt_lim(f2 v) { return mk(rsq_normal(limit_scaling_f(v.x)), rsq_normal(limit_scaling_f(v.y))); }

// OSQP scale_data (Ruiz equilibration of the KKT matrix + cost scaling), paired layout.
template <int LPS>
__device__ __forceinline__ void ruiz_scale2(const GroupComm<LPS>& cm, Stage2& s, int iters, int nvar) {
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.D[i] = bc(1.0f); s.Eb[i] = bc(1.0f); }
#pragma unroll
    for (int i = 0; i < 3; ++i) s.Ed[i] = bc(1.0f);
    s.cs = 1.0f;
    const float inv_nvar = 1.0f / (float)nvar;
    for (int it = 0; it < iters; ++it) {
        f2 aa[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) aa[i] = pabs(s.a[i]);
        f2 col[5];
        col[0] = pmax(pmax(aa[0], aa[2]), aa[4]);
        col[1] = pmax(aa[1], aa[3]);
        col[2] = aa[5];
        col[3] = aa[7];
        col[4] = aa[6];
#pragma unroll
        for (int i = 0; i < 3; ++i) col[i] = pmax(col[i], pabs(s.c[i]));
#pragma unroll
        for (int i = 0; i < 5; ++i) col[i] = pmax(pmax(col[i], pabs(s.e[i])), pabs(s.P[i]));
        f2 ro[3];
        ro[0] = pmax(aa[0], aa[1]);
        ro[1] = pmax(pmax(aa[2], aa[3]), aa[6]);
        ro[2] = pmax(pmax(aa[4], aa[5]), aa[7]);
        f2 Dt[5], Edt[3], Ebt[5], En[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) Edt[i] = prsqrt_lim(pmax(pabs(s.c[i]), cm.to_next(ro[i])));
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            Dt[i] = prsqrt_lim(col[i]);
            Ebt[i] = prsqrt_lim(pabs(s.e[i]));
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) En[i] = cm.from_next(Edt[i]);
#pragma unroll
        for (int i = 0; i < 5; ++i) s.P[i] = pmul(pmul(s.P[i], Dt[i]), Dt[i]);
        s.a[0] = pmul(pmul(s.a[0], En[0]), Dt[0]); s.a[1] = pmul(pmul(s.a[1], En[0]), Dt[1]);
        s.a[2] = pmul(pmul(s.a[2], En[1]), Dt[0]); s.a[3] = pmul(pmul(s.a[3], En[1]), Dt[1]);
        s.a[4] = pmul(pmul(s.a[4], En[2]), Dt[0]); s.a[5] = pmul(pmul(s.a[5], En[2]), Dt[2]);
        s.a[6] = pmul(pmul(s.a[6], En[1]), Dt[4]); s.a[7] = pmul(pmul(s.a[7], En[2]), Dt[3]);
#pragma unroll
        for (int i = 0; i < 3; ++i) { s.c[i] = pmul(pmul(s.c[i], Edt[i]), Dt[i]); s.Ed[i] = pmul(s.Ed[i], Edt[i]); }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            s.e[i] = pmul(pmul(s.e[i], Ebt[i]), Dt[i]);
            s.q[i] = pmul(s.q[i], Dt[i]);
            s.D[i] = pmul(s.D[i], Dt[i]);
            s.Eb[i] = pmul(s.Eb[i], Ebt[i]);
        }
        // cost scaling
        float sp = 0.0f, mq = 0.0f;
#pragma unroll
        for (int i = 0; i < 5; ++i) { sp += fabsf(s.P[i].x) + fabsf(s.P[i].y); amax(mq, s.q[i]); }
        sp = cm.sum(sp) * inv_nvar;
        mq = limit_scaling_f(cm.max(mq));
        float ct = 1.0f / limit_scaling_f(fmaxf(sp, mq));
        const f2 ct2 = bc(ct);
#pragma unroll
        for (int i = 0; i < 5; ++i) { s.P[i] = pmul(s.P[i], ct2); s.q[i] = pmul(s.q[i], ct2); }
        s.cs *= ct;
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) s.d[i] = pmul(s.d[i], s.Ed[i]);
#pragma unroll
    for (int i = 0; i < 5; ++i) { s.lo[i] = pmul(s.lo[i], s.Eb[i]); s.hi[i] = pmul(s.hi[i], s.Eb[i]); }
}

// PCR coefficients of levels 0 .. NLEV-2 ((-alpha, -beta) packed, 9 float2 per level): 54 registers at LPS = 16 that ptxas
// cannot keep next to the iterate and parks in local memory (33 LDL.64 per pass).  MPC_PCR_COEF_SMEM keeps them in shared
// memory instead, one float4 column per lane ([k][32 lanes], conflict-free), and a pass re-reads them with LDS.128: the
// same bytes through the LSU, less than half the instructions.
// With the PCR coefficients out of the register file, the per-pass operands P, lo, hi fit back into registers
// (measured at 4096 cars, step with L2 flushed: 0.1590 ms all-LDL/LDS -> 0.1580 coefficients in smem -> 0.1556 with both
// switches on; bit-identical results).
#ifndef MPC_PASS_P_REG
#define MPC_PASS_P_REG 1
#endif
#ifndef MPC_PASS_BOUNDS_REG
#define MPC_PASS_BOUNDS_REG 1
#endif
#ifndef MPC_PCR_COEF_SMEM
#define MPC_PCR_COEF_SMEM 1
#endif
// Compensated curvature bound row: its v (and with it z = clip(v)) is kept as an unevaluated fp32 sum (high, low).  v is an
// O(1) number (|kappa| reaches 6.47) that collects one small increment w per pass; rounding that sum to 24 bits perturbs z
// by 6e-8 |z| every pass, and the QP's weakly determined directions (the curvature inputs: reduced-Hessian eigenvalues
// ~ 1e-7, tools/precision_study.py) integrate the perturbation.  TwoSum keeps the rounding error of v + w in the low word,
// z inherits it while the row is inactive, and r / dy are updated with both words of z+ - z.  Measured on the 66 solved
// golden QPs (executable model): max |x - oracle| 1.17e-3 -> 5.0e-4, the same as keeping every v, z in fp64; compensating
// the other bound rows as well changes nothing (4.9e-4), so only row 4 pays for it.
#ifndef MPC_COMPENSATED_V
#define MPC_COMPENSATED_V 1
#endif
// y-form (A/B switch, OFF): the duals are state -- y_d of the dynamics rows is accumulated (y_d += dy_d), y_b of a bound row is
// rho (v - z) by construction -- and the right-hand side is  P x + q + A'(y + rho r)  with ONE A' product per pass, instead of
// tracking u = q + A'y through a second product u += A'dy.  Measured on the GPU (round 2): on QPs that OSQP solves it is
// 3.5x closer to the fp64 oracle (4096 random QPs, identical traces: max |x - oracle| 7.7e-4 -> 2.2e-4; 76 golden: 3.5e-4 ->
// 1.4e-4), but it is NOT faster (the second product sits off the dependent chain: 116.9 vs 117.7 us) and it degrades on
// primal-infeasible QPs: there y grows without bound while A'y stays bounded, so A'y computed from y loses what u, fed by the
// vanishing increments A'dy, keeps -- certificates come a check apart 8x more often (0.07 % -> 0.76 % of the QPs) and one
// infeasible QP in 741 is never certified (status 2 instead of -3, i.e. the controller would apply an iterate where the
// reference replays its previous plan).  The status is what the reference reacts to, so the u-form stays.
#ifndef MPC_Y_FORM
#define MPC_Y_FORM 0
#endif
#ifndef MPC_PASS_Q_REG
#define MPC_PASS_Q_REG 0
#endif
template <int LPS> struct PcrCoef {  // float4 per lane
    static constexpr int kF4 = (9 * ((LPS == 32 ? 5 : (LPS == 16 ? 4 : (LPS == 8 ? 3 : 2))) - 1) + 1) / 2;
};
__device__ __forceinline__ float4 lds128v(const float4* p) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "r"((unsigned)__cvta_generic_to_shared(p)));
    return v;
}

template <int LPS> struct PairFactor {
    static constexpr int NLEV = LPS == 32 ? 5 : (LPS == 16 ? 4 : (LPS == 8 ? 3 : 2));
    f2 iv, ik, nsxv0, nsxv2, nsxk0, nsxk1, nfv, nfk;  // input elimination (per stage); couplings stored negated
    float UA[6], LA[6], DAi[6];                 // in-lane cyclic-reduction level: nonzeros of U_A (0 1 2 3 4 8), of Lo_A
                                                // (= U_B(l-1)': same six, transposed), DA^-1 (symmetric)
#if !MPC_PCR_COEF_SMEM
    f2 nab[NLEV - 1][9];                        // PCR levels 0 .. NLEV-2: (-alpha, -beta) packed
#endif
    float last[9];                              // PCR level NLEV-1: one partner (gl ^ LPS/2)
    float Dinv[6];                              // symmetric: 00 01 02 11 12 22
};

__device__ __forceinline__ void inv3sym6(const float* M /*00 01 02 11 12 22*/, float* R) {
    const float a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5];
    const float A = d * f - e * e, B = c * e - b * f, C = b * e - c * d;
    const float r = 1.0f / (a * A + b * B + c * C);
    R[0] = A * r; R[1] = B * r; R[2] = C * r;
    R[3] = (a * f - c * c) * r; R[4] = (b * c - a * e) * r; R[5] = (a * d - b * b) * r;
}
__device__ __forceinline__ void sym6_to9(const float* S, float* M) {
    M[0] = S[0]; M[1] = M[3] = S[1]; M[2] = M[6] = S[2]; M[4] = S[3]; M[5] = M[7] = S[4]; M[8] = S[5];
}

// Build S = P + sigma I + A' R A for both stages of the lane, eliminate the inputs, eliminate stage A,
// PCR-factorise the chain of B stages.
template <int LPS, bool LOOSE>
__device__ __forceinline__ void factorize2(const GroupComm<LPS>& cm, const Stage2& s, PairFactor<LPS>& f, float sigma,
                                           float rdf, const f2 rb[5], const f2* sm, float4* cf) {
    constexpr int NLEV = PairFactor<LPS>::NLEV;
    const f2* a = s.a;
    const f2 rd = bc(rdf), sg = bc(sigma);
    f2 diag[5];
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        const bool lz = LOOSE && (i == 1 || i == 2);  // loose rows: rho = rho_min, e kept in shared memory
        const f2 ei = lz ? sm[(34 + i) * LPS + cm.gl] : s.e[i];
        const f2 ri = lz ? bc((float)kRhoMin) : rb[i];
        diag[i] = pfma(pmul(ri, ei), ei, padd(sm[(49 + i) * LPS + cm.gl], sg));
    }
    f2 cn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) cn[i] = cm.from_next(s.c[i]);
    f2 D00 = pfma(rd, pfma(a[4], a[4], pfma(a[2], a[2], pfma(a[0], a[0], pmul(s.c[0], s.c[0])))), diag[0]);
    f2 D11 = pfma(rd, pfma(a[3], a[3], pfma(a[1], a[1], pmul(s.c[1], s.c[1]))), diag[1]);
    f2 D22 = pfma(rd, pfma(a[5], a[5], pmul(s.c[2], s.c[2])), diag[2]);
    f2 D01 = pmul(rd, pfma(a[2], a[3], pmul(a[0], a[1])));
    f2 D02 = pmul(rd, pmul(a[4], a[5]));
    const f2 Svv = pfma(rd, pmul(a[7], a[7]), diag[3]);
    const f2 Skk = pfma(rd, pmul(a[6], a[6]), diag[4]);
    f.iv = mk(1.0f / Svv.x, 1.0f / Svv.y);
    f.ik = mk(1.0f / Skk.x, 1.0f / Skk.y);
    const f2 ra7 = pmul(rd, a[7]), ra6 = pmul(rd, a[6]);
    const f2 sxv0 = pmul(ra7, a[4]), sxv2 = pmul(ra7, a[5]);
    const f2 sxk0 = pmul(ra6, a[2]), sxk1 = pmul(ra6, a[3]);
    const f2 fv = pmul(ra7, cn[2]), fk = pmul(ra6, cn[1]);
    // coupling block (row j, col j+1): U[3 i + r] = rd * (coef of x_i in row r of block j+1) * c_{j+1}[r]
    const f2 rc0 = pmul(rd, cn[0]), rc1 = pmul(rd, cn[1]), rc2 = pmul(rd, cn[2]);
    f2 U0 = pmul(a[0], rc0), U1 = pmul(a[2], rc1), U2 = pmul(a[4], rc2);
    f2 U3 = pmul(a[1], rc0), U4 = pmul(a[3], rc1), U8 = pmul(a[5], rc2);  // U5 = U6 = U7 = 0
    // Schur complement of the (diagonal) input block
    const f2 ivs0 = pmul(f.iv, sxv0), ivs2 = pmul(f.iv, sxv2), iks0 = pmul(f.ik, sxk0), iks1 = pmul(f.ik, sxk1);
    D00 = psub(D00, pfma(iks0, sxk0, pmul(ivs0, sxv0)));
    D01 = psub(D01, pmul(iks0, sxk1));
    D02 = psub(D02, pmul(ivs0, sxv2));
    D11 = psub(D11, pmul(iks1, sxk1));
    D22 = psub(D22, pmul(ivs2, sxv2));
    U2 = psub(U2, pmul(ivs0, fv)); U8 = psub(U8, pmul(ivs2, fv));
    U1 = psub(U1, pmul(iks0, fk)); U4 = psub(U4, pmul(iks1, fk));
    D11 = psub(D11, cm.to_next(pmul(pmul(f.ik, fk), fk)));
    D22 = psub(D22, cm.to_next(pmul(pmul(f.iv, fv), fv)));
    {
        const f2 m1 = bc(-1.0f);
        f.nsxv0 = pmul(sxv0, m1); f.nsxv2 = pmul(sxv2, m1); f.nsxk0 = pmul(sxk0, m1); f.nsxk1 = pmul(sxk1, m1);
        f.nfv = pmul(fv, m1); f.nfk = pmul(fk, m1);
    }
    // ---- eliminate stage A inside the lane ----
    {
        const float DA[6] = {D00.x, D01.x, D02.x, D11.x, 0.0f, D22.x};
        inv3sym6(DA, f.DAi);
    }
    float Di9[9];
    sym6_to9(f.DAi, Di9);
    const float UA[9] = {U0.x, U1.x, U2.x, U3.x, U4.x, 0.0f, 0.0f, 0.0f, U8.x};
    const float UB[9] = {U0.y, U1.y, U2.y, U3.y, U4.y, 0.0f, 0.0f, 0.0f, U8.y};
    float LoA[9];  // (U_B of the previous lane)'
    {
        const float p0 = cm.prev(UB[0]), p1 = cm.prev(UB[1]), p2 = cm.prev(UB[2]), p3 = cm.prev(UB[3]),
                    p4 = cm.prev(UB[4]), p8 = cm.prev(UB[8]);
        LoA[0] = p0; LoA[1] = p3; LoA[2] = 0.0f;
        LoA[3] = p1; LoA[4] = p4; LoA[5] = 0.0f;
        LoA[6] = p2; LoA[7] = 0.0f; LoA[8] = p8;
    }
    float G[9], H[9];  // G = DA^-1 U_A, H = DA^-1 Lo_A
    mm3(Di9, UA, G);
    mm3(Di9, LoA, H);
    f.UA[0] = UA[0]; f.UA[1] = UA[1]; f.UA[2] = UA[2]; f.UA[3] = UA[3]; f.UA[4] = UA[4]; f.UA[5] = UA[8];
    f.LA[0] = LoA[0]; f.LA[1] = LoA[3]; f.LA[2] = LoA[6]; f.LA[3] = LoA[1]; f.LA[4] = LoA[4]; f.LA[5] = LoA[8];
    float Hn[9], Gn[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { Hn[i] = cm.next(H[i]); Gn[i] = cm.next(G[i]); }
    float Dm[9], U[9], Lo[9];
    {
        const float DB[9] = {D00.y, D01.y, D02.y, D01.y, D11.y, 0.0f, D02.y, 0.0f, D22.y};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float acc = DB[3 * i + k];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc -= UA[3 * j + i] * G[3 * j + k] + UB[3 * i + j] * Hn[3 * j + k];
                Dm[3 * i + k] = acc;
            }
        float t[9];
        mm3(UB, Gn, t);
#pragma unroll
        for (int i = 0; i < 9; ++i) U[i] = -t[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) Lo[3 * i + k] = cm.prev(U[3 * k + i]);
    // ---- PCR over the B chain ----
#pragma unroll
    for (int lev = 0; lev < NLEV; ++lev) {
        const int sft = 1 << lev;
        float Di[9];
        inv3sym(Dm, Di);
        float Dup[9], Ddn[9], Uup[9], Ldn[9], Lup[9], Udn[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Dup[i] = cm.up(Di[i], sft); Ddn[i] = cm.dn(Di[i], sft);
            Uup[i] = cm.up(U[i], sft);  Ldn[i] = cm.dn(Lo[i], sft);
            Lup[i] = cm.up(Lo[i], sft); Udn[i] = cm.dn(U[i], sft);
        }
        const bool has_up = cm.gl >= sft, has_dn = cm.gl + sft < LPS;
        float al[9], be[9];
        mm3(Lo, Dup, al);
        mm3(U, Ddn, be);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            if (!has_up) al[i] = 0.0f;
            if (!has_dn) be[i] = 0.0f;
        }
        float t1[9], t2[9];
        mm3(al, Uup, t1);
        mm3(be, Ldn, t2);
#pragma unroll
        for (int i = 0; i < 9; ++i) Dm[i] -= t1[i] + t2[i];
        mm3(al, Lup, t1);
        mm3(be, Udn, t2);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            Lo[i] = -t1[i]; U[i] = -t2[i];
#if MPC_PCR_COEF_SMEM
            if (lev < NLEV - 1) {
                const int j = 9 * lev + i;  // float2 index: float4 column j / 2 of this lane, half j & 1
                reinterpret_cast<f2*>(cf + (j >> 1) * 32)[j & 1] = mk(-al[i], -be[i]);
            }
#else
            if (lev < NLEV - 1) f.nab[lev < NLEV - 1 ? lev : 0][i] = mk(-al[i], -be[i]);
#endif
            else f.last[i] = has_up ? al[i] : be[i];
        }
    }
    float Di[9];
    inv3sym(Dm, Di);
    f.Dinv[0] = Di[0]; f.Dinv[1] = Di[1]; f.Dinv[2] = Di[2]; f.Dinv[3] = Di[4]; f.Dinv[4] = Di[5]; f.Dinv[5] = Di[8];
}

// x = S^-1 b for both stages of the lane
template <int LPS>
__device__ __forceinline__ void kkt_solve2(const GroupComm<LPS>& cm, const PairFactor<LPS>& f, const f2 b[5], f2 x[5],
                                           const float4* cf) {
    constexpr int NLEV = PairFactor<LPS>::NLEV;
#if MPC_PCR_COEF_SMEM
    float4 cq[PcrCoef<LPS>::kF4 > 0 ? PcrCoef<LPS>::kF4 : 1];
#pragma unroll
    for (int k = 0; k < PcrCoef<LPS>::kF4; ++k) cq[k] = lds128v(cf + k * 32);
#endif
    const f2 bv = pmul(f.iv, b[3]), bk = pmul(f.ik, b[4]);
    f2 bx0 = pfma(bk, f.nsxk0, pfma(bv, f.nsxv0, b[0]));
    f2 bx1 = pfma(bk, f.nsxk1, b[1]);
    f2 bx2 = pfma(bv, f.nsxv2, b[2]);
    bx1 = padd(bx1, cm.to_next(pmul(bk, f.nfk)));
    bx2 = padd(bx2, cm.to_next(pmul(bv, f.nfv)));
    // in-lane cyclic-reduction level: t = DA^-1 b_A,  b_B' = b_B - U_A' t - [Lo_A' t](l+1)
    const float bA0 = bx0.x, bA1 = bx1.x, bA2 = bx2.x;
    const float t0 = fmaf(f.DAi[2], bA2, fmaf(f.DAi[1], bA1, f.DAi[0] * bA0));
    const float t1 = fmaf(f.DAi[4], bA2, fmaf(f.DAi[3], bA1, f.DAi[1] * bA0));
    const float t2 = fmaf(f.DAi[5], bA2, fmaf(f.DAi[4], bA1, f.DAi[2] * bA0));
    // U_A = [u0 u1 u2; u3 u4 0; 0 0 u5],  Lo_A = [l0 l3 0; l1 l4 0; l2 0 l5]   (indices into f.UA / f.LA)
    const float h0 = fmaf(f.LA[2], t2, fmaf(f.LA[1], t1, f.LA[0] * t0));
    const float h1 = fmaf(f.LA[4], t1, f.LA[3] * t0);
    const float h2 = f.LA[5] * t2;
    f2 R0 = mk(fmaf(-f.UA[3], t1, fmaf(-f.UA[0], t0, bx0.y)), 0.0f);
    f2 R1 = mk(fmaf(-f.UA[4], t1, fmaf(-f.UA[1], t0, bx1.y)), 0.0f);
    f2 R2 = mk(fmaf(-f.UA[5], t2, fmaf(-f.UA[2], t0, bx2.y)), 0.0f);
    R0.x -= cm.next(h0); R1.x -= cm.next(h1); R2.x -= cm.next(h2);
    // PCR: one packed FMA chain per row accumulates r - alpha.up - beta.down as (r - alpha.up, -beta.down)
#pragma unroll
    for (int lev = 0; lev < NLEV - 1; ++lev) {
        const int sft = 1 << lev;
        const f2 n0 = mk(cm.up(R0.x, sft), cm.dn(R0.x, sft));
        const f2 n1 = mk(cm.up(R1.x, sft), cm.dn(R1.x, sft));
        const f2 n2 = mk(cm.up(R2.x, sft), cm.dn(R2.x, sft));
#if MPC_PCR_COEF_SMEM
        f2 nab[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = 9 * lev + i;
            nab[i] = (j & 1) ? mk(cq[j >> 1].z, cq[j >> 1].w) : mk(cq[j >> 1].x, cq[j >> 1].y);
        }
#else
        const f2* nab = f.nab[lev];
#endif
        const f2 s0 = pfma(nab[2], n2, pfma(nab[1], n1, pfma(nab[0], n0, R0)));
        const f2 s1 = pfma(nab[5], n2, pfma(nab[4], n1, pfma(nab[3], n0, R1)));
        const f2 s2 = pfma(nab[8], n2, pfma(nab[7], n1, pfma(nab[6], n0, R2)));
        R0.x = s0.x + s0.y; R1.x = s1.x + s1.y; R2.x = s2.x + s2.y;
    }
    float r0 = R0.x, r1 = R1.x, r2 = R2.x;
    {
        const int sft = LPS / 2;
        const float n0 = cm.bfly(r0, sft), n1 = cm.bfly(r1, sft), n2 = cm.bfly(r2, sft);
        r0 = fmaf(-f.last[2], n2, fmaf(-f.last[1], n1, fmaf(-f.last[0], n0, r0)));
        r1 = fmaf(-f.last[5], n2, fmaf(-f.last[4], n1, fmaf(-f.last[3], n0, r1)));
        r2 = fmaf(-f.last[8], n2, fmaf(-f.last[7], n1, fmaf(-f.last[6], n0, r2)));
    }
    const float xB0 = fmaf(f.Dinv[2], r2, fmaf(f.Dinv[1], r1, f.Dinv[0] * r0));
    const float xB1 = fmaf(f.Dinv[4], r2, fmaf(f.Dinv[3], r1, f.Dinv[1] * r0));
    const float xB2 = fmaf(f.Dinv[5], r2, fmaf(f.Dinv[4], r1, f.Dinv[2] * r0));
    // back-substitution: x_A = t - DA^-1 (U_A x_B + Lo_A x_B(l-1))
    const float p0 = cm.prev(xB0), p1 = cm.prev(xB1), p2 = cm.prev(xB2);
    const float w0 = fmaf(f.LA[3], p1, fmaf(f.LA[0], p0, fmaf(f.UA[2], xB2, fmaf(f.UA[1], xB1, f.UA[0] * xB0))));
    const float w1 = fmaf(f.LA[4], p1, fmaf(f.LA[1], p0, fmaf(f.UA[4], xB1, f.UA[3] * xB0)));
    const float w2 = fmaf(f.LA[5], p2, fmaf(f.LA[2], p0, f.UA[5] * xB2));
    const float xA0 = fmaf(-f.DAi[2], w2, fmaf(-f.DAi[1], w1, fmaf(-f.DAi[0], w0, t0)));
    const float xA1 = fmaf(-f.DAi[4], w2, fmaf(-f.DAi[3], w1, fmaf(-f.DAi[1], w0, t1)));
    const float xA2 = fmaf(-f.DAi[5], w2, fmaf(-f.DAi[4], w1, fmaf(-f.DAi[2], w0, t2)));
    x[0] = mk(xA0, xB0); x[1] = mk(xA1, xB1); x[2] = mk(xA2, xB2);
    const f2 xn1 = cm.from_next(x[1]), xn2 = cm.from_next(x[2]);  // fv, fk are 0 where there is no successor
    x[3] = pmul(f.iv, pfma(f.nfv, xn2, pfma(f.nsxv2, x[2], pfma(f.nsxv0, x[0], b[3]))));
    x[4] = pmul(f.ik, pfma(f.nfk, xn1, pfma(f.nsxk1, x[1], pfma(f.nsxk0, x[0], b[4]))));
}

// z = A w for both stages: zd (dynamics block of the stage) and zb (bound rows)
template <int LPS, bool LOOSE>
__device__ __forceinline__ void A_apply2(const GroupComm<LPS>& cm, const Stage2& s, const f2 w[5], f2 zd[3], f2 zb[5]) {
    const f2 o0 = pfma(s.a[1], w[1], pmul(s.a[0], w[0]));
    const f2 o1 = pfma(s.a[6], w[4], pfma(s.a[3], w[1], pmul(s.a[2], w[0])));
    const f2 o2 = pfma(s.a[7], w[3], pfma(s.a[5], w[2], pmul(s.a[4], w[0])));
    zd[0] = pfma(s.c[0], w[0], cm.to_next(o0));
    zd[1] = pfma(s.c[1], w[1], cm.to_next(o1));
    zd[2] = pfma(s.c[2], w[2], cm.to_next(o2));
#pragma unroll
    for (int i = 0; i < 5; ++i)
        if (!(LOOSE && (i == 1 || i == 2))) zb[i] = pmul(s.e[i], w[i]);
}

// r = acc + A' y for both stages (LOOSE: yb[1], yb[2] are identically zero and not read)
template <int LPS, bool LOOSE>
__device__ __forceinline__ void At_apply2(const GroupComm<LPS>& cm, const Stage2& s, const f2 yd[3], const f2 yb[5],
                                          const f2 acc[5], f2 r[5]) {
    const f2 g0 = cm.from_next(yd[0]), g1 = cm.from_next(yd[1]), g2 = cm.from_next(yd[2]);
    r[0] = pfma(s.a[4], g2, pfma(s.a[2], g1, pfma(s.a[0], g0, pfma(s.c[0], yd[0], pfma(s.e[0], yb[0], acc[0])))));
    r[1] = pfma(s.a[3], g1, pfma(s.a[1], g0, pfma(s.c[1], yd[1], acc[1])));
    r[2] = pfma(s.a[5], g2, pfma(s.c[2], yd[2], acc[2]));
    if (!LOOSE) { r[1] = pfma(s.e[1], yb[1], r[1]); r[2] = pfma(s.e[2], yb[2], r[2]); }
    r[3] = pfma(s.a[7], g2, pfma(s.e[3], yb[3], acc[3]));
    r[4] = pfma(s.a[6], g1, pfma(s.e[4], yb[4], acc[4]));
}

__device__ __forceinline__ float rho_of(float lo, float hi, float rho, float thr) {
    return (lo < -thr && hi > thr) ? (float)kRhoMin : ((hi - lo < (float)kRhoTol) ? (float)kRhoEqOverIneq * rho : rho);
}
template <int LPS, bool LOOSE>
__device__ __forceinline__ void set_rho2(const f2* sm, int gl, float rho, float& rd, f2 rb[5]) {
    const float thr = (float)(kOsqpInfty * kMinScaling);
    rd = (float)kRhoEqOverIneq * rho;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (LOOSE && (i == 1 || i == 2)) continue;
        const f2 lo = sm[(39 + i) * LPS + gl], hi = sm[(44 + i) * LPS + gl];
        rb[i] = mk(rho_of(lo.x, hi.x, rho, thr), rho_of(lo.y, hi.y, rho, thr));
    }
}

// per-scenario shared constants, [kPairRows][LPS] float2 -- what only the termination checks, the certificates and the
// (re)factorisation read, so that it does not occupy registers across the iteration:
//   0..2 d | 3..7 D | 8..10 Ed | 11..15 Eb | 16..20 1/D | 21..23 1/Ed | 24..28 1/Eb | 29..33 q | 34..38 e | 39..43 lo | 44..48 hi
// and what the iteration reads once per pass (one LDS.64 each instead of a register pair held for the whole solve):
//   49..53 P | 54 (|q|_scaled, |q|_unscaled) | 55 (c, 1/c)
// and what the pass in front of a termination check leaves for the certificates (so that it is not carried in registers
// from pass to pass):   56..60 alpha D (dx) | 61..63 dy of the dynamics rows | 64..68 dy of the bound rows
constexpr int kPairRows = 69;

// -DMPC_PHASE_CLOCK: a diagnostic build in which lane 0 of every 128th CTA prints the SM clock at the phase boundaries of
// its solve (kernel entry, assembled, scaled, factorised, first result emitted, loop left).  One warp per CTA assumed.
#ifdef MPC_PHASE_CLOCK
__device__ __forceinline__ long long* phase_clock_buf() { __shared__ long long t[8]; return t; }
#define MPC_PHASE_MARK(k) do { if ((threadIdx.x & 31) == 0) phase_clock_buf()[k] = clock64(); } while (0)
#else
#define MPC_PHASE_MARK(k) do { } while (0)
#endif

// The OSQP loop.  ALL lanes of the warp call this together (every group = one scenario).  Control flow around
// the collectives is warp-uniform: a branch that only some scenarios need is taken by the whole warp when ANY
// scenario votes for it and its result is ignored elsewhere (a refactorisation with an unchanged rho reproduces
// the same factor bit for bit).  A scenario that terminates hands its result to `emit(w, result)` at once --
// w[5] = the UNSCALED primal stage vectors, NaN when OSQP would return no solution -- and then keeps iterating as
// a bystander until every scenario of the warp is done; `live` = false marks a group without a scenario.
// unroll factor of the quiet-pass loop.  Measured at 4096 cars (step, L2 flushed): 1 -> 121.8 us, 2 -> 124.0, 3 -> 128.0: the
// register copies at the loop's tail that unrolling removes cost less than the instruction fetches it adds.
#ifndef MPC_QUIET_UNROLL
#define MPC_QUIET_UNROLL 1
#endif
constexpr int kQuietUnroll = MPC_QUIET_UNROLL;
template <int LPS, bool LOOSE, typename Emit>
__device__ __forceinline__ void admm_solve2(const GroupComm<LPS>& cm, Stage2& s, const AdmmSettings& st, const f2 al2,
                                            const f2 nal2, int nvar, f2* sm, float4* cf, bool live, Emit emit) {
    typedef GroupComm<LPS> GC;
    const int gl = cm.gl;
    MPC_PHASE_MARK(1);
    if (st.scaling > 0) ruiz_scale2<LPS>(cm, s, st.scaling, nvar);
    else {
#pragma unroll
        for (int i = 0; i < 5; ++i) { s.D[i] = bc(1.0f); s.Eb[i] = bc(1.0f); }
#pragma unroll
        for (int i = 0; i < 3; ++i) s.Ed[i] = bc(1.0f);
        s.cs = 1.0f;
    }
    const float thr = (float)(kOsqpInfty * kMinScaling);
    MPC_PHASE_MARK(2);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        sm[i * LPS + gl] = s.d[i]; sm[(8 + i) * LPS + gl] = s.Ed[i];
        sm[(21 + i) * LPS + gl] = mk(1.0f / s.Ed[i].x, 1.0f / s.Ed[i].y);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        sm[(3 + i) * LPS + gl] = s.D[i]; sm[(11 + i) * LPS + gl] = s.Eb[i];
        sm[(16 + i) * LPS + gl] = mk(1.0f / s.D[i].x, 1.0f / s.D[i].y);
        sm[(24 + i) * LPS + gl] = mk(1.0f / s.Eb[i].x, 1.0f / s.Eb[i].y);
        sm[(29 + i) * LPS + gl] = s.q[i]; sm[(34 + i) * LPS + gl] = s.e[i];
        sm[(39 + i) * LPS + gl] = s.lo[i]; sm[(44 + i) * LPS + gl] = s.hi[i]; sm[(49 + i) * LPS + gl] = s.P[i];
    }
    float rho = (float)st.rho, rdf;
    f2 rb[5];
    const float sigma = (float)st.sigma;
    set_rho2<LPS, LOOSE>(sm, gl, rho, rdf, rb);
    PairFactor<LPS> f;
    factorize2<LPS, LOOSE>(cm, s, f, sigma, rdf, rb, sm, cf);
    MPC_PHASE_MARK(3);
    float nq_s = 0.0f, nq_u = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i) { amax(nq_s, s.q[i]); amax(nq_u, pmul(s.q[i], sm[(16 + i) * LPS + gl])); }
    sm[54 * LPS + gl] = mk(cm.max(nq_s), cm.max(nq_u));
    sm[55 * LPS + gl] = mk(s.cs, 1.0f / s.cs);
    const f2 zero = bc(0.0f);
    f2 x[5], u[5], yd[3], vb[5], zb[5], rbd[5], rdy[3];   // u: u-form only; yd: y-form only
    f2 vl4 = zero, zl4 = zero;  // low words of v and z of the curvature row (MPC_COMPENSATED_V)
#pragma unroll
    for (int i = 0; i < 5; ++i) { x[i] = zero; u[i] = s.q[i]; vb[i] = zero; zb[i] = zero; rbd[i] = zero; }
#pragma unroll
    for (int i = 0; i < 3; ++i) { rdy[i] = zero; yd[i] = zero; }
    f2 rd = bc(rdf);
    bool done = !live;
    int iter = 0;
    int chk = st.check_termination > 0 ? st.check_termination : -1;
    int adp = st.adaptive_rho_interval > 0 ? st.adaptive_rho_interval : -1;
    auto finish = [&](int status, int it) {
        f2 w[5];
        const bool nan_out = (status == -3 || status == -4 || status == -7 || status == 3 || status == 4);  // OSQP: no solution
#pragma unroll
        for (int i = 0; i < 5; ++i) w[i] = nan_out ? bc(NAN) : pmul(sm[(3 + i) * LPS + gl], x[i]);
        SolveResult r;
        r.iters = it;
        r.status = status;
        MPC_PHASE_MARK(4);
        emit(w, r);
        MPC_PHASE_MARK(5);
        done = true;
    };
    // one ADMM pass
    auto pass = [&](const bool first, const bool keep) __attribute__((always_inline)) {
        f2 td[3], tb[5], rhs[5], s1d[3], s1b[5], dl[5], ed[3], eb[5];
#pragma unroll
        for (int i = 0; i < 3; ++i) td[i] = MPC_Y_FORM ? pfma(rd, rdy[i], yd[i]) : pmul(rd, rdy[i]);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const f2 lin = MPC_Y_FORM ? (MPC_PASS_Q_REG ? s.q[i] : ldsv(&sm[(29 + i) * LPS + gl])) : u[i];
            rhs[i] = pfma(MPC_PASS_P_REG ? s.P[i] : ldsv(&sm[(49 + i) * LPS + gl]), x[i], lin);
            if (LOOSE && (i == 1 || i == 2)) continue;
            if (MPC_Y_FORM) {  // y + rho r = rho ((v - z) + r)
                f2 vz = psub(vb[i], zb[i]);
                if (MPC_COMPENSATED_V && i == 4) vz = padd(vz, psub(vl4, zl4));
                tb[i] = pmul(rb[i], padd(vz, rbd[i]));
            } else {
                tb[i] = pmul(rb[i], rbd[i]);
            }
        }
        At_apply2<LPS, LOOSE>(cm, s, td, tb, rhs, rhs);  // rhs = P x + q + A'(y + rho r);  S D = -rhs
        kkt_solve2<LPS>(cm, f, rhs, dl, cf);
#pragma unroll
        for (int i = 0; i < 5; ++i) { dl[i] = pmul(dl[i], nal2); x[i] = padd(x[i], dl[i]); }  // dl = alpha D
        A_apply2<LPS, LOOSE>(cm, s, dl, s1d, s1b);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const f2 wv = pfma(al2, rdy[i], s1d[i]);  // v - z_prev
            rdy[i] = padd(rdy[i], s1d[i]);
            ed[i] = pmul(rd, wv);                     // dy of the dynamics rows
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            if (LOOSE && (i == 1 || i == 2)) continue;
            const f2 wv = pfma(al2, rbd[i], s1b[i]);
            if (MPC_COMPENSATED_V && i == 4) {
                const f2 vs = padd(vb[i], wv), bb = psub(vs, vb[i]);             // TwoSum(v, w)
                vl4 = padd(vl4, padd(psub(vb[i], psub(vs, bb)), psub(wv, bb)));
                vb[i] = vs;
            } else {
                vb[i] = padd(vb[i], wv);
            }
            const f2 zn = MPC_PASS_BOUNDS_REG ? pmin(pmax(vb[i], s.lo[i]), s.hi[i])
                                              : pmin(pmax(vb[i], ldsv(&sm[(39 + i) * LPS + gl])), ldsv(&sm[(44 + i) * LPS + gl]));
            const f2 step = psub(zn, zb[i]);
            zb[i] = zn;
            if (MPC_COMPENSATED_V && i == 4) {
                const f2 zln = mk(zn.x == vb[i].x ? vl4.x : 0.0f, zn.y == vb[i].y ? vl4.y : 0.0f);  // inactive row: z = v
                const f2 stl = psub(zln, zl4);
                zl4 = zln;
                rbd[i] = psub(psub(padd(rbd[i], s1b[i]), step), stl);
                eb[i] = pmul(rb[i], psub(psub(wv, step), stl));  // dy of the bound rows
            } else {
                rbd[i] = psub(padd(rbd[i], s1b[i]), step);
                eb[i] = pmul(rb[i], psub(wv, step));      // dy of the bound rows
            }
        }
        if (first) {  // iteration 1: the dynamics z jumped from the cold start 0 to d (z+ - z = d instead of 0)
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const f2 dd = ldsv(&sm[i * LPS + gl]);
                rdy[i] = psub(rdy[i], dd);
                ed[i] = psub(ed[i], pmul(rd, dd));
            }
        }
        if (MPC_Y_FORM) {
#pragma unroll
            for (int i = 0; i < 3; ++i) yd[i] = padd(yd[i], ed[i]);
        } else {
            At_apply2<LPS, LOOSE>(cm, s, ed, eb, u, u);
        }
        if (keep) {  // the next after_pass() checks: leave the certificates' operands in shared memory
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                sm[(56 + i) * LPS + gl] = dl[i];
                if (!(LOOSE && (i == 1 || i == 2))) sm[(64 + i) * LPS + gl] = eb[i];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) sm[(61 + i) * LPS + gl] = ed[i];
        }
    };
    // termination check / rho adaptation after a pass; returns true when every scenario of the warp is done
    // phase 0 = iterating.  After max_iter passes OSQP (osqp.c, after its main loop) runs a NORMAL termination check if the
    // last pass was not a check pass (phase 1), then the APPROXIMATE one (phase 2: every tolerance x 10, statuses 2 / 3 / 4),
    // else reports max-iter (-2); both go through this same check code.
    int phase = 0;
    float tol = 1.0f;
    auto after_pass = [&]() __attribute__((always_inline)) -> bool {
        bool can_check = true, can_adapt = false;
        if (phase == 0) {
            can_check = (--chk == 0); can_adapt = (--adp == 0);
            if (can_check) chk = st.check_termination;
            if (can_adapt) adp = st.adaptive_rho_interval;
        }
        if (can_check || can_adapt) {
            f2 axd[3], axb[5], zd[3], Di[5], Edi[3], Ebi[5];
#pragma unroll
            for (int i = 0; i < 3; ++i) { zd[i] = ldsv(&sm[i * LPS + gl]); Edi[i] = ldsv(&sm[(21 + i) * LPS + gl]); }
#pragma unroll
            for (int i = 0; i < 5; ++i) { Di[i] = ldsv(&sm[(16 + i) * LPS + gl]); Ebi[i] = ldsv(&sm[(24 + i) * LPS + gl]); }
            A_apply2<LPS, LOOSE>(cm, s, x, axd, axb);
            if (LOOSE) {  // loose rows: z follows A x
                axb[1] = pmul(ldsv(&sm[35 * LPS + gl]), x[1]); axb[2] = pmul(ldsv(&sm[36 * LPS + gl]), x[2]);
                zb[1] = axb[1]; zb[2] = axb[2];
            }
            float pr_s = 0, pr_u = 0, nz_s = 0, nz_u = 0, nax_s = 0, nax_u = 0;
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const f2 r = psub(axd[i], zd[i]);
                amax(pr_s, r); amax(pr_u, pmul(r, Edi[i]));
                amax(nz_s, zd[i]); amax(nz_u, pmul(zd[i], Edi[i]));
                amax(nax_s, axd[i]); amax(nax_u, pmul(axd[i], Edi[i]));
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const f2 r = psub(axb[i], zb[i]);
                amax(pr_s, r); amax(pr_u, pmul(r, Ebi[i]));
                amax(nz_s, zb[i]); amax(nz_u, pmul(zb[i], Ebi[i]));
                amax(nax_s, axb[i]); amax(nax_u, pmul(axb[i], Ebi[i]));
            }
            float du_s = 0, du_u = 0, npx_s = 0, npx_u = 0, naty_s = 0, naty_u = 0;
            const float nq_s = ldsv(&sm[54 * LPS + gl]).x, nq_u = ldsv(&sm[54 * LPS + gl]).y, cs = ldsv(&sm[55 * LPS + gl]).x, cinv = ldsv(&sm[55 * LPS + gl]).y;
            f2 aty5[5];
            if (MPC_Y_FORM) {  // A'y from the duals themselves: y_b = rho (v - z)
                f2 yb5[5];
                const f2 z5[5] = {zero, zero, zero, zero, zero};
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    if (LOOSE && (i == 1 || i == 2)) { yb5[i] = zero; continue; }
                    f2 vz = psub(vb[i], zb[i]);
                    if (MPC_COMPENSATED_V && i == 4) vz = padd(vz, psub(vl4, zl4));
                    yb5[i] = pmul(rb[i], vz);
                }
                At_apply2<LPS, LOOSE>(cm, s, yd, yb5, z5, aty5);
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const f2 px = pmul(ldsv(&sm[(49 + i) * LPS + gl]), x[i]);
                const f2 qi = ldsv(&sm[(29 + i) * LPS + gl]);
                const f2 r = MPC_Y_FORM ? padd(padd(px, qi), aty5[i]) : padd(px, u[i]);
                const f2 aty = MPC_Y_FORM ? aty5[i] : psub(u[i], qi);
                amax(du_s, r); amax(du_u, pmul(r, Di[i]));
                amax(npx_s, px); amax(npx_u, pmul(px, Di[i]));
                amax(naty_s, aty); amax(naty_u, pmul(aty, Di[i]));
            }
            pr_s = cm.max(pr_s); pr_u = cm.max(pr_u); du_s = cm.max(du_s); du_u = cm.max(du_u) * cinv;
            nz_s = cm.max(nz_s); nz_u = cm.max(nz_u); nax_s = cm.max(nax_s); nax_u = cm.max(nax_u);
            npx_s = cm.max(npx_s); npx_u = cm.max(npx_u); naty_s = cm.max(naty_s); naty_u = cm.max(naty_u);
            if (can_check) {
                int status = 0;
                if (pr_u > (float)kOsqpInfty || du_u > (float)kOsqpInfty) status = -7;
                const float eps_prim = tol * ((float)st.eps_abs + (float)st.eps_rel * fmaxf(nz_u, nax_u));
                const float eps_dual = tol * ((float)st.eps_abs + (float)st.eps_rel * cinv * fmaxf(fmaxf(nq_u, naty_u), npx_u));
                const bool prim_ok = pr_u < eps_prim, dual_ok = du_u < eps_dual;
                if (status == 0 && prim_ok && dual_ok) status = phase == 2 ? 2 : 1;
                const bool open = !done && status == 0;  // this scenario still needs the certificates
                bool pinf = false, dinf = false;
                if (GC::warp_any(open && !prim_ok)) {  // is_primal_infeasible
                    const float epi = tol * (float)st.eps_prim_inf;
                    f2 pyb[5], ed[3], eb[5];
#pragma unroll
                    for (int i = 0; i < 3; ++i) ed[i] = ldsv(&sm[(61 + i) * LPS + gl]);
#pragma unroll
                    for (int i = 0; i < 5; ++i) eb[i] = (LOOSE && (i == 1 || i == 2)) ? zero : ldsv(&sm[(64 + i) * LPS + gl]);
                    float ndy = 0, lhs = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        amax(ndy, pmul(ldsv(&sm[(8 + i) * LPS + gl]), ed[i]));
                        const f2 t = pmul(zd[i], ed[i]);  // u*max(dy,0) + l*min(dy,0) with l = u = d
                        lhs += t.x + t.y;
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        if (LOOSE && (i == 1 || i == 2)) { pyb[i] = zero; continue; }
                        float dv[2] = {eb[i].x, eb[i].y};
                        const f2 lo2 = ldsv(&sm[(39 + i) * LPS + gl]), hi2 = ldsv(&sm[(44 + i) * LPS + gl]);
                        const float lov[2] = {lo2.x, lo2.y}, hiv[2] = {hi2.x, hi2.y};
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            float d = dv[h];
                            if (hiv[h] > thr) d = (lov[h] < -thr) ? 0.0f : fminf(d, 0.0f);
                            else if (lov[h] < -thr) d = fmaxf(d, 0.0f);
                            dv[h] = d;
                            lhs += hiv[h] * fmaxf(d, 0.0f) + lov[h] * fminf(d, 0.0f);
                        }
                        pyb[i] = mk(dv[0], dv[1]);
                        amax(ndy, pmul(ldsv(&sm[(11 + i) * LPS + gl]), pyb[i]));
                    }
                    ndy = cm.max(ndy);
                    lhs = cm.sum(lhs);
                    const bool cand = open && !prim_ok && ndy > epi && lhs < -epi * ndy;
                    if (GC::warp_any(cand)) {
                        f2 atdy[5];
                        const f2 z5[5] = {zero, zero, zero, zero, zero};
                        float na = 0;
                        At_apply2<LPS, LOOSE>(cm, s, ed, pyb, z5, atdy);
#pragma unroll
                        for (int i = 0; i < 5; ++i) amax(na, pmul(atdy[i], Di[i]));
                        na = cm.max(na);
                        pinf = cand && na < epi * ndy;
                    }
                }
                if (GC::warp_any(open && !dual_ok && !pinf)) {  // is_dual_infeasible (dx = alpha D of this iteration)
                    const float edi = tol * (float)st.eps_dual_inf;
                    f2 dl[5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) dl[i] = ldsv(&sm[(56 + i) * LPS + gl]);
                    float ndx = 0, qdx = 0, npdx = 0;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        amax(ndx, pmul(ldsv(&sm[(3 + i) * LPS + gl]), dl[i]));
                        const f2 t = pmul(ldsv(&sm[(29 + i) * LPS + gl]), dl[i]);
                        qdx += t.x + t.y;
                        amax(npdx, pmul(pmul(ldsv(&sm[(49 + i) * LPS + gl]), dl[i]), Di[i]));
                    }
                    ndx = cm.max(ndx);
                    qdx = cm.sum(qdx);
                    npdx = cm.max(npdx);
                    const bool cand = open && !dual_ok && !pinf && ndx > edi && qdx < -cs * edi * ndx && npdx < cs * edi * ndx;
                    if (GC::warp_any(cand)) {
                        f2 adxd[3], adxb[5];
                        A_apply2<LPS, LOOSE>(cm, s, dl, adxd, adxb);
                        if (LOOSE) { adxb[1] = pmul(ldsv(&sm[35 * LPS + gl]), dl[1]); adxb[2] = pmul(ldsv(&sm[36 * LPS + gl]), dl[2]); }
                        int bad = 0;
                        const float lim = edi * ndx;
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const f2 v = pmul(adxd[i], Edi[i]);  // equality rows have finite bounds
                            if (fabsf(v.x) > lim || fabsf(v.y) > lim) bad = 1;
                        }
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const f2 v = pmul(adxb[i], Ebi[i]);
                            const f2 lo2 = ldsv(&sm[(39 + i) * LPS + gl]), hi2 = ldsv(&sm[(44 + i) * LPS + gl]);
                            if ((hi2.x < thr && v.x > lim) || (lo2.x > -thr && v.x < -lim)) bad = 1;
                            if ((hi2.y < thr && v.y > lim) || (lo2.y > -thr && v.y < -lim)) bad = 1;
                        }
                        dinf = cand && !cm.any(bad != 0);
                    }
                }
                if (status == 0 && pinf) status = phase == 2 ? 3 : -3;
                if (status == 0 && dinf) status = phase == 2 ? 4 : -4;
                if (status == 0 && phase == 2) status = -2;
                if (!done && status != 0) finish(status, phase ? st.max_iter : iter);
                if (GC::warp_all(done)) return true;
            }
            if (can_adapt) {  // adapt_rho / compute_rho_estimate on the scaled residuals
                const float pn = pr_s / (fmaxf(nz_s, nax_s) + 1e-10f);
                const float dn = du_s / (fmaxf(fmaxf(nq_s, naty_s), npx_s) + 1e-10f);
                float rnew = rho * sqrtf(pn / (dn + 1e-10f));
                rnew = fminf(fmaxf(rnew, (float)kRhoMin), (float)kRhoMax);
                const bool upd = !done && (rnew > rho * (float)st.adaptive_rho_tolerance ||
                                           rnew < rho / (float)st.adaptive_rho_tolerance);
                if (GC::warp_any(upd)) {  // scenarios that keep their rho recompute an identical factor
                    if (upd) {
                        const float ratio = rho / rnew;  // y is unchanged: v = z + (v - z) rho_old / rho_new
                        rho = rnew;
                        const float thr2 = thr;
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (LOOSE && (i == 1 || i == 2)) continue;
                            // rows with rho fixed at rho_min (loose) keep their v
                            const f2 lo2 = ldsv(&sm[(39 + i) * LPS + gl]), hi2 = ldsv(&sm[(44 + i) * LPS + gl]);
                            const f2 rr = mk((lo2.x < -thr2 && hi2.x > thr2) ? 1.0f : ratio,
                                             (lo2.y < -thr2 && hi2.y > thr2) ? 1.0f : ratio);
                            if (MPC_COMPENSATED_V && i == 4) {
                                vb[i] = pfma(padd(psub(vb[i], zb[i]), psub(vl4, zl4)), rr, zb[i]);
                                vl4 = zl4;
                            } else {
                                vb[i] = pfma(psub(vb[i], zb[i]), rr, zb[i]);
                            }
                        }
                    }
                    set_rho2<LPS, LOOSE>(sm, gl, rho, rdf, rb);
                    rd = bc(rdf);
                    factorize2<LPS, LOOSE>(cm, s, f, sigma, rdf, rb, sm, cf);
                }
            }
        }
        return false;
    };
    for (iter = 1;; ++iter) {
        // Passes after which nothing happens (no check, no rho adaptation, not the last one) run in a loop of their own:
        // one straight-line body and one backward branch, instead of a round trip through the check code's branches
        // (which ptxas lays out far from the pass and which showed up as instruction-fetch stalls).  Same passes, same
        // order; `after_pass` would only have decremented the two counters.
        if (phase == 0 && iter > 1) {
            int quiet = st.max_iter - iter;
            if (chk > 0) quiet = min(quiet, chk - 1);
            if (adp > 0) quiet = min(quiet, adp - 1);
#pragma unroll kQuietUnroll
            for (int i = 0; i < quiet; ++i) pass(false, false);
            if (quiet > 0) {
                iter += quiet;
                if (chk > 0) chk -= quiet;
                if (adp > 0) adp -= quiet;
            }
        }
#ifdef MPC_QUAD_MARK   // PMTRIG markers around the pass (tools/sass_pass.py)
        asm volatile("pmevent 1;");
#endif
        if (phase == 0) pass(iter == 1, chk == 1 || iter >= st.max_iter);
#ifdef MPC_QUAD_MARK
        asm volatile("pmevent 2;");
#endif
        if (after_pass()) break;   // phase 2 always ends here: every scenario still open is finished with -2
#ifdef MPC_PHASE_CLOCK
        if (iter == 1) MPC_PHASE_MARK(7);
#endif
        if (phase == 1 || (phase == 0 && iter >= st.max_iter)) {
            // the last pass was a check pass iff check_termination divides max_iter
            const bool checked = phase == 0 && st.check_termination > 0 && (st.max_iter % st.check_termination == 0);
            phase = (phase == 1 || checked) ? 2 : 1;
            if (phase == 2) tol = 10.0f;
        }
    }
}

}  // namespace mpcb
