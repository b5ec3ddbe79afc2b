// admm_quad.cuh -- K1 + K2, four stages per lane: the one-wave fp32 variant for sm_100a (N + 1 <= 4 * LPS stages).
//
// Why.  The paired kernel (admm_pair.cuh) is a latency-bound dependent chain at two warps per scheduler: its register file
// is full (255 registers for two stages per lane), so 4096 cars are 2048 warps on 1184 resident slots -- two rounds.  Here a
// scenario occupies LPS = 8 lanes and a lane owns FOUR consecutive stages, s0 .. s3 = 4l .. 4l + 3, packed as two float2
// slices  E = (s0, s2)  and  O = (s1, s3):
//   * four scenarios share a warp, 4096 cars are 1024 warps: ONE round on 148 x 7 slots;
//   * every element-wise operation exists twice (slice E, slice O), independent of each other: ILP 2 on the whole pass;
//   * the neighbour of a stage is mostly in the same lane: the stage BEFORE an O stage is the E stage with the same half,
//     the stage AFTER an E stage is the O stage with the same half -- no data movement at all; the other two directions
//     cost one shuffle per value (prev(O.y) / next(E.x)), i.e. one shuffle per quantity per FOUR stages;
//   * the block-tridiagonal solve eliminates the E stages inside the lane as PACKED 3x3 algebra (cyclic reduction level 1,
//     both E stages at once), then s1 (level 2, scalar), and runs parallel cyclic reduction over the s3 chain of LPS
//     blocks: log2(8) = 3 levels instead of 4, 31 shuffles per pass for four scenarios instead of 31 for two.
// Same OSQP iteration, same v-form / u-form bookkeeping, same termination logic as admm_pair.cuh (which documents them);
// e_psi / t rows are the reference's unbounded ("loose") rows -- other configurations use the paired kernel.
#pragma once
#include "admm_pair.cuh"

namespace mpcb {

template <int LPS> struct QuadComm : GroupComm<LPS> {
    // the value the PREVIOUS / NEXT stage holds, for the two stages of slice sl (0 = E, 1 = O), given both slices' values
    __device__ __forceinline__ f2 prev_of(int sl, f2 vE, f2 vO) const { return sl == 0 ? mk(this->prev(vO.y), vO.x) : vE; }
    __device__ __forceinline__ f2 next_of(int sl, f2 vE, f2 vO) const { return sl == 0 ? vO : mk(vE.y, this->next(vE.x)); }
};

// packed 3x3 helpers (row-major, both halves at once)
__device__ __forceinline__ void mm3p(const f2* A, const f2* B, f2* C) {
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[3 * i + j] = pfma(A[3 * i + 2], B[6 + j], pfma(A[3 * i + 1], B[3 + j], pmul(A[3 * i], B[j])));
}
__device__ __forceinline__ f2 prcp(f2 v) { return mk(1.0f / v.x, 1.0f / v.y); }

// per-scenario shared constants: the LOGICAL rows of admm_pair.cuh (kPairRows), one set per slice: [row][slice][LPS] float2.
// Not stored: rows 16 .. 28 (reciprocals of the scalings: recomputed at the checks) and rows 34, 37, 38 (e of the bound rows
// that live in registers) -- 40 physical rows, 20 KB per warp, so that eight warps fit an SM's shared memory.
constexpr int kQuadRows = 40;
__host__ __device__ constexpr int quad_phys_row(int row) {
    return row < 16 ? row : (row < 29 ? -1 : (row < 34 ? row - 13 : (row == 35 ? 21 : (row == 36 ? 22 : (row >= 39 ? row - 16 : -1)))));
}
template <int LPS> __device__ __forceinline__ int qrow(int row, int sl, int gl) { return (quad_phys_row(row) * LPS + gl) * 2 + sl; }
// both slices of one row with a single LDS.128: (x, y) = slice E, (z, w) = slice O.  A quarter-warp (one scenario) reads 128
// contiguous bytes, the four scenarios of the warp four different lines: conflict-free.
template <int LPS> __device__ __forceinline__ float4 qrow2(const f2* sm, int row, int gl) {
    return lds128v(reinterpret_cast<const float4*>(sm + (quad_phys_row(row) * LPS + gl) * 2));
}
__device__ __forceinline__ f2 slice_of(const float4& v, int sl) { return sl == 0 ? mk(v.x, v.y) : mk(v.z, v.w); }
// rho of the bound rows 0 / 3 / 4 of the four stages, as 2-bit codes (0 inequality: rho, 1 equality-typed: 1e3 rho, 2 loose:
// rho_min) in one register instead of six float2: bit pair 2 * (6 sl + 3 half + j), j = 0, 1, 2 for rows 0, 3, 4
__device__ __forceinline__ float rho_code(unsigned codes, int sl, int half, int j, float rho, float rdf) {
    const unsigned c = (codes >> (2 * (6 * sl + 3 * half + j))) & 3u;
    return c == 0 ? rho : (c == 1 ? rdf : (float)kRhoMin);
}
__device__ __forceinline__ f2 rho_row(unsigned codes, int sl, int i, float rho, float rdf) {
    const int j = i == 0 ? 0 : i - 2;
    return mk(rho_code(codes, sl, 0, j, rho, rdf), rho_code(codes, sl, 1, j, rho, rdf));
}

template <int LPS> struct QuadFactor {
    static constexpr int NLEV = PairFactor<LPS>::NLEV;
    f2 iv[2], ik[2], nsxv0[2], nsxv2[2], nsxk0[2], nsxk1[2], nfv[2], nfk[2];  // input elimination per slice (negated)
    f2 DEi[6], UE[6], UO[6];      // level 1: DE^-1 (00 01 02 11 12 22), nonzeros (0 1 2 3 4 8) of U of the E and the O stages
    float DAi[6], UA[9], UB[9];   // level 2: A = s1 eliminated; UA couples s1 -> s3, UB couples s3 -> s1 of the next lane
    float last[9], Dinv[6];       // PCR over the s3 chain (levels 0 .. NLEV-2 in shared memory, as in admm_pair.cuh)
};

__device__ __forceinline__ void inv3sym6p(const f2* M /*00 01 02 11 12 22*/, f2* R) {
    const f2 a = M[0], b = M[1], c = M[2], d = M[3], e = M[4], f = M[5];
    const f2 A = psub(pmul(d, f), pmul(e, e)), B = psub(pmul(c, e), pmul(b, f)), C = psub(pmul(b, e), pmul(c, d));
    const f2 r = prcp(pfma(c, C, pfma(b, B, pmul(a, A))));
    R[0] = pmul(A, r); R[1] = pmul(B, r); R[2] = pmul(C, r);
    R[3] = pmul(psub(pmul(a, f), pmul(c, c)), r); R[4] = pmul(psub(pmul(b, c), pmul(a, e)), r);
    R[5] = pmul(psub(pmul(a, d), pmul(b, b)), r);
}

// OSQP scale_data, four stages per lane
template <int LPS>
__device__ __forceinline__ void ruiz_scale4(const QuadComm<LPS>& cm, Stage2 (&s)[2], int iters, int nvar) {
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 5; ++i) { s[sl].D[i] = bc(1.0f); s[sl].Eb[i] = bc(1.0f); }
#pragma unroll
        for (int i = 0; i < 3; ++i) s[sl].Ed[i] = bc(1.0f);
    }
    float cs = 1.0f;
    const float inv_nvar = 1.0f / (float)nvar;
    for (int it = 0; it < iters; ++it) {
        f2 Dt[2][5], Edt[2][3], Ebt[2][5], En[2][3], ro[2][3];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            f2 aa[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) aa[i] = pabs(s[sl].a[i]);
            f2 col[5];
            col[0] = pmax(pmax(aa[0], aa[2]), aa[4]);
            col[1] = pmax(aa[1], aa[3]);
            col[2] = aa[5];
            col[3] = aa[7];
            col[4] = aa[6];
#pragma unroll
            for (int i = 0; i < 3; ++i) col[i] = pmax(col[i], pabs(s[sl].c[i]));
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                col[i] = pmax(pmax(col[i], pabs(s[sl].e[i])), pabs(s[sl].P[i]));
                Dt[sl][i] = prsqrt_lim(col[i]);
                Ebt[sl][i] = prsqrt_lim(pabs(s[sl].e[i]));
            }
            ro[sl][0] = pmax(aa[0], aa[1]);
            ro[sl][1] = pmax(pmax(aa[2], aa[3]), aa[6]);
            ro[sl][2] = pmax(pmax(aa[4], aa[5]), aa[7]);
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 3; ++i)
                Edt[sl][i] = prsqrt_lim(pmax(pabs(s[sl].c[i]), cm.prev_of(sl, ro[0][i], ro[1][i])));
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 3; ++i) En[sl][i] = cm.next_of(sl, Edt[0][i], Edt[1][i]);
        float sp = 0.0f, mq = 0.0f;
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            Stage2& t = s[sl];
            const f2* D = Dt[sl];
            const f2* E = En[sl];
#pragma unroll
            for (int i = 0; i < 5; ++i) t.P[i] = pmul(pmul(t.P[i], D[i]), D[i]);
            t.a[0] = pmul(pmul(t.a[0], E[0]), D[0]); t.a[1] = pmul(pmul(t.a[1], E[0]), D[1]);
            t.a[2] = pmul(pmul(t.a[2], E[1]), D[0]); t.a[3] = pmul(pmul(t.a[3], E[1]), D[1]);
            t.a[4] = pmul(pmul(t.a[4], E[2]), D[0]); t.a[5] = pmul(pmul(t.a[5], E[2]), D[2]);
            t.a[6] = pmul(pmul(t.a[6], E[1]), D[4]); t.a[7] = pmul(pmul(t.a[7], E[2]), D[3]);
#pragma unroll
            for (int i = 0; i < 3; ++i) { t.c[i] = pmul(pmul(t.c[i], Edt[sl][i]), D[i]); t.Ed[i] = pmul(t.Ed[i], Edt[sl][i]); }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                t.e[i] = pmul(pmul(t.e[i], Ebt[sl][i]), D[i]);
                t.q[i] = pmul(t.q[i], D[i]);
                t.D[i] = pmul(t.D[i], D[i]);
                t.Eb[i] = pmul(t.Eb[i], Ebt[sl][i]);
                sp += fabsf(t.P[i].x) + fabsf(t.P[i].y);
                amax(mq, t.q[i]);
            }
        }
        sp = cm.sum(sp) * inv_nvar;
        mq = limit_scaling_f(cm.max(mq));
        const float ct = 1.0f / limit_scaling_f(fmaxf(sp, mq));
        const f2 ct2 = bc(ct);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 5; ++i) { s[sl].P[i] = pmul(s[sl].P[i], ct2); s[sl].q[i] = pmul(s[sl].q[i], ct2); }
        cs *= ct;
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        s[sl].cs = cs;
#pragma unroll
        for (int i = 0; i < 3; ++i) s[sl].d[i] = pmul(s[sl].d[i], s[sl].Ed[i]);
#pragma unroll
        for (int i = 0; i < 5; ++i) { s[sl].lo[i] = pmul(s[sl].lo[i], s[sl].Eb[i]); s[sl].hi[i] = pmul(s[sl].hi[i], s[sl].Eb[i]); }
    }
}

// z = A w: zd (dynamics rows of each stage) and zb (bound rows 0, 3, 4; rows 1, 2 are loose)
template <int LPS>
__device__ __forceinline__ void A_apply4(const QuadComm<LPS>& cm, const Stage2 (&s)[2], const f2 (&w)[2][5], f2 (&zd)[2][3],
                                         f2 (&zb)[2][5]) {
    f2 o[2][3];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2* a = s[sl].a;
        o[sl][0] = pfma(a[1], w[sl][1], pmul(a[0], w[sl][0]));
        o[sl][1] = pfma(a[6], w[sl][4], pfma(a[3], w[sl][1], pmul(a[2], w[sl][0])));
        o[sl][2] = pfma(a[7], w[sl][3], pfma(a[5], w[sl][2], pmul(a[4], w[sl][0])));
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 3; ++i) zd[sl][i] = pfma(s[sl].c[i], w[sl][i], cm.prev_of(sl, o[0][i], o[1][i]));
        zb[sl][0] = pmul(s[sl].e[0], w[sl][0]);
        zb[sl][3] = pmul(s[sl].e[3], w[sl][3]);
        zb[sl][4] = pmul(s[sl].e[4], w[sl][4]);
    }
}

// r = acc + A' y (yb[1], yb[2] are identically zero and not read)
template <int LPS>
__device__ __forceinline__ void At_apply4(const QuadComm<LPS>& cm, const Stage2 (&s)[2], const f2 (&yd)[2][3],
                                          const f2 (&yb)[2][5], const f2 (&acc)[2][5], f2 (&r)[2][5]) {
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2* a = s[sl].a;
        const f2 g0 = cm.next_of(sl, yd[0][0], yd[1][0]), g1 = cm.next_of(sl, yd[0][1], yd[1][1]),
                 g2 = cm.next_of(sl, yd[0][2], yd[1][2]);
        const f2 r0 = pfma(a[4], g2, pfma(a[2], g1, pfma(a[0], g0, pfma(s[sl].c[0], yd[sl][0], pfma(s[sl].e[0], yb[sl][0], acc[sl][0])))));
        const f2 r1 = pfma(a[3], g1, pfma(a[1], g0, pfma(s[sl].c[1], yd[sl][1], acc[sl][1])));
        const f2 r2 = pfma(a[5], g2, pfma(s[sl].c[2], yd[sl][2], acc[sl][2]));
        const f2 r3 = pfma(a[7], g2, pfma(s[sl].e[3], yb[sl][3], acc[sl][3]));
        const f2 r4 = pfma(a[6], g1, pfma(s[sl].e[4], yb[sl][4], acc[sl][4]));
        r[sl][0] = r0; r[sl][1] = r1; r[sl][2] = r2; r[sl][3] = r3; r[sl][4] = r4;
    }
}

// S = P + sigma I + A' R A for the four stages of the lane; inputs eliminated; E stages eliminated (level 1, packed);
// s1 eliminated (level 2); PCR factorisation of the s3 chain.
template <int LPS>
__device__ __forceinline__ void factorize4(const QuadComm<LPS>& cm, const Stage2 (&s)[2], QuadFactor<LPS>& f, float sigma,
                                           float rho, float rdf, unsigned codes, const f2* sm, float4* cf) {
    constexpr int NLEV = QuadFactor<LPS>::NLEV;
    const f2 rd = bc(rdf), sg = bc(sigma);
    f2 D00[2], D01[2], D02[2], D11[2], D22[2], U0[2], U1[2], U2[2], U3[2], U4[2], U8[2], tkk[2], tvv[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2* a = s[sl].a;
        f2 diag[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const bool lz = (i == 1 || i == 2);  // loose rows: rho = rho_min, e kept in shared memory
            const f2 ei = lz ? sm[qrow<LPS>(34 + i, sl, cm.gl)] : s[sl].e[i];
            const f2 ri = lz ? bc((float)kRhoMin) : rho_row(codes, sl, i, rho, rdf);
            diag[i] = pfma(pmul(ri, ei), ei, padd(sm[qrow<LPS>(49 + i, sl, cm.gl)], sg));
        }
        f2 cn[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cn[i] = cm.next_of(sl, s[0].c[i], s[1].c[i]);
        const f2* c = s[sl].c;
        D00[sl] = pfma(rd, pfma(a[4], a[4], pfma(a[2], a[2], pfma(a[0], a[0], pmul(c[0], c[0])))), diag[0]);
        D11[sl] = pfma(rd, pfma(a[3], a[3], pfma(a[1], a[1], pmul(c[1], c[1]))), diag[1]);
        D22[sl] = pfma(rd, pfma(a[5], a[5], pmul(c[2], c[2])), diag[2]);
        D01[sl] = pmul(rd, pfma(a[2], a[3], pmul(a[0], a[1])));
        D02[sl] = pmul(rd, pmul(a[4], a[5]));
        const f2 Svv = pfma(rd, pmul(a[7], a[7]), diag[3]);
        const f2 Skk = pfma(rd, pmul(a[6], a[6]), diag[4]);
        f.iv[sl] = prcp(Svv);
        f.ik[sl] = prcp(Skk);
        const f2 ra7 = pmul(rd, a[7]), ra6 = pmul(rd, a[6]);
        const f2 sxv0 = pmul(ra7, a[4]), sxv2 = pmul(ra7, a[5]);
        const f2 sxk0 = pmul(ra6, a[2]), sxk1 = pmul(ra6, a[3]);
        const f2 fv = pmul(ra7, cn[2]), fk = pmul(ra6, cn[1]);
        const f2 rc0 = pmul(rd, cn[0]), rc1 = pmul(rd, cn[1]), rc2 = pmul(rd, cn[2]);
        U0[sl] = pmul(a[0], rc0); U1[sl] = pmul(a[2], rc1); U2[sl] = pmul(a[4], rc2);
        U3[sl] = pmul(a[1], rc0); U4[sl] = pmul(a[3], rc1); U8[sl] = pmul(a[5], rc2);
        const f2 ivs0 = pmul(f.iv[sl], sxv0), ivs2 = pmul(f.iv[sl], sxv2), iks0 = pmul(f.ik[sl], sxk0), iks1 = pmul(f.ik[sl], sxk1);
        D00[sl] = psub(D00[sl], pfma(iks0, sxk0, pmul(ivs0, sxv0)));
        D01[sl] = psub(D01[sl], pmul(iks0, sxk1));
        D02[sl] = psub(D02[sl], pmul(ivs0, sxv2));
        D11[sl] = psub(D11[sl], pmul(iks1, sxk1));
        D22[sl] = psub(D22[sl], pmul(ivs2, sxv2));
        U2[sl] = psub(U2[sl], pmul(ivs0, fv)); U8[sl] = psub(U8[sl], pmul(ivs2, fv));
        U1[sl] = psub(U1[sl], pmul(iks0, fk)); U4[sl] = psub(U4[sl], pmul(iks1, fk));
        tkk[sl] = pmul(pmul(f.ik[sl], fk), fk);
        tvv[sl] = pmul(pmul(f.iv[sl], fv), fv);
        const f2 m1 = bc(-1.0f);
        f.nsxv0[sl] = pmul(sxv0, m1); f.nsxv2[sl] = pmul(sxv2, m1); f.nsxk0[sl] = pmul(sxk0, m1); f.nsxk1[sl] = pmul(sxk1, m1);
        f.nfv[sl] = pmul(fv, m1); f.nfk[sl] = pmul(fk, m1);
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {  // what the previous stage's input elimination leaves on this stage's diagonal
        D11[sl] = psub(D11[sl], cm.prev_of(sl, tkk[0], tkk[1]));
        D22[sl] = psub(D22[sl], cm.prev_of(sl, tvv[0], tvv[1]));
    }
    // ---- level 1: eliminate the E stages (packed) ----
    {
        const f2 DE[6] = {D00[0], D01[0], D02[0], D11[0], bc(0.0f), D22[0]};
        inv3sym6p(DE, f.DEi);
    }
    f.UE[0] = U0[0]; f.UE[1] = U1[0]; f.UE[2] = U2[0]; f.UE[3] = U3[0]; f.UE[4] = U4[0]; f.UE[5] = U8[0];
    f.UO[0] = U0[1]; f.UO[1] = U1[1]; f.UO[2] = U2[1]; f.UO[3] = U3[1]; f.UO[4] = U4[1]; f.UO[5] = U8[1];
    const f2 z2 = bc(0.0f);
    const f2 Di9[9] = {f.DEi[0], f.DEi[1], f.DEi[2], f.DEi[1], f.DEi[3], f.DEi[4], f.DEi[2], f.DEi[4], f.DEi[5]};
    const f2 UE9[9] = {U0[0], U1[0], U2[0], U3[0], U4[0], z2, z2, z2, U8[0]};
    const f2 UO9[9] = {U0[1], U1[1], U2[1], U3[1], U4[1], z2, z2, z2, U8[1]};
    f2 LoE[9];  // (U of the odd stage before each E stage)'
    {
        f2 p[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) p[i] = (i == 5 || i == 6 || i == 7) ? z2 : cm.prev_of(0, z2, UO9[i]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) LoE[3 * i + k] = p[3 * k + i];
    }
    f2 G[9], H[9], Gn[9], Hn[9];
    mm3p(Di9, UE9, G);
    mm3p(Di9, LoE, H);
#pragma unroll
    for (int i = 0; i < 9; ++i) { Gn[i] = cm.next_of(1, G[i], z2); Hn[i] = cm.next_of(1, H[i], z2); }
    f2 DO9[9], UOn[9];  // the odd chain after level 1: diagonal blocks and couplings s1 -> s3 (.x), s3 -> next s1 (.y)
    {
        const f2 DB[9] = {D00[1], D01[1], D02[1], D01[1], D11[1], z2, D02[1], z2, D22[1]};
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                f2 acc = DB[3 * i + k];
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    acc = psub(acc, pfma(UO9[3 * i + j], Hn[3 * j + k], pmul(UE9[3 * j + i], G[3 * j + k])));
                DO9[3 * i + k] = acc;
            }
        f2 t[9];
        mm3p(UO9, Gn, t);
#pragma unroll
        for (int i = 0; i < 9; ++i) UOn[i] = pmul(t[i], bc(-1.0f));
    }
    // ---- level 2: eliminate s1 (= .x of the odd chain) inside the lane ----
    float DA9[9], DB9[9], UA[9], UB[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { DA9[i] = DO9[i].x; DB9[i] = DO9[i].y; UA[i] = UOn[i].x; UB[i] = UOn[i].y; f.UA[i] = UA[i]; f.UB[i] = UB[i]; }
    float DAi9[9];
    inv3sym(DA9, DAi9);
    f.DAi[0] = DAi9[0]; f.DAi[1] = DAi9[1]; f.DAi[2] = DAi9[2]; f.DAi[3] = DAi9[4]; f.DAi[4] = DAi9[5]; f.DAi[5] = DAi9[8];
    float LoA[9];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) LoA[3 * i + k] = cm.prev(UB[3 * k + i]);
    float G2[9], H2[9], G2n[9], H2n[9];
    mm3(DAi9, UA, G2);
    mm3(DAi9, LoA, H2);
#pragma unroll
    for (int i = 0; i < 9; ++i) { G2n[i] = cm.next(G2[i]); H2n[i] = cm.next(H2[i]); }
    float Dm[9], U[9], Lo[9];
    {
#pragma unroll
        for (int i = 0; i < 3; ++i)
#pragma unroll
            for (int k = 0; k < 3; ++k) {
                float acc = DB9[3 * i + k];
#pragma unroll
                for (int j = 0; j < 3; ++j) acc -= UA[3 * j + i] * G2[3 * j + k] + UB[3 * i + j] * H2n[3 * j + k];
                Dm[3 * i + k] = acc;
            }
        float t[9];
        mm3(UB, G2n, t);
#pragma unroll
        for (int i = 0; i < 9; ++i) U[i] = -t[i];
    }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int k = 0; k < 3; ++k) Lo[3 * i + k] = cm.prev(U[3 * k + i]);
    // ---- PCR over the s3 chain (as admm_pair.cuh::factorize2) ----
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
            if (lev < NLEV - 1) {
                const int j = 9 * lev + i;  // float2 index: float4 column j / 2 of this lane, half j & 1
                reinterpret_cast<f2*>(cf + (j >> 1) * 32)[j & 1] = mk(-al[i], -be[i]);
            } else {
                f.last[i] = has_up ? al[i] : be[i];
            }
        }
    }
    float Di[9];
    inv3sym(Dm, Di);
    f.Dinv[0] = Di[0]; f.Dinv[1] = Di[1]; f.Dinv[2] = Di[2]; f.Dinv[3] = Di[4]; f.Dinv[4] = Di[5]; f.Dinv[5] = Di[8];
}

// x = S^-1 b for the four stages of the lane
template <int LPS>
__device__ __forceinline__ void kkt_solve4(const QuadComm<LPS>& cm, const QuadFactor<LPS>& f, const f2 (&b)[2][5], f2 (&x)[2][5],
                                           const float4* cf) {
    constexpr int NLEV = QuadFactor<LPS>::NLEV;
    float4 cq[PcrCoef<LPS>::kF4 > 0 ? PcrCoef<LPS>::kF4 : 1];
#pragma unroll
    for (int k = 0; k < PcrCoef<LPS>::kF4; ++k) cq[k] = lds128v(cf + k * 32);
    // input elimination
    f2 bx[2][3], tk[2], tv[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2 bv = pmul(f.iv[sl], b[sl][3]), bk = pmul(f.ik[sl], b[sl][4]);
        bx[sl][0] = pfma(bk, f.nsxk0[sl], pfma(bv, f.nsxv0[sl], b[sl][0]));
        bx[sl][1] = pfma(bk, f.nsxk1[sl], b[sl][1]);
        bx[sl][2] = pfma(bv, f.nsxv2[sl], b[sl][2]);
        tk[sl] = pmul(bk, f.nfk[sl]);
        tv[sl] = pmul(bv, f.nfv[sl]);
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        bx[sl][1] = padd(bx[sl][1], cm.prev_of(sl, tk[0], tk[1]));
        bx[sl][2] = padd(bx[sl][2], cm.prev_of(sl, tv[0], tv[1]));
    }
    // level 1 forward: tE = DE^-1 b_E;  b_O' = b_O - U_E' tE - U_O tE(next even stage)
    const f2* DEi = f.DEi;
    f2 tE[3];
    tE[0] = pfma(DEi[2], bx[0][2], pfma(DEi[1], bx[0][1], pmul(DEi[0], bx[0][0])));
    tE[1] = pfma(DEi[4], bx[0][2], pfma(DEi[3], bx[0][1], pmul(DEi[1], bx[0][0])));
    tE[2] = pfma(DEi[5], bx[0][2], pfma(DEi[4], bx[0][1], pmul(DEi[2], bx[0][0])));
    const f2 z2 = bc(0.0f);
    f2 tn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) tn[i] = cm.next_of(1, tE[i], z2);
    // U = [u0 u1 u2; u3 u4 0; 0 0 u5]:  (U' t)_k = sum_i U[i][k] t_i,  (U t)_i = sum_k U[i][k] t_k
    const f2 m1 = bc(-1.0f);
    f2 rO[3];
    rO[0] = psub(psub(bx[1][0], pfma(f.UE[3], tE[1], pmul(f.UE[0], tE[0]))), pfma(f.UO[2], tn[2], pfma(f.UO[1], tn[1], pmul(f.UO[0], tn[0]))));
    rO[1] = psub(psub(bx[1][1], pfma(f.UE[4], tE[1], pmul(f.UE[1], tE[0]))), pfma(f.UO[4], tn[1], pmul(f.UO[3], tn[0])));
    rO[2] = psub(psub(bx[1][2], pfma(f.UE[5], tE[2], pmul(f.UE[2], tE[0]))), pmul(f.UO[5], tn[2]));
    (void)m1;
    // level 2 forward (scalar): tA = DA^-1 r_A;  r_B' = r_B - UA' tA - UB tA(next lane)
    const float rA0 = rO[0].x, rA1 = rO[1].x, rA2 = rO[2].x;
    const float tA0 = fmaf(f.DAi[2], rA2, fmaf(f.DAi[1], rA1, f.DAi[0] * rA0));
    const float tA1 = fmaf(f.DAi[4], rA2, fmaf(f.DAi[3], rA1, f.DAi[1] * rA0));
    const float tA2 = fmaf(f.DAi[5], rA2, fmaf(f.DAi[4], rA1, f.DAi[2] * rA0));
    const float n0 = cm.next(tA0), n1 = cm.next(tA1), n2 = cm.next(tA2);
    f2 R0 = mk(rO[0].y, 0.0f), R1 = mk(rO[1].y, 0.0f), R2 = mk(rO[2].y, 0.0f);
    R0.x -= fmaf(f.UA[6], tA2, fmaf(f.UA[3], tA1, f.UA[0] * tA0)) + fmaf(f.UB[2], n2, fmaf(f.UB[1], n1, f.UB[0] * n0));
    R1.x -= fmaf(f.UA[7], tA2, fmaf(f.UA[4], tA1, f.UA[1] * tA0)) + fmaf(f.UB[5], n2, fmaf(f.UB[4], n1, f.UB[3] * n0));
    R2.x -= fmaf(f.UA[8], tA2, fmaf(f.UA[5], tA1, f.UA[2] * tA0)) + fmaf(f.UB[8], n2, fmaf(f.UB[7], n1, f.UB[6] * n0));
    // PCR over the s3 chain
#pragma unroll
    for (int lev = 0; lev < NLEV - 1; ++lev) {
        const int sft = 1 << lev;
        const f2 m0 = mk(cm.up(R0.x, sft), cm.dn(R0.x, sft));
        const f2 m1_ = mk(cm.up(R1.x, sft), cm.dn(R1.x, sft));
        const f2 m2 = mk(cm.up(R2.x, sft), cm.dn(R2.x, sft));
        f2 nab[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = 9 * lev + i;
            nab[i] = (j & 1) ? mk(cq[j >> 1].z, cq[j >> 1].w) : mk(cq[j >> 1].x, cq[j >> 1].y);
        }
        const f2 s0 = pfma(nab[2], m2, pfma(nab[1], m1_, pfma(nab[0], m0, R0)));
        const f2 s1 = pfma(nab[5], m2, pfma(nab[4], m1_, pfma(nab[3], m0, R1)));
        const f2 s2 = pfma(nab[8], m2, pfma(nab[7], m1_, pfma(nab[6], m0, R2)));
        R0.x = s0.x + s0.y; R1.x = s1.x + s1.y; R2.x = s2.x + s2.y;
    }
    float r0 = R0.x, r1 = R1.x, r2 = R2.x;
    {
        const int sft = LPS / 2;
        const float q0 = cm.bfly(r0, sft), q1 = cm.bfly(r1, sft), q2 = cm.bfly(r2, sft);
        r0 = fmaf(-f.last[2], q2, fmaf(-f.last[1], q1, fmaf(-f.last[0], q0, r0)));
        r1 = fmaf(-f.last[5], q2, fmaf(-f.last[4], q1, fmaf(-f.last[3], q0, r1)));
        r2 = fmaf(-f.last[8], q2, fmaf(-f.last[7], q1, fmaf(-f.last[6], q0, r2)));
    }
    const float xB0 = fmaf(f.Dinv[2], r2, fmaf(f.Dinv[1], r1, f.Dinv[0] * r0));
    const float xB1 = fmaf(f.Dinv[4], r2, fmaf(f.Dinv[3], r1, f.Dinv[1] * r0));
    const float xB2 = fmaf(f.Dinv[5], r2, fmaf(f.Dinv[4], r1, f.Dinv[2] * r0));
    // level 2 back: x_A = tA - DA^-1 (UA x_B + [UB' x_B](previous lane))
    const float mB0 = fmaf(f.UB[6], xB2, fmaf(f.UB[3], xB1, f.UB[0] * xB0));
    const float mB1 = fmaf(f.UB[7], xB2, fmaf(f.UB[4], xB1, f.UB[1] * xB0));
    const float mB2 = fmaf(f.UB[8], xB2, fmaf(f.UB[5], xB1, f.UB[2] * xB0));
    const float w0 = fmaf(f.UA[2], xB2, fmaf(f.UA[1], xB1, fmaf(f.UA[0], xB0, cm.prev(mB0))));
    const float w1 = fmaf(f.UA[5], xB2, fmaf(f.UA[4], xB1, fmaf(f.UA[3], xB0, cm.prev(mB1))));
    const float w2 = fmaf(f.UA[8], xB2, fmaf(f.UA[7], xB1, fmaf(f.UA[6], xB0, cm.prev(mB2))));
    const float xA0 = fmaf(-f.DAi[2], w2, fmaf(-f.DAi[1], w1, fmaf(-f.DAi[0], w0, tA0)));
    const float xA1 = fmaf(-f.DAi[4], w2, fmaf(-f.DAi[3], w1, fmaf(-f.DAi[1], w0, tA1)));
    const float xA2 = fmaf(-f.DAi[5], w2, fmaf(-f.DAi[4], w1, fmaf(-f.DAi[2], w0, tA2)));
    x[1][0] = mk(xA0, xB0); x[1][1] = mk(xA1, xB1); x[1][2] = mk(xA2, xB2);
    // level 1 back: x_E = tE - DE^-1 (U_E x_O + [U_O' x_O](odd stage before))
    f2 mO[3];
    mO[0] = pfma(f.UO[3], x[1][1], pmul(f.UO[0], x[1][0]));
    mO[1] = pfma(f.UO[4], x[1][1], pmul(f.UO[1], x[1][0]));
    mO[2] = pfma(f.UO[5], x[1][2], pmul(f.UO[2], x[1][0]));
    f2 wE[3];
    wE[0] = pfma(f.UE[2], x[1][2], pfma(f.UE[1], x[1][1], pfma(f.UE[0], x[1][0], cm.prev_of(0, z2, mO[0]))));
    wE[1] = pfma(f.UE[4], x[1][1], pfma(f.UE[3], x[1][0], cm.prev_of(0, z2, mO[1])));
    wE[2] = pfma(f.UE[5], x[1][2], cm.prev_of(0, z2, mO[2]));
    x[0][0] = psub(tE[0], pfma(DEi[2], wE[2], pfma(DEi[1], wE[1], pmul(DEi[0], wE[0]))));
    x[0][1] = psub(tE[1], pfma(DEi[4], wE[2], pfma(DEi[3], wE[1], pmul(DEi[1], wE[0]))));
    x[0][2] = psub(tE[2], pfma(DEi[5], wE[2], pfma(DEi[4], wE[1], pmul(DEi[2], wE[0]))));
    // inputs (fv, fk are 0 where there is no successor)
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2 xn1 = cm.next_of(sl, x[0][1], x[1][1]), xn2 = cm.next_of(sl, x[0][2], x[1][2]);
        x[sl][3] = pmul(f.iv[sl], pfma(f.nfv[sl], xn2, pfma(f.nsxv2[sl], x[sl][2], pfma(f.nsxv0[sl], x[sl][0], b[sl][3]))));
        x[sl][4] = pmul(f.ik[sl], pfma(f.nfk[sl], xn1, pfma(f.nsxk1[sl], x[sl][1], pfma(f.nsxk0[sl], x[sl][0], b[sl][4]))));
    }
}

template <int LPS>
__device__ __forceinline__ unsigned rho_codes4(const f2* sm, int gl) {
    const float thr = (float)(kOsqpInfty * kMinScaling);
    unsigned codes = 0;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = j == 0 ? 0 : j + 2;
            const f2 lo = sm[qrow<LPS>(39 + i, sl, gl)], hi = sm[qrow<LPS>(44 + i, sl, gl)];
            const unsigned c0 = (lo.x < -thr && hi.x > thr) ? 2u : ((hi.x - lo.x < (float)kRhoTol) ? 1u : 0u);
            const unsigned c1 = (lo.y < -thr && hi.y > thr) ? 2u : ((hi.y - lo.y < (float)kRhoTol) ? 1u : 0u);
            codes |= c0 << (2 * (6 * sl + j)) | c1 << (2 * (6 * sl + 3 + j));
        }
    return codes;
}

// The OSQP loop, four stages per lane.  ALL lanes of the warp call this together (every group of LPS lanes = one scenario);
// control flow around the collectives is warp-uniform by voting, exactly as in admm_pair.cuh::admm_solve2.  `emit(w, result)`
// receives the UNSCALED primal stage vectors of both slices (w[0] = stages (4l, 4l + 2), w[1] = (4l + 1, 4l + 3)).
template <int LPS, typename Emit>
__device__ __forceinline__ void admm_solve4(const QuadComm<LPS>& cm, Stage2 (&s)[2], const AdmmSettings& st, const f2 al2,
                                            const f2 nal2, int nvar, f2* sm, float4* cf, bool live, Emit emit) {
    typedef GroupComm<LPS> GC;
    const int gl = cm.gl;
    if (st.scaling > 0) ruiz_scale4<LPS>(cm, s, st.scaling, nvar);
    else {
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
            for (int i = 0; i < 5; ++i) { s[sl].D[i] = bc(1.0f); s[sl].Eb[i] = bc(1.0f); }
#pragma unroll
            for (int i = 0; i < 3; ++i) s[sl].Ed[i] = bc(1.0f);
            s[sl].cs = 1.0f;
        }
    }
    const float thr = (float)(kOsqpInfty * kMinScaling);
    auto R = [&](int row, int sl) -> f2& { return sm[qrow<LPS>(row, sl, gl)]; };
    float nq_s = 0.0f, nq_u = 0.0f;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            R(i, sl) = s[sl].d[i]; R(8 + i, sl) = s[sl].Ed[i];
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            R(3 + i, sl) = s[sl].D[i]; R(11 + i, sl) = s[sl].Eb[i];
            R(29 + i, sl) = s[sl].q[i];
            if (i == 1 || i == 2) R(34 + i, sl) = s[sl].e[i];
            R(39 + i, sl) = s[sl].lo[i]; R(44 + i, sl) = s[sl].hi[i]; R(49 + i, sl) = s[sl].P[i];
            amax(nq_s, s[sl].q[i]); amax(nq_u, pmul(s[sl].q[i], prcp(s[sl].D[i])));
        }
    }
    const float cs0 = s[0].cs;
    R(54, 0) = mk(cm.max(nq_s), cm.max(nq_u));
    R(55, 0) = mk(cs0, 1.0f / cs0);
    float rho = (float)st.rho, rdf = (float)kRhoEqOverIneq * rho;
    const float sigma = (float)st.sigma;
    __syncwarp();
    const unsigned codes = rho_codes4<LPS>(sm, gl);
    QuadFactor<LPS> f;
    factorize4<LPS>(cm, s, f, sigma, rho, rdf, codes, sm, cf);
    const f2 zero = bc(0.0f);
    f2 x[2][5], u[2][5], vb[2][5], zb[2][5], rbd[2][5], rdy[2][3], vl4[2], zl4[2];
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 5; ++i) { x[sl][i] = zero; u[sl][i] = s[sl].q[i]; vb[sl][i] = zero; zb[sl][i] = zero; rbd[sl][i] = zero; }
#pragma unroll
        for (int i = 0; i < 3; ++i) rdy[sl][i] = zero;
        vl4[sl] = zero; zl4[sl] = zero;
    }
    f2 rd = bc(rdf);
    bool done = !live;
    int iter = 0;
    int chk = st.check_termination > 0 ? st.check_termination : -1;
    int adp = st.adaptive_rho_interval > 0 ? st.adaptive_rho_interval : -1;
    auto finish = [&](int status, int it) {
        f2 w[2][5];
        const bool nan_out = (status == -3 || status == -4 || status == -7 || status == 3 || status == 4);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 5; ++i) w[sl][i] = nan_out ? bc(NAN) : pmul(R(3 + i, sl), x[sl][i]);
        SolveResult r;
        r.iters = it;
        r.status = status;
        emit(w, r);
        done = true;
    };
    f2 dl[2][5], ed[2][3], eb[2][5];
    auto pass = [&](const bool first) __attribute__((always_inline)) {
        f2 td[2][3], tb[2][5], rhs[2][5], s1d[2][3], s1b[2][5];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
            for (int i = 0; i < 3; ++i) td[sl][i] = pmul(rd, rdy[sl][i]);
#pragma unroll
            for (int i = 0; i < 5; ++i) tb[sl][i] = (i == 1 || i == 2) ? zero : pmul(rho_row(codes, sl, i, rho, rdf), rbd[sl][i]);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float4 P2 = qrow2<LPS>(sm, 49 + i, gl);
            rhs[0][i] = pfma(slice_of(P2, 0), x[0][i], u[0][i]);
            rhs[1][i] = pfma(slice_of(P2, 1), x[1][i], u[1][i]);
        }
        At_apply4<LPS>(cm, s, td, tb, rhs, rhs);  // rhs = P x + u + A'(rho r);  S D = -rhs
        kkt_solve4<LPS>(cm, f, rhs, dl, cf);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 5; ++i) { dl[sl][i] = pmul(dl[sl][i], nal2); x[sl][i] = padd(x[sl][i], dl[sl][i]); }  // dl = alpha D
        A_apply4<LPS>(cm, s, dl, s1d, s1b);
        float4 lo4[5], hi4[5];
#pragma unroll
        for (int i = 0; i < 5; ++i)
            if (!(i == 1 || i == 2)) { lo4[i] = qrow2<LPS>(sm, 39 + i, gl); hi4[i] = qrow2<LPS>(sm, 44 + i, gl); }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const f2 wv = pfma(al2, rdy[sl][i], s1d[sl][i]);  // v - z_prev
                rdy[sl][i] = padd(rdy[sl][i], s1d[sl][i]);
                ed[sl][i] = pmul(rd, wv);                         // dy of the dynamics rows
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i == 1 || i == 2) { eb[sl][i] = zero; continue; }
                const f2 rbi = rho_row(codes, sl, i, rho, rdf);
                const f2 lo_i = slice_of(lo4[i], sl), hi_i = slice_of(hi4[i], sl);
                const f2 wv = pfma(al2, rbd[sl][i], s1b[sl][i]);
                if (MPC_COMPENSATED_V && i == 4) {
                    const f2 vs = padd(vb[sl][i], wv), bb = psub(vs, vb[sl][i]);  // TwoSum(v, w)
                    vl4[sl] = padd(vl4[sl], padd(psub(vb[sl][i], psub(vs, bb)), psub(wv, bb)));
                    vb[sl][i] = vs;
                } else {
                    vb[sl][i] = padd(vb[sl][i], wv);
                }
                const f2 zn = pmin(pmax(vb[sl][i], lo_i), hi_i);
                const f2 step = psub(zn, zb[sl][i]);
                zb[sl][i] = zn;
                if (MPC_COMPENSATED_V && i == 4) {
                    const f2 zln = mk(zn.x == vb[sl][i].x ? vl4[sl].x : 0.0f, zn.y == vb[sl][i].y ? vl4[sl].y : 0.0f);
                    const f2 stl = psub(zln, zl4[sl]);
                    zl4[sl] = zln;
                    rbd[sl][i] = psub(psub(padd(rbd[sl][i], s1b[sl][i]), step), stl);
                    eb[sl][i] = pmul(rbi, psub(psub(wv, step), stl));
                } else {
                    rbd[sl][i] = psub(padd(rbd[sl][i], s1b[sl][i]), step);
                    eb[sl][i] = pmul(rbi, psub(wv, step));
                }
            }
            if (first) {  // iteration 1: the dynamics z jumped from the cold start 0 to d
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const f2 dd = ldsv(&R(i, sl));
                    rdy[sl][i] = psub(rdy[sl][i], dd);
                    ed[sl][i] = psub(ed[sl][i], pmul(rd, dd));
                }
            }
        }
        At_apply4<LPS>(cm, s, ed, eb, u, u);
    };
    int phase = 0;  // 0 iterating, 1 final normal check, 2 final approximate check (see admm_pair.cuh)
    float tol = 1.0f;
    auto after_pass = [&]() __attribute__((always_inline)) -> bool {
        bool can_check = true, can_adapt = false;
        if (phase == 0) {
            can_check = (--chk == 0); can_adapt = (--adp == 0);
            if (can_check) chk = st.check_termination;
            if (can_adapt) adp = st.adaptive_rho_interval;
        }
        if (can_check || can_adapt) {
            f2 axd[2][3], axb[2][5];
            A_apply4<LPS>(cm, s, x, axd, axb);
            float pr_s = 0, pr_u = 0, nz_s = 0, nz_u = 0, nax_s = 0, nax_u = 0;
            float du_s = 0, du_u = 0, npx_s = 0, npx_u = 0, naty_s = 0, naty_u = 0;
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) {
                // loose rows: z follows A x
                axb[sl][1] = pmul(ldsv(&R(35, sl)), x[sl][1]); axb[sl][2] = pmul(ldsv(&R(36, sl)), x[sl][2]);
                zb[sl][1] = axb[sl][1]; zb[sl][2] = axb[sl][2];
#pragma unroll
                for (int i = 0; i < 3; ++i) {
                    const f2 zd = ldsv(&R(i, sl)), Edi = prcp(ldsv(&R(8 + i, sl)));
                    const f2 r = psub(axd[sl][i], zd);
                    amax(pr_s, r); amax(pr_u, pmul(r, Edi));
                    amax(nz_s, zd); amax(nz_u, pmul(zd, Edi));
                    amax(nax_s, axd[sl][i]); amax(nax_u, pmul(axd[sl][i], Edi));
                }
#pragma unroll
                for (int i = 0; i < 5; ++i) {
                    const f2 Ebi = prcp(ldsv(&R(11 + i, sl))), Di = prcp(ldsv(&R(3 + i, sl)));
                    const f2 r = psub(axb[sl][i], zb[sl][i]);
                    amax(pr_s, r); amax(pr_u, pmul(r, Ebi));
                    amax(nz_s, zb[sl][i]); amax(nz_u, pmul(zb[sl][i], Ebi));
                    amax(nax_s, axb[sl][i]); amax(nax_u, pmul(axb[sl][i], Ebi));
                    const f2 px = pmul(ldsv(&R(49 + i, sl)), x[sl][i]);
                    const f2 rr = padd(px, u[sl][i]);
                    const f2 aty = psub(u[sl][i], ldsv(&R(29 + i, sl)));
                    amax(du_s, rr); amax(du_u, pmul(rr, Di));
                    amax(npx_s, px); amax(npx_u, pmul(px, Di));
                    amax(naty_s, aty); amax(naty_u, pmul(aty, Di));
                }
            }
            const float nq_s = ldsv(&R(54, 0)).x, nq_u = ldsv(&R(54, 0)).y, cs = ldsv(&R(55, 0)).x, cinv = ldsv(&R(55, 0)).y;
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
                const bool open = !done && status == 0;
                bool pinf = false, dinf = false;
                if (GC::warp_any(open && !prim_ok)) {  // is_primal_infeasible
                    const float epi = tol * (float)st.eps_prim_inf;
                    f2 pyb[2][5];
                    float ndy = 0, lhs = 0;
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            amax(ndy, pmul(ldsv(&R(8 + i, sl)), ed[sl][i]));
                            const f2 t = pmul(ldsv(&R(i, sl)), ed[sl][i]);
                            lhs += t.x + t.y;
                        }
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            if (i == 1 || i == 2) { pyb[sl][i] = zero; continue; }
                            float dv[2] = {eb[sl][i].x, eb[sl][i].y};
                            const f2 lo2 = ldsv(&R(39 + i, sl)), hi2 = ldsv(&R(44 + i, sl));
                            const float lov[2] = {lo2.x, lo2.y}, hiv[2] = {hi2.x, hi2.y};
#pragma unroll
                            for (int h = 0; h < 2; ++h) {
                                float d = dv[h];
                                if (hiv[h] > thr) d = (lov[h] < -thr) ? 0.0f : fminf(d, 0.0f);
                                else if (lov[h] < -thr) d = fmaxf(d, 0.0f);
                                dv[h] = d;
                                lhs += hiv[h] * fmaxf(d, 0.0f) + lov[h] * fminf(d, 0.0f);
                            }
                            pyb[sl][i] = mk(dv[0], dv[1]);
                            amax(ndy, pmul(ldsv(&R(11 + i, sl)), pyb[sl][i]));
                        }
                    }
                    ndy = cm.max(ndy);
                    lhs = cm.sum(lhs);
                    const bool cand = open && !prim_ok && ndy > epi && lhs < -epi * ndy;
                    if (GC::warp_any(cand)) {
                        f2 atdy[2][5], z5[2][5];
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                            for (int i = 0; i < 5; ++i) z5[sl][i] = zero;
                        float na = 0;
                        At_apply4<LPS>(cm, s, ed, pyb, z5, atdy);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                            for (int i = 0; i < 5; ++i) amax(na, pmul(atdy[sl][i], prcp(ldsv(&R(3 + i, sl)))));
                        na = cm.max(na);
                        pinf = cand && na < epi * ndy;
                    }
                }
                if (GC::warp_any(open && !dual_ok && !pinf)) {  // is_dual_infeasible (dx = alpha D of this iteration)
                    const float edi = tol * (float)st.eps_dual_inf;
                    float ndx = 0, qdx = 0, npdx = 0;
#pragma unroll
                    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            amax(ndx, pmul(ldsv(&R(3 + i, sl)), dl[sl][i]));
                            const f2 t = pmul(ldsv(&R(29 + i, sl)), dl[sl][i]);
                            qdx += t.x + t.y;
                            amax(npdx, pmul(pmul(ldsv(&R(49 + i, sl)), dl[sl][i]), prcp(ldsv(&R(3 + i, sl)))));
                        }
                    ndx = cm.max(ndx);
                    qdx = cm.sum(qdx);
                    npdx = cm.max(npdx);
                    const bool cand = open && !dual_ok && !pinf && ndx > edi && qdx < -cs * edi * ndx && npdx < cs * edi * ndx;
                    if (GC::warp_any(cand)) {
                        f2 adxd[2][3], adxb[2][5];
                        A_apply4<LPS>(cm, s, dl, adxd, adxb);
                        int bad = 0;
                        const float lim = edi * ndx;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            adxb[sl][1] = pmul(ldsv(&R(35, sl)), dl[sl][1]); adxb[sl][2] = pmul(ldsv(&R(36, sl)), dl[sl][2]);
#pragma unroll
                            for (int i = 0; i < 3; ++i) {
                                const f2 v = pmul(adxd[sl][i], prcp(ldsv(&R(8 + i, sl))));
                                if (fabsf(v.x) > lim || fabsf(v.y) > lim) bad = 1;
                            }
#pragma unroll
                            for (int i = 0; i < 5; ++i) {
                                const f2 v = pmul(adxb[sl][i], prcp(ldsv(&R(11 + i, sl))));
                                const f2 lo2 = ldsv(&R(39 + i, sl)), hi2 = ldsv(&R(44 + i, sl));
                                if ((hi2.x < thr && v.x > lim) || (lo2.x > -thr && v.x < -lim)) bad = 1;
                                if ((hi2.y < thr && v.y > lim) || (lo2.y > -thr && v.y < -lim)) bad = 1;
                            }
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
            if (can_adapt) {
                const float pn = pr_s / (fmaxf(nz_s, nax_s) + 1e-10f);
                const float dn = du_s / (fmaxf(fmaxf(nq_s, naty_s), npx_s) + 1e-10f);
                float rnew = rho * sqrtf(pn / (dn + 1e-10f));
                rnew = fminf(fmaxf(rnew, (float)kRhoMin), (float)kRhoMax);
                const bool upd = !done && (rnew > rho * (float)st.adaptive_rho_tolerance ||
                                           rnew < rho / (float)st.adaptive_rho_tolerance);
                if (GC::warp_any(upd)) {
                    if (upd) {
                        const float ratio = rho / rnew;  // y is unchanged: v = z + (v - z) rho_old / rho_new
                        rho = rnew;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
                            for (int i = 0; i < 5; ++i) {
                                if (i == 1 || i == 2) continue;
                                const f2 lo2 = ldsv(&R(39 + i, sl)), hi2 = ldsv(&R(44 + i, sl));
                                const f2 rr = mk((lo2.x < -thr && hi2.x > thr) ? 1.0f : ratio,
                                                 (lo2.y < -thr && hi2.y > thr) ? 1.0f : ratio);
                                if (MPC_COMPENSATED_V && i == 4) {
                                    vb[sl][i] = pfma(padd(psub(vb[sl][i], zb[sl][i]), psub(vl4[sl], zl4[sl])), rr, zb[sl][i]);
                                    vl4[sl] = zl4[sl];
                                } else {
                                    vb[sl][i] = pfma(psub(vb[sl][i], zb[sl][i]), rr, zb[sl][i]);
                                }
                            }
                    }
                    rdf = (float)kRhoEqOverIneq * rho;
                    rd = bc(rdf);
                    factorize4<LPS>(cm, s, f, sigma, rho, rdf, codes, sm, cf);
                }
            }
        }
        return false;
    };
    for (iter = 1;; ++iter) {
        if (phase == 0) pass(iter == 1);
        if (after_pass()) break;
        if (phase == 1 || (phase == 0 && iter >= st.max_iter)) {
            const bool checked = phase == 0 && st.check_termination > 0 && (st.max_iter % st.check_termination == 0);
            phase = (phase == 1 || checked) ? 2 : 1;
            if (phase == 2) tol = 10.0f;
        }
    }
}

}  // namespace mpcb
