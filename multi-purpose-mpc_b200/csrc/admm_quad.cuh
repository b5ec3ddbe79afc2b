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

// Memory plan of a warp (four scenarios), all indices per lane:
//   HOT  = what every pass reads, in shared memory as float4 columns [k][32 lanes] (LDS.128, conflict-free; a float4 holds the
//          E-slice and the O-slice value of one quantity, or four scalars): 43 float4 = 688 B per lane, 22 KB per warp, so
//          that eight warps fit an SM.  The iterate itself and the stage matrices a, c, e stay in registers.
//   COLD = what only the termination checks, the certificates, the re-factorisation and the result read (scalings, q, d,
//          the e of the loose rows, norms): global memory, float4 columns [k][32 lanes] per warp, L2-resident.
enum : int {
    kHP = 0,     // 5: P
    kHLo = 5,    // 3: lo of the bound rows 0, 3, 4
    kHHi = 8,    // 3: hi
    kHEl = 11,   // 8: input elimination iv, ik, -sxv0, -sxv2, -sxk0, -sxk1, -fv, -fk
    kHL1 = 19,   // 9: level 1, 18 f2 = DE^-1 (00 01 02 11 12 22), U_E (0 1 2 3 4 8), U_O (0 1 2 3 4 8), two f2 per float4
    kHL2 = 28,   // 6: level 2, 24 floats = DA^-1 (6), UA (9), UB (9)
    kHEnd = 34,  // 4: PCR last level (9 floats) and the final D^-1 (00 01 02 11 12 22), 15 floats
    kHC = 38,    // 3: c (the stage's own -1 entries of the dynamics rows, scaled)
    kHE = 41,    // 3: e of the bound rows 0, 3, 4
    kHPcr = 44,  // PcrCoef<LPS>::kF4: PCR levels 0 .. NLEV-2
};
template <int LPS> struct QuadHot { static constexpr int kF4 = kHPcr + PcrCoef<LPS>::kF4; };
enum : int {
    kCd = 0,     // 3: d
    kCD = 3,     // 5: D
    kCEd = 8,    // 3: Ed
    kCEb = 11,   // 5: Eb
    kCq = 16,    // 5: q
    kCel = 21,   // 2: e of the loose rows 1, 2
    kCmisc = 23, // 1: (|q| scaled, |q| unscaled, c, 1 / c)
    kCdl = 24,   // 5: alpha D of the pass in front of a termination check (dx of the dual-infeasibility certificate)
    kCed = 29,   // 3: dy of the dynamics rows of that pass
    kCeb = 32,   // 3: dy of the bound rows 0, 3, 4 of that pass
    kQuadCold = 35,
};
__device__ __forceinline__ f2 slice_of(const float4& v, int sl) { return sl == 0 ? mk(v.x, v.y) : mk(v.z, v.w); }
__device__ __forceinline__ float4 both(f2 e, f2 o) { return make_float4(e.x, e.y, o.x, o.y); }
__device__ __forceinline__ void smem_fence() { asm volatile("" ::: "memory"); }
// LDS.128 that cannot be scheduled before `dep` exists: its address carries (bits(dep) & zero), `zero` being a kernel argument
// that is always 0 -- the compiler cannot fold it.  The HOT columns are re-read where they are used; without the pin the
// scheduler hoists all of a pass's 60-odd loads to its top and the register allocator parks the iterate in local memory.
__device__ __forceinline__ float4 lds128_after(const float4* p, float dep, int zero) {
    float4 v;
    const unsigned a = (unsigned)__cvta_generic_to_shared(p) + (unsigned)(__float_as_int(dep) & zero);
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
// rho of the bound rows 0 / 3 / 4 of the four stages, as 2-bit codes (0 inequality: rho, 1 equality-typed: 1e3 rho, 2 loose:
// rho_min) in one register instead of six float2: bit pair 2 * (6 sl + 3 half + j), j = 0, 1, 2 for rows 0, 3, 4
__device__ __forceinline__ float rho_code(unsigned codes, int sl, int half, int j, float rho, float rdf) {
    const unsigned c = codes >> (2 * (6 * sl + 3 * half + j));
    const float r = (c & 1u) ? rdf : rho;  // two selects, no branch
    return (c & 2u) ? (float)kRhoMin : r;
}
__device__ __forceinline__ f2 rho_row(unsigned codes, int sl, int i, float rho, float rdf) {
    const int j = i == 0 ? 0 : i - 2;
    return mk(rho_code(codes, sl, 0, j, rho, rdf), rho_code(codes, sl, 1, j, rho, rdf));
}

template <int LPS> struct QuadFactor {
    static constexpr int NLEV = PairFactor<LPS>::NLEV;
    // no data members: the whole factor lives in the HOT columns (the type only carries NLEV)
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
                                         f2 (&zb)[2][5], const float4* hot, int zpin) {
    f2 o[2][3];
    float4 c4[3], e4[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c4[i] = lds128_after(hot + (kHC + i) * 32, w[0][0].x, zpin);
        e4[i] = lds128_after(hot + (kHE + i) * 32, w[0][0].x, zpin);
    }
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
        for (int i = 0; i < 3; ++i) zd[sl][i] = pfma(slice_of(c4[i], sl), w[sl][i], cm.prev_of(sl, o[0][i], o[1][i]));
        zb[sl][0] = pmul(slice_of(e4[0], sl), w[sl][0]);
        zb[sl][3] = pmul(slice_of(e4[1], sl), w[sl][3]);
        zb[sl][4] = pmul(slice_of(e4[2], sl), w[sl][4]);
    }
}

// r = acc + A' y (yb[1], yb[2] are identically zero and not read)
template <int LPS>
__device__ __forceinline__ void At_apply4(const QuadComm<LPS>& cm, const Stage2 (&s)[2], const f2 (&yd)[2][3],
                                          const f2 (&yb)[2][5], const f2 (&acc)[2][5], f2 (&r)[2][5], const float4* hot, int zpin) {
    float4 c4[3], e4[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        c4[i] = lds128_after(hot + (kHC + i) * 32, yd[0][0].x, zpin);
        e4[i] = lds128_after(hot + (kHE + i) * 32, yd[0][0].x, zpin);
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2* a = s[sl].a;
        const f2 g0 = cm.next_of(sl, yd[0][0], yd[1][0]), g1 = cm.next_of(sl, yd[0][1], yd[1][1]),
                 g2 = cm.next_of(sl, yd[0][2], yd[1][2]);
        const f2 r0 = pfma(a[4], g2, pfma(a[2], g1, pfma(a[0], g0, pfma(slice_of(c4[0], sl), yd[sl][0], pfma(slice_of(e4[0], sl), yb[sl][0], acc[sl][0])))));
        const f2 r1 = pfma(a[3], g1, pfma(a[1], g0, pfma(slice_of(c4[1], sl), yd[sl][1], acc[sl][1])));
        const f2 r2 = pfma(a[5], g2, pfma(slice_of(c4[2], sl), yd[sl][2], acc[sl][2]));
        const f2 r3 = pfma(a[7], g2, pfma(slice_of(e4[1], sl), yb[sl][3], acc[sl][3]));
        const f2 r4 = pfma(a[6], g1, pfma(slice_of(e4[2], sl), yb[sl][4], acc[sl][4]));
        r[sl][0] = r0; r[sl][1] = r1; r[sl][2] = r2; r[sl][3] = r3; r[sl][4] = r4;
    }
}

// S = P + sigma I + A' R A for the four stages of the lane; inputs eliminated; E stages eliminated (level 1, packed);
// s1 eliminated (level 2); PCR factorisation of the s3 chain.  Everything a pass needs of it goes to the HOT columns.
template <int LPS>
__device__ __forceinline__ void factorize4(const QuadComm<LPS>& cm, const Stage2 (&s)[2], QuadFactor<LPS>& f, float sigma,
                                           float rho, float rdf, unsigned codes, float4* hot, const float4* cold) {
    constexpr int NLEV = QuadFactor<LPS>::NLEV;
    const f2 rd = bc(rdf), sg = bc(sigma);
    f2 D00[2], D01[2], D02[2], D11[2], D22[2], U0[2], U1[2], U2[2], U3[2], U4[2], U8[2], tkk[2], tvv[2];
    f2 el[8][2];
    float4 P4[5], el4[2];
#pragma unroll
    for (int i = 0; i < 5; ++i) P4[i] = lds128v(hot + (kHP + i) * 32);
    el4[0] = cold[(kCel + 0) * 32]; el4[1] = cold[(kCel + 1) * 32];
    float4 c4[3], e4[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) { c4[i] = lds128v(hot + (kHC + i) * 32); e4[i] = lds128v(hot + (kHE + i) * 32); }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        const f2* a = s[sl].a;
        f2 diag[5];
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const bool lz = (i == 1 || i == 2);  // loose rows: rho = rho_min, their e is COLD
            const f2 ei = lz ? slice_of(el4[i == 1 ? 0 : 1], sl) : slice_of(e4[i == 0 ? 0 : i - 2], sl);
            const f2 ri = lz ? bc((float)kRhoMin) : rho_row(codes, sl, i, rho, rdf);
            diag[i] = pfma(pmul(ri, ei), ei, padd(slice_of(P4[i], sl), sg));
        }
        f2 cn[3];
#pragma unroll
        for (int i = 0; i < 3; ++i) cn[i] = cm.next_of(sl, slice_of(c4[i], 0), slice_of(c4[i], 1));
        const f2 c[3] = {slice_of(c4[0], sl), slice_of(c4[1], sl), slice_of(c4[2], sl)};
        D00[sl] = pfma(rd, pfma(a[4], a[4], pfma(a[2], a[2], pfma(a[0], a[0], pmul(c[0], c[0])))), diag[0]);
        D11[sl] = pfma(rd, pfma(a[3], a[3], pfma(a[1], a[1], pmul(c[1], c[1]))), diag[1]);
        D22[sl] = pfma(rd, pfma(a[5], a[5], pmul(c[2], c[2])), diag[2]);
        D01[sl] = pmul(rd, pfma(a[2], a[3], pmul(a[0], a[1])));
        D02[sl] = pmul(rd, pmul(a[4], a[5]));
        const f2 Svv = pfma(rd, pmul(a[7], a[7]), diag[3]);
        const f2 Skk = pfma(rd, pmul(a[6], a[6]), diag[4]);
        const f2 iv = prcp(Svv), ik = prcp(Skk);
        const f2 ra7 = pmul(rd, a[7]), ra6 = pmul(rd, a[6]);
        const f2 sxv0 = pmul(ra7, a[4]), sxv2 = pmul(ra7, a[5]);
        const f2 sxk0 = pmul(ra6, a[2]), sxk1 = pmul(ra6, a[3]);
        const f2 fv = pmul(ra7, cn[2]), fk = pmul(ra6, cn[1]);
        const f2 rc0 = pmul(rd, cn[0]), rc1 = pmul(rd, cn[1]), rc2 = pmul(rd, cn[2]);
        U0[sl] = pmul(a[0], rc0); U1[sl] = pmul(a[2], rc1); U2[sl] = pmul(a[4], rc2);
        U3[sl] = pmul(a[1], rc0); U4[sl] = pmul(a[3], rc1); U8[sl] = pmul(a[5], rc2);
        const f2 ivs0 = pmul(iv, sxv0), ivs2 = pmul(iv, sxv2), iks0 = pmul(ik, sxk0), iks1 = pmul(ik, sxk1);
        D00[sl] = psub(D00[sl], pfma(iks0, sxk0, pmul(ivs0, sxv0)));
        D01[sl] = psub(D01[sl], pmul(iks0, sxk1));
        D02[sl] = psub(D02[sl], pmul(ivs0, sxv2));
        D11[sl] = psub(D11[sl], pmul(iks1, sxk1));
        D22[sl] = psub(D22[sl], pmul(ivs2, sxv2));
        U2[sl] = psub(U2[sl], pmul(ivs0, fv)); U8[sl] = psub(U8[sl], pmul(ivs2, fv));
        U1[sl] = psub(U1[sl], pmul(iks0, fk)); U4[sl] = psub(U4[sl], pmul(iks1, fk));
        tkk[sl] = pmul(pmul(ik, fk), fk);
        tvv[sl] = pmul(pmul(iv, fv), fv);
        const f2 m1 = bc(-1.0f);
        el[0][sl] = iv; el[1][sl] = ik; el[2][sl] = pmul(sxv0, m1); el[3][sl] = pmul(sxv2, m1); el[4][sl] = pmul(sxk0, m1);
        el[5][sl] = pmul(sxk1, m1); el[6][sl] = pmul(fv, m1); el[7][sl] = pmul(fk, m1);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) hot[(kHEl + k) * 32] = both(el[k][0], el[k][1]);
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {  // what the previous stage's input elimination leaves on this stage's diagonal
        D11[sl] = psub(D11[sl], cm.prev_of(sl, tkk[0], tkk[1]));
        D22[sl] = psub(D22[sl], cm.prev_of(sl, tvv[0], tvv[1]));
    }
    // ---- level 1: eliminate the E stages (packed) ----
    f2 DEi[6];
    {
        const f2 DE[6] = {D00[0], D01[0], D02[0], D11[0], bc(0.0f), D22[0]};
        inv3sym6p(DE, DEi);
    }
    {
        const f2 l1[18] = {DEi[0], DEi[1], DEi[2], DEi[3], DEi[4], DEi[5], U0[0], U1[0], U2[0], U3[0], U4[0], U8[0],
                           U0[1], U1[1], U2[1], U3[1], U4[1], U8[1]};
#pragma unroll
        for (int k = 0; k < 9; ++k) hot[(kHL1 + k) * 32] = both(l1[2 * k], l1[2 * k + 1]);
    }
    const f2 z2 = bc(0.0f);
    const f2 Di9[9] = {DEi[0], DEi[1], DEi[2], DEi[1], DEi[3], DEi[4], DEi[2], DEi[4], DEi[5]};
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
    for (int i = 0; i < 9; ++i) { DA9[i] = DO9[i].x; DB9[i] = DO9[i].y; UA[i] = UOn[i].x; UB[i] = UOn[i].y; }
    float DAi9[9];
    inv3sym(DA9, DAi9);
    {
        const float l2[24] = {DAi9[0], DAi9[1], DAi9[2], DAi9[4], DAi9[5], DAi9[8], UA[0], UA[1], UA[2], UA[3], UA[4], UA[5],
                              UA[6], UA[7], UA[8], UB[0], UB[1], UB[2], UB[3], UB[4], UB[5], UB[6], UB[7], UB[8]};
#pragma unroll
        for (int k = 0; k < 6; ++k) hot[(kHL2 + k) * 32] = make_float4(l2[4 * k], l2[4 * k + 1], l2[4 * k + 2], l2[4 * k + 3]);
    }
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
    float4* cf = hot + kHPcr * 32;
    float last[9];
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
                last[i] = has_up ? al[i] : be[i];
            }
        }
    }
    float Di[9];
    inv3sym(Dm, Di);
    hot[(kHEnd + 0) * 32] = make_float4(last[0], last[1], last[2], last[3]);
    hot[(kHEnd + 1) * 32] = make_float4(last[4], last[5], last[6], last[7]);
    hot[(kHEnd + 2) * 32] = make_float4(last[8], Di[0], Di[1], Di[2]);
    hot[(kHEnd + 3) * 32] = make_float4(Di[4], Di[5], Di[8], 0.0f);
    smem_fence();
}

// x = S^-1 b for the four stages of the lane
template <int LPS>
__device__ __forceinline__ void kkt_solve4(const QuadComm<LPS>& cm, const QuadFactor<LPS>& f, const f2 (&b)[2][5], f2 (&x)[2][5],
                                           const float4* hot, int zero) {
    constexpr int NLEV = QuadFactor<LPS>::NLEV;
    const f2 z2 = bc(0.0f);
    // input elimination
    f2 bx[2][3], tk[2], tv[2];
    {
        float4 e4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e4[k] = lds128_after(hot + (kHEl + k) * 32, b[0][3].x, zero);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const f2 bv = pmul(slice_of(e4[0], sl), b[sl][3]), bk = pmul(slice_of(e4[1], sl), b[sl][4]);
            bx[sl][0] = pfma(bk, slice_of(e4[4], sl), pfma(bv, slice_of(e4[2], sl), b[sl][0]));
            bx[sl][1] = pfma(bk, slice_of(e4[5], sl), b[sl][1]);
            bx[sl][2] = pfma(bv, slice_of(e4[3], sl), b[sl][2]);
            tk[sl] = pmul(bk, slice_of(e4[7], sl));
            tv[sl] = pmul(bv, slice_of(e4[6], sl));
        }
    }
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
        bx[sl][1] = padd(bx[sl][1], cm.prev_of(sl, tk[0], tk[1]));
        bx[sl][2] = padd(bx[sl][2], cm.prev_of(sl, tv[0], tv[1]));
    }
    // level 1 forward: tE = DE^-1 b_E;  b_O' = b_O - U_E' tE - U_O tE(next even stage)
    f2 tE[3], rO[3];
    {
    f2 DEi[6], UE[6], UO[6];
    {
        float4 l1[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) l1[k] = lds128_after(hot + (kHL1 + k) * 32, bx[0][1].x, zero);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            DEi[2 * k] = mk(l1[k].x, l1[k].y); DEi[2 * k + 1] = mk(l1[k].z, l1[k].w);
            UE[2 * k] = mk(l1[3 + k].x, l1[3 + k].y); UE[2 * k + 1] = mk(l1[3 + k].z, l1[3 + k].w);
            UO[2 * k] = mk(l1[6 + k].x, l1[6 + k].y); UO[2 * k + 1] = mk(l1[6 + k].z, l1[6 + k].w);
        }
    }
    tE[0] = pfma(DEi[2], bx[0][2], pfma(DEi[1], bx[0][1], pmul(DEi[0], bx[0][0])));
    tE[1] = pfma(DEi[4], bx[0][2], pfma(DEi[3], bx[0][1], pmul(DEi[1], bx[0][0])));
    tE[2] = pfma(DEi[5], bx[0][2], pfma(DEi[4], bx[0][1], pmul(DEi[2], bx[0][0])));
    f2 tn[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) tn[i] = cm.next_of(1, tE[i], z2);
    // U = [u0 u1 u2; u3 u4 0; 0 0 u5]:  (U' t)_k = sum_i U[i][k] t_i,  (U t)_i = sum_k U[i][k] t_k
    rO[0] = psub(psub(bx[1][0], pfma(UE[3], tE[1], pmul(UE[0], tE[0]))), pfma(UO[2], tn[2], pfma(UO[1], tn[1], pmul(UO[0], tn[0]))));
    rO[1] = psub(psub(bx[1][1], pfma(UE[4], tE[1], pmul(UE[1], tE[0]))), pfma(UO[4], tn[1], pmul(UO[3], tn[0])));
    rO[2] = psub(psub(bx[1][2], pfma(UE[5], tE[2], pmul(UE[2], tE[0]))), pmul(UO[5], tn[2]));
    }
    // level 2 forward (scalar): tA = DA^-1 r_A;  r_B' = r_B - UA' tA - UB tA(next lane)
    float DAi[6], UA[9], UB[9];
    {
        float4 l2[6];
#pragma unroll
        for (int k = 0; k < 6; ++k) l2[k] = lds128_after(hot + (kHL2 + k) * 32, rO[0].x, zero);
        const float v[24] = {l2[0].x, l2[0].y, l2[0].z, l2[0].w, l2[1].x, l2[1].y, l2[1].z, l2[1].w, l2[2].x, l2[2].y, l2[2].z, l2[2].w,
                             l2[3].x, l2[3].y, l2[3].z, l2[3].w, l2[4].x, l2[4].y, l2[4].z, l2[4].w, l2[5].x, l2[5].y, l2[5].z, l2[5].w};
#pragma unroll
        for (int k = 0; k < 6; ++k) DAi[k] = v[k];
#pragma unroll
        for (int k = 0; k < 9; ++k) { UA[k] = v[6 + k]; UB[k] = v[15 + k]; }
    }
    const float rA0 = rO[0].x, rA1 = rO[1].x, rA2 = rO[2].x;
    const float tA0 = fmaf(DAi[2], rA2, fmaf(DAi[1], rA1, DAi[0] * rA0));
    const float tA1 = fmaf(DAi[4], rA2, fmaf(DAi[3], rA1, DAi[1] * rA0));
    const float tA2 = fmaf(DAi[5], rA2, fmaf(DAi[4], rA1, DAi[2] * rA0));
    const float n0 = cm.next(tA0), n1 = cm.next(tA1), n2 = cm.next(tA2);
    f2 R0 = mk(rO[0].y, 0.0f), R1 = mk(rO[1].y, 0.0f), R2 = mk(rO[2].y, 0.0f);
    R0.x -= fmaf(UA[6], tA2, fmaf(UA[3], tA1, UA[0] * tA0)) + fmaf(UB[2], n2, fmaf(UB[1], n1, UB[0] * n0));
    R1.x -= fmaf(UA[7], tA2, fmaf(UA[4], tA1, UA[1] * tA0)) + fmaf(UB[5], n2, fmaf(UB[4], n1, UB[3] * n0));
    R2.x -= fmaf(UA[8], tA2, fmaf(UA[5], tA1, UA[2] * tA0)) + fmaf(UB[8], n2, fmaf(UB[7], n1, UB[6] * n0));
    // PCR over the s3 chain (the coefficients of a level are read when the level starts)
    {
        const float4* cf = hot + kHPcr * 32;
#pragma unroll
        for (int lev = 0; lev < NLEV - 1; ++lev) {
            const int sft = 1 << lev;
            const f2 m0 = mk(cm.up(R0.x, sft), cm.dn(R0.x, sft));
            const f2 m1 = mk(cm.up(R1.x, sft), cm.dn(R1.x, sft));
            const f2 m2 = mk(cm.up(R2.x, sft), cm.dn(R2.x, sft));
            f2 nab[9];
            float4 cq[6];
            const int j0 = (9 * lev) >> 1;  // first float4 column of this level
#pragma unroll
            for (int k = 0; k < 6; ++k)
                if (j0 + k < PcrCoef<LPS>::kF4) cq[k] = lds128_after(cf + (j0 + k) * 32, R0.x, zero);
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int j = 9 * lev + i;
                const int c = (j >> 1) - j0;
                nab[i] = (j & 1) ? mk(cq[c].z, cq[c].w) : mk(cq[c].x, cq[c].y);
            }
            const f2 s0 = pfma(nab[2], m2, pfma(nab[1], m1, pfma(nab[0], m0, R0)));
            const f2 s1 = pfma(nab[5], m2, pfma(nab[4], m1, pfma(nab[3], m0, R1)));
            const f2 s2 = pfma(nab[8], m2, pfma(nab[7], m1, pfma(nab[6], m0, R2)));
            R0.x = s0.x + s0.y; R1.x = s1.x + s1.y; R2.x = s2.x + s2.y;
        }
    }
    float r0 = R0.x, r1 = R1.x, r2 = R2.x;
    float xB0, xB1, xB2;
    {
        float4 en[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) en[k] = lds128_after(hot + (kHEnd + k) * 32, r0, zero);
        const float last[9] = {en[0].x, en[0].y, en[0].z, en[0].w, en[1].x, en[1].y, en[1].z, en[1].w, en[2].x};
        const float Dinv[6] = {en[2].y, en[2].z, en[2].w, en[3].x, en[3].y, en[3].z};
        const int sft = LPS / 2;
        const float q0 = cm.bfly(r0, sft), q1 = cm.bfly(r1, sft), q2 = cm.bfly(r2, sft);
        r0 = fmaf(-last[2], q2, fmaf(-last[1], q1, fmaf(-last[0], q0, r0)));
        r1 = fmaf(-last[5], q2, fmaf(-last[4], q1, fmaf(-last[3], q0, r1)));
        r2 = fmaf(-last[8], q2, fmaf(-last[7], q1, fmaf(-last[6], q0, r2)));
        xB0 = fmaf(Dinv[2], r2, fmaf(Dinv[1], r1, Dinv[0] * r0));
        xB1 = fmaf(Dinv[4], r2, fmaf(Dinv[3], r1, Dinv[1] * r0));
        xB2 = fmaf(Dinv[5], r2, fmaf(Dinv[4], r1, Dinv[2] * r0));
    }
    // level 2 back: x_A = tA - DA^-1 (UA x_B + [UB' x_B](previous lane))
    const float mB0 = fmaf(UB[6], xB2, fmaf(UB[3], xB1, UB[0] * xB0));
    const float mB1 = fmaf(UB[7], xB2, fmaf(UB[4], xB1, UB[1] * xB0));
    const float mB2 = fmaf(UB[8], xB2, fmaf(UB[5], xB1, UB[2] * xB0));
    const float w0 = fmaf(UA[2], xB2, fmaf(UA[1], xB1, fmaf(UA[0], xB0, cm.prev(mB0))));
    const float w1 = fmaf(UA[5], xB2, fmaf(UA[4], xB1, fmaf(UA[3], xB0, cm.prev(mB1))));
    const float w2 = fmaf(UA[8], xB2, fmaf(UA[7], xB1, fmaf(UA[6], xB0, cm.prev(mB2))));
    const float xA0 = fmaf(-DAi[2], w2, fmaf(-DAi[1], w1, fmaf(-DAi[0], w0, tA0)));
    const float xA1 = fmaf(-DAi[4], w2, fmaf(-DAi[3], w1, fmaf(-DAi[1], w0, tA1)));
    const float xA2 = fmaf(-DAi[5], w2, fmaf(-DAi[4], w1, fmaf(-DAi[2], w0, tA2)));
    x[1][0] = mk(xA0, xB0); x[1][1] = mk(xA1, xB1); x[1][2] = mk(xA2, xB2);
    // level 1 back: x_E = tE - DE^-1 (U_E x_O + [U_O' x_O](odd stage before))   (coefficients re-read)
    f2 DEi[6], UE[6], UO[6];
    {
        float4 l1[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) l1[k] = lds128_after(hot + (kHL1 + k) * 32, xA0, zero);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            DEi[2 * k] = mk(l1[k].x, l1[k].y); DEi[2 * k + 1] = mk(l1[k].z, l1[k].w);
            UE[2 * k] = mk(l1[3 + k].x, l1[3 + k].y); UE[2 * k + 1] = mk(l1[3 + k].z, l1[3 + k].w);
            UO[2 * k] = mk(l1[6 + k].x, l1[6 + k].y); UO[2 * k + 1] = mk(l1[6 + k].z, l1[6 + k].w);
        }
    }
    f2 mO[3];
    mO[0] = pfma(UO[3], x[1][1], pmul(UO[0], x[1][0]));
    mO[1] = pfma(UO[4], x[1][1], pmul(UO[1], x[1][0]));
    mO[2] = pfma(UO[5], x[1][2], pmul(UO[2], x[1][0]));
    f2 wE[3];
    wE[0] = pfma(UE[2], x[1][2], pfma(UE[1], x[1][1], pfma(UE[0], x[1][0], cm.prev_of(0, z2, mO[0]))));
    wE[1] = pfma(UE[4], x[1][1], pfma(UE[3], x[1][0], cm.prev_of(0, z2, mO[1])));
    wE[2] = pfma(UE[5], x[1][2], cm.prev_of(0, z2, mO[2]));
    x[0][0] = psub(tE[0], pfma(DEi[2], wE[2], pfma(DEi[1], wE[1], pmul(DEi[0], wE[0]))));
    x[0][1] = psub(tE[1], pfma(DEi[4], wE[2], pfma(DEi[3], wE[1], pmul(DEi[1], wE[0]))));
    x[0][2] = psub(tE[2], pfma(DEi[5], wE[2], pfma(DEi[4], wE[1], pmul(DEi[2], wE[0]))));
    // inputs (fv, fk are 0 where there is no successor)
    {
        float4 e4[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) e4[k] = lds128_after(hot + (kHEl + k) * 32, x[0][0].x, zero);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
            const f2 xn1 = cm.next_of(sl, x[0][1], x[1][1]), xn2 = cm.next_of(sl, x[0][2], x[1][2]);
            x[sl][3] = pmul(slice_of(e4[0], sl), pfma(slice_of(e4[6], sl), xn2, pfma(slice_of(e4[3], sl), x[sl][2], pfma(slice_of(e4[2], sl), x[sl][0], b[sl][3]))));
            x[sl][4] = pmul(slice_of(e4[1], sl), pfma(slice_of(e4[7], sl), xn1, pfma(slice_of(e4[5], sl), x[sl][1], pfma(slice_of(e4[4], sl), x[sl][0], b[sl][4]))));
        }
    }
}

__device__ __forceinline__ unsigned rho_codes4(const Stage2 (&s)[2]) {
    const float thr = (float)(kOsqpInfty * kMinScaling);
    unsigned codes = 0;
#pragma unroll
    for (int sl = 0; sl < 2; ++sl)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const int i = j == 0 ? 0 : j + 2;
            const f2 lo = s[sl].lo[i], hi = s[sl].hi[i];
            const unsigned c0 = (lo.x < -thr && hi.x > thr) ? 2u : ((hi.x - lo.x < (float)kRhoTol) ? 1u : 0u);
            const unsigned c1 = (lo.y < -thr && hi.y > thr) ? 2u : ((hi.y - lo.y < (float)kRhoTol) ? 1u : 0u);
            codes |= c0 << (2 * (6 * sl + j)) | c1 << (2 * (6 * sl + 3 + j));
        }
    return codes;
}

// The OSQP loop, four stages per lane.  ALL lanes of the warp call this together (every group of LPS lanes = one scenario);
// control flow around the collectives is warp-uniform by voting, exactly as in admm_pair.cuh::admm_solve2.  `emit(w, result)`
// receives the UNSCALED primal stage vectors of both slices (w[0] = stages (4l, 4l + 2), w[1] = (4l + 1, 4l + 3)).
// hot / cold: this LANE's float4 columns (stride 32 float4 between consecutive entries), see the memory plan above.
template <int LPS, typename Emit>
__device__ __forceinline__ void admm_solve4(const QuadComm<LPS>& cm, Stage2 (&s)[2], const AdmmSettings& st, const f2 al2,
                                            const f2 nal2, int nvar, float4* hot, float4* cold, bool live, int zpin, Emit emit) {
    typedef GroupComm<LPS> GC;
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
    auto C = [&](int k) -> float4& { return cold[k * 32]; };
    auto H = [&](int k) -> float4 { return lds128v(hot + k * 32); };
    float nq_s = 0.0f, nq_u = 0.0f;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        C(kCd + i) = both(s[0].d[i], s[1].d[i]);
        C(kCEd + i) = both(s[0].Ed[i], s[1].Ed[i]);
    }
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        C(kCD + i) = both(s[0].D[i], s[1].D[i]);
        C(kCEb + i) = both(s[0].Eb[i], s[1].Eb[i]);
        C(kCq + i) = both(s[0].q[i], s[1].q[i]);
        hot[(kHP + i) * 32] = both(s[0].P[i], s[1].P[i]);
        if (i == 1 || i == 2) C(kCel + i - 1) = both(s[0].e[i], s[1].e[i]);
        else {
            const int j = i == 0 ? 0 : i - 2;
            hot[(kHLo + j) * 32] = both(s[0].lo[i], s[1].lo[i]);
            hot[(kHHi + j) * 32] = both(s[0].hi[i], s[1].hi[i]);
        }
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) { amax(nq_s, s[sl].q[i]); amax(nq_u, pmul(s[sl].q[i], prcp(s[sl].D[i]))); }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) hot[(kHC + i) * 32] = both(s[0].c[i], s[1].c[i]);
    hot[(kHE + 0) * 32] = both(s[0].e[0], s[1].e[0]); hot[(kHE + 1) * 32] = both(s[0].e[3], s[1].e[3]);
    hot[(kHE + 2) * 32] = both(s[0].e[4], s[1].e[4]);
    const float cs0 = s[0].cs;
    C(kCmisc) = make_float4(cm.max(nq_s), cm.max(nq_u), cs0, 1.0f / cs0);
    smem_fence();
    float rho = (float)st.rho, rdf = (float)kRhoEqOverIneq * rho;
    const float sigma = (float)st.sigma;
    const unsigned codes = rho_codes4(s);
    QuadFactor<LPS> f;
    factorize4<LPS>(cm, s, f, sigma, rho, rdf, codes, hot, cold);
    const f2 zero = bc(0.0f);
    f2 x[2][5], u[2][5], vb[2][5], rbd[2][5], rdy[2][3], vl4[2];   // z = clip(v) is recomputed where it is needed
#pragma unroll
    for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
        for (int i = 0; i < 5; ++i) { x[sl][i] = zero; u[sl][i] = s[sl].q[i]; vb[sl][i] = zero; rbd[sl][i] = zero; }
#pragma unroll
        for (int i = 0; i < 3; ++i) rdy[sl][i] = zero;
        vl4[sl] = zero;
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
        for (int i = 0; i < 5; ++i) {
            const float4 D4 = C(kCD + i);
#pragma unroll
            for (int sl = 0; sl < 2; ++sl) w[sl][i] = nan_out ? bc(NAN) : pmul(slice_of(D4, sl), x[sl][i]);
        }
        SolveResult r;
        r.iters = it;
        r.status = status;
        emit(w, r);
        done = true;
    };
    // dl, ed, eb (the certificates' operands) are NOT carried from pass to pass: 44 registers the solve needs; the pass in
    // front of a termination check leaves them in the COLD columns instead
    auto pass = [&](const bool first, const bool keep) __attribute__((always_inline)) {
        f2 td[2][3], tb[2][5], rhs[2][5], s1d[2][3], s1b[2][5], dl[2][5], ed[2][3], eb[2][5];
#pragma unroll
        for (int sl = 0; sl < 2; ++sl) {
#pragma unroll
            for (int i = 0; i < 3; ++i) td[sl][i] = pmul(rd, rdy[sl][i]);
#pragma unroll
            for (int i = 0; i < 5; ++i) tb[sl][i] = (i == 1 || i == 2) ? zero : pmul(rho_row(codes, sl, i, rho, rdf), rbd[sl][i]);
        }
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const float4 P4 = H(kHP + i);
            rhs[0][i] = pfma(slice_of(P4, 0), x[0][i], u[0][i]);
            rhs[1][i] = pfma(slice_of(P4, 1), x[1][i], u[1][i]);
        }
        At_apply4<LPS>(cm, s, td, tb, rhs, rhs, hot, zpin);  // rhs = P x + u + A'(rho r);  S D = -rhs
        kkt_solve4<LPS>(cm, f, rhs, dl, hot, zpin);
#pragma unroll
        for (int sl = 0; sl < 2; ++sl)
#pragma unroll
            for (int i = 0; i < 5; ++i) { dl[sl][i] = pmul(dl[sl][i], nal2); x[sl][i] = padd(x[sl][i], dl[sl][i]); }  // dl = alpha D
        A_apply4<LPS>(cm, s, dl, s1d, s1b, hot, zpin);
        float4 lo4[3], hi4[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            lo4[j] = lds128_after(hot + (kHLo + j) * 32, s1b[0][0].x, zpin);
            hi4[j] = lds128_after(hot + (kHHi + j) * 32, s1b[0][0].x, zpin);
        }
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
                const int j = i == 0 ? 0 : i - 2;
                const f2 rbi = rho_row(codes, sl, i, rho, rdf);
                const f2 lo_i = slice_of(lo4[j], sl), hi_i = slice_of(hi4[j], sl);
                const f2 wv = pfma(al2, rbd[sl][i], s1b[sl][i]);
                // z of the previous pass; the cold start is z = 0 whatever the bounds say (iteration 1)
                const f2 zc = pmin(pmax(vb[sl][i], lo_i), hi_i);
                const f2 zo = first ? zero : zc;
                f2 zlo = zero;
                if (MPC_COMPENSATED_V && i == 4) {
                    zlo = mk(zo.x == vb[sl][i].x ? vl4[sl].x : 0.0f, zo.y == vb[sl][i].y ? vl4[sl].y : 0.0f);
                    const f2 vs = padd(vb[sl][i], wv), bb = psub(vs, vb[sl][i]);  // TwoSum(v, w)
                    vl4[sl] = padd(vl4[sl], padd(psub(vb[sl][i], psub(vs, bb)), psub(wv, bb)));
                    vb[sl][i] = vs;
                } else {
                    vb[sl][i] = padd(vb[sl][i], wv);
                }
                const f2 zn = pmin(pmax(vb[sl][i], lo_i), hi_i);
                const f2 step = psub(zn, zo);
                if (MPC_COMPENSATED_V && i == 4) {
                    const f2 zln = mk(zn.x == vb[sl][i].x ? vl4[sl].x : 0.0f, zn.y == vb[sl][i].y ? vl4[sl].y : 0.0f);
                    const f2 stl = psub(zln, zlo);
                    rbd[sl][i] = psub(psub(padd(rbd[sl][i], s1b[sl][i]), step), stl);
                    eb[sl][i] = pmul(rbi, psub(psub(wv, step), stl));
                } else {
                    rbd[sl][i] = psub(padd(rbd[sl][i], s1b[sl][i]), step);
                    eb[sl][i] = pmul(rbi, psub(wv, step));
                }
            }
        }
        if (first) {  // iteration 1: the dynamics z jumped from the cold start 0 to d
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float4 d4 = C(kCd + i);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const f2 dd = slice_of(d4, sl);
                    rdy[sl][i] = psub(rdy[sl][i], dd);
                    ed[sl][i] = psub(ed[sl][i], pmul(rd, dd));
                }
            }
        }
        At_apply4<LPS>(cm, s, ed, eb, u, u, hot, zpin);
        if (keep) {
#pragma unroll
            for (int i = 0; i < 5; ++i) C(kCdl + i) = both(dl[0][i], dl[1][i]);
#pragma unroll
            for (int i = 0; i < 3; ++i) C(kCed + i) = both(ed[0][i], ed[1][i]);
            C(kCeb + 0) = both(eb[0][0], eb[1][0]); C(kCeb + 1) = both(eb[0][3], eb[1][3]); C(kCeb + 2) = both(eb[0][4], eb[1][4]);
        }
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
            f2 axd[2][3], axb[2][5], zb[2][5];
            A_apply4<LPS>(cm, s, x, axd, axb, hot, zpin);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const int i = j == 0 ? 0 : j + 2;
                const float4 lo4 = H(kHLo + j), hi4 = H(kHHi + j);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) zb[sl][i] = pmin(pmax(vb[sl][i], slice_of(lo4, sl)), slice_of(hi4, sl));
            }
            float pr_s = 0, pr_u = 0, nz_s = 0, nz_u = 0, nax_s = 0, nax_u = 0;
            float du_s = 0, du_u = 0, npx_s = 0, npx_u = 0, naty_s = 0, naty_u = 0;
            {
                const float4 e1 = C(kCel + 0), e2 = C(kCel + 1);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {  // loose rows: z follows A x
                    axb[sl][1] = pmul(slice_of(e1, sl), x[sl][1]); axb[sl][2] = pmul(slice_of(e2, sl), x[sl][2]);
                    zb[sl][1] = axb[sl][1]; zb[sl][2] = axb[sl][2];
                }
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) {
                const float4 d4 = C(kCd + i), E4 = C(kCEd + i);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const f2 zd = slice_of(d4, sl), Edi = prcp(slice_of(E4, sl));
                    const f2 r = psub(axd[sl][i], zd);
                    amax(pr_s, r); amax(pr_u, pmul(r, Edi));
                    amax(nz_s, zd); amax(nz_u, pmul(zd, Edi));
                    amax(nax_s, axd[sl][i]); amax(nax_u, pmul(axd[sl][i], Edi));
                }
            }
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                const float4 E4 = C(kCEb + i), D4 = C(kCD + i), P4 = H(kHP + i), q4 = C(kCq + i);
#pragma unroll
                for (int sl = 0; sl < 2; ++sl) {
                    const f2 Ebi = prcp(slice_of(E4, sl)), Di = prcp(slice_of(D4, sl));
                    const f2 r = psub(axb[sl][i], zb[sl][i]);
                    amax(pr_s, r); amax(pr_u, pmul(r, Ebi));
                    amax(nz_s, zb[sl][i]); amax(nz_u, pmul(zb[sl][i], Ebi));
                    amax(nax_s, axb[sl][i]); amax(nax_u, pmul(axb[sl][i], Ebi));
                    const f2 px = pmul(slice_of(P4, sl), x[sl][i]);
                    const f2 rr = padd(px, u[sl][i]);
                    const f2 aty = psub(u[sl][i], slice_of(q4, sl));
                    amax(du_s, rr); amax(du_u, pmul(rr, Di));
                    amax(npx_s, px); amax(npx_u, pmul(px, Di));
                    amax(naty_s, aty); amax(naty_u, pmul(aty, Di));
                }
            }
            const float4 misc = C(kCmisc);
            const float nq_s = misc.x, nq_u = misc.y, cs = misc.z, cinv = misc.w;
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
                    f2 pyb[2][5], ed[2][3], eb[2][5];
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float4 v = C(kCed + i), w = C(kCeb + i);
                        const int r = i == 0 ? 0 : i + 2;
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) { ed[sl][i] = slice_of(v, sl); eb[sl][r] = slice_of(w, sl); }
                    }
                    float ndy = 0, lhs = 0;
#pragma unroll
                    for (int i = 0; i < 3; ++i) {
                        const float4 E4 = C(kCEd + i), d4 = C(kCd + i);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            amax(ndy, pmul(slice_of(E4, sl), ed[sl][i]));
                            const f2 t = pmul(slice_of(d4, sl), ed[sl][i]);
                            lhs += t.x + t.y;
                        }
                    }
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        if (i == 1 || i == 2) { pyb[0][i] = zero; pyb[1][i] = zero; continue; }
                        const int j = i == 0 ? 0 : i - 2;
                        const float4 lo4 = H(kHLo + j), hi4 = H(kHHi + j), E4 = C(kCEb + i);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            float dv[2] = {eb[sl][i].x, eb[sl][i].y};
                            const f2 lo2 = slice_of(lo4, sl), hi2 = slice_of(hi4, sl);
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
                            amax(ndy, pmul(slice_of(E4, sl), pyb[sl][i]));
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
                        At_apply4<LPS>(cm, s, ed, pyb, z5, atdy, hot, zpin);
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const float4 D4 = C(kCD + i);
#pragma unroll
                            for (int sl = 0; sl < 2; ++sl) amax(na, pmul(atdy[sl][i], prcp(slice_of(D4, sl))));
                        }
                        na = cm.max(na);
                        pinf = cand && na < epi * ndy;
                    }
                }
                if (GC::warp_any(open && !dual_ok && !pinf)) {  // is_dual_infeasible (dx = alpha D of this iteration)
                    const float edi = tol * (float)st.eps_dual_inf;
                    f2 dl[2][5];
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 v = C(kCdl + i);
                        dl[0][i] = slice_of(v, 0); dl[1][i] = slice_of(v, 1);
                    }
                    float ndx = 0, qdx = 0, npdx = 0;
#pragma unroll
                    for (int i = 0; i < 5; ++i) {
                        const float4 D4 = C(kCD + i), q4 = C(kCq + i), P4 = H(kHP + i);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) {
                            amax(ndx, pmul(slice_of(D4, sl), dl[sl][i]));
                            const f2 t = pmul(slice_of(q4, sl), dl[sl][i]);
                            qdx += t.x + t.y;
                            amax(npdx, pmul(pmul(slice_of(P4, sl), dl[sl][i]), prcp(slice_of(D4, sl))));
                        }
                    }
                    ndx = cm.max(ndx);
                    qdx = cm.sum(qdx);
                    npdx = cm.max(npdx);
                    const bool cand = open && !dual_ok && !pinf && ndx > edi && qdx < -cs * edi * ndx && npdx < cs * edi * ndx;
                    if (GC::warp_any(cand)) {
                        f2 adxd[2][3], adxb[2][5];
                        A_apply4<LPS>(cm, s, dl, adxd, adxb, hot, zpin);
                        int bad = 0;
                        const float lim = edi * ndx;
                        const float4 e1 = C(kCel + 0), e2 = C(kCel + 1);
#pragma unroll
                        for (int sl = 0; sl < 2; ++sl) { adxb[sl][1] = pmul(slice_of(e1, sl), dl[sl][1]); adxb[sl][2] = pmul(slice_of(e2, sl), dl[sl][2]); }
#pragma unroll
                        for (int i = 0; i < 3; ++i) {
                            const float4 E4 = C(kCEd + i);
#pragma unroll
                            for (int sl = 0; sl < 2; ++sl) {
                                const f2 v = pmul(adxd[sl][i], prcp(slice_of(E4, sl)));  // equality rows have finite bounds
                                if (fabsf(v.x) > lim || fabsf(v.y) > lim) bad = 1;
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 5; ++i) {
                            const float4 E4 = C(kCEb + i);
                            const bool lz = (i == 1 || i == 2);
                            const int j = i == 0 ? 0 : i - 2;
                            const float4 lo4 = lz ? make_float4(-1e30f, -1e30f, -1e30f, -1e30f) : H(kHLo + j);
                            const float4 hi4 = lz ? make_float4(1e30f, 1e30f, 1e30f, 1e30f) : H(kHHi + j);
#pragma unroll
                            for (int sl = 0; sl < 2; ++sl) {
                                const f2 v = pmul(adxb[sl][i], prcp(slice_of(E4, sl)));
                                const f2 lo2 = slice_of(lo4, sl), hi2 = slice_of(hi4, sl);
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
                                const int j = i == 0 ? 0 : i - 2;
                                // rows with rho fixed at rho_min (loose) keep their v
                                const f2 rr = mk(((codes >> (2 * (6 * sl + j))) & 3u) == 2u ? 1.0f : ratio,
                                                 ((codes >> (2 * (6 * sl + 3 + j))) & 3u) == 2u ? 1.0f : ratio);
                                const f2 zc = zb[sl][i];  // clip(v) of this check
                                if (MPC_COMPENSATED_V && i == 4) {
                                    const f2 zl = mk(zc.x == vb[sl][i].x ? vl4[sl].x : 0.0f, zc.y == vb[sl][i].y ? vl4[sl].y : 0.0f);
                                    vb[sl][i] = pfma(padd(psub(vb[sl][i], zc), psub(vl4[sl], zl)), rr, zc);
                                    vl4[sl] = zl;
                                } else {
                                    vb[sl][i] = pfma(psub(vb[sl][i], zc), rr, zc);
                                }
                            }
                    }
                    rdf = (float)kRhoEqOverIneq * rho;
                    rd = bc(rdf);
                    factorize4<LPS>(cm, s, f, sigma, rho, rdf, codes, hot, cold);
                }
            }
        }
        return false;
    };
#ifdef MPC_QUAD_MARK   // PMTRIG markers around the pass, for counting its instructions in the SASS (tools/sass_pass.py)
#define QUAD_MARK(n) asm volatile("pmevent " #n ";")
#else
#define QUAD_MARK(n)
#endif
    for (iter = 1;; ++iter) {
        // passes after which nothing happens run in a loop of their own (see admm_pair.cuh)
        if (phase == 0 && iter > 1) {
            int quiet = st.max_iter - iter;
            if (chk > 0) quiet = min(quiet, chk - 1);
            if (adp > 0) quiet = min(quiet, adp - 1);
#pragma unroll 1
            for (int i = 0; i < quiet; ++i) pass(false, false);
            if (quiet > 0) {
                iter += quiet;
                if (chk > 0) chk -= quiet;
                if (adp > 0) adp -= quiet;
            }
        }
        QUAD_MARK(1);
        if (phase == 0) pass(iter == 1, chk == 1 || iter >= st.max_iter);
        QUAD_MARK(2);
        if (after_pass()) break;
        if (phase == 1 || (phase == 0 && iter >= st.max_iter)) {
            const bool checked = phase == 0 && st.check_termination > 0 && (st.max_iter % st.check_termination == 0);
            phase = (phase == 1 || checked) ? 2 : 1;
            if (phase == 2) tol = 10.0f;
        }
    }
}

}  // namespace mpcb
