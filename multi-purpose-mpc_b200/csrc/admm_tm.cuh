// admm_tm.cuh -- K1 + K2, paired-stage fp32 ADMM with its constants in TENSOR MEMORY: the 16-warps-per-SM variant for sm_100a.
//
// Why.  admm_pair.cuh is a latency-bound dependent chain whose 255 registers per thread cap an SM at eight warps, so 4096 cars
// (2048 warps) need two rounds.  Of those registers only ~48 are the iterate; ~120 hold CONSTANTS of the solve -- the stage
// matrices a, c, e, P, the bounds, the input-elimination and cyclic-reduction coefficients -- re-read every pass.  A B200 SM
// has a second 256 KB on-chip memory besides the register file: tensor memory (512 columns x 128 lanes x 32 bit), 12-cycle
// loads (LDTM), addressed per (lane, column) -- with tcgen05.ld/st.32x32b a warp reads and writes "column c of MY lane", i.e.
// TMEM works as a software-managed extension of the register file.  Here a CTA of four warps allocates 128 columns; warp w
// owns lanes 32 w .. 32 w + 31 of them, so every thread has 128 private 32-bit slots.  The constants live there (116 slots),
// the PCR coefficients stay in shared memory (7 KB per warp), the check-only rows move to global memory (L2), the iterate and
// the temporaries fit 128 registers: four CTAs = 16 warps per SM, 2368 warp slots for the 2048 warps of the bench -- ONE round,
// four warps per scheduler to hide the chain's latency.  Tensor cores are not involved: no MMA ever touches these columns.
// The arithmetic is admm_pair.cuh's, operation for operation (same functions, fed from TMEM): results are bit-identical.
#pragma once
#include "admm_pair.cuh"

namespace mpcb {

// tcgen05.ld / st, shape 32x32b: thread t of the warp moves N consecutive columns of lane (quarter base + t)
#define TM_R16(v, o) "r"(v[o]), "r"(v[o + 1]), "r"(v[o + 2]), "r"(v[o + 3]), "r"(v[o + 4]), "r"(v[o + 5]), "r"(v[o + 6]), "r"(v[o + 7]), \
                     "r"(v[o + 8]), "r"(v[o + 9]), "r"(v[o + 10]), "r"(v[o + 11]), "r"(v[o + 12]), "r"(v[o + 13]), "r"(v[o + 14]), "r"(v[o + 15])
#define TM_W16(v, o) "=r"(v[o]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]), \
                     "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
__device__ __forceinline__ void tm_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), TM_R16(v, 0) : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : TM_W16(v, 0) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
#undef TM_R16
#undef TM_W16

// column map of a thread's 128 slots (all groups start on a multiple of their load size)
enum : int {
    kTmA = 0,      // 16: a[0..7] (.x, .y interleaved)
    kTmCE = 16,    // 16: c[0..2], e[0], e[3], e[4] (12 used)
    kTmP = 32,     // 16: P[0..4] (10 used)
    kTmLH = 48,    // 16: lo[0,3,4], hi[0,3,4] (12 used)
    kTmEl = 64,    // 16: iv, ik, -sxv0, -sxv2, -sxk0, -sxk1, -fv, -fk
    kTmCR = 80,    // 16 + 4: DA^-1 (6), UA (6), LA (6)
    kTmEnd = 100,  // 16: last (9), D^-1 (6)
    kTmCols = 128,
};

struct TmStore {
    uint32_t base;  // TMEM address of column 0 of this warp's lane quarter
    __device__ __forceinline__ static uint32_t fu(float v) { return __float_as_uint(v); }
    __device__ __forceinline__ static float uf(uint32_t v) { return __uint_as_float(v); }
    __device__ __forceinline__ void store_stage(const Stage2& s) const {
        uint32_t v[16];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = fu(s.a[i].x); v[2 * i + 1] = fu(s.a[i].y); }
        tm_st16(base + kTmA, v);
        const f2 ce[8] = {s.c[0], s.c[1], s.c[2], s.e[0], s.e[3], s.e[4], mk(0.f, 0.f), mk(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = fu(ce[i].x); v[2 * i + 1] = fu(ce[i].y); }
        tm_st16(base + kTmCE, v);
        const f2 pp[8] = {s.P[0], s.P[1], s.P[2], s.P[3], s.P[4], mk(0.f, 0.f), mk(0.f, 0.f), mk(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = fu(pp[i].x); v[2 * i + 1] = fu(pp[i].y); }
        tm_st16(base + kTmP, v);
        const f2 lh[8] = {s.lo[0], s.lo[3], s.lo[4], s.hi[0], s.hi[3], s.hi[4], mk(0.f, 0.f), mk(0.f, 0.f)};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = fu(lh[i].x); v[2 * i + 1] = fu(lh[i].y); }
        tm_st16(base + kTmLH, v);
        tm_wait_st();
    }
    template <int LPS> __device__ __forceinline__ void store_factor(const PairFactor<LPS>& f) const {
        uint32_t v[20];
        const f2 el[8] = {f.iv, f.ik, f.nsxv0, f.nsxv2, f.nsxk0, f.nsxk1, f.nfv, f.nfk};
#pragma unroll
        for (int i = 0; i < 8; ++i) { v[2 * i] = fu(el[i].x); v[2 * i + 1] = fu(el[i].y); }
        tm_st16(base + kTmEl, v);
#pragma unroll
        for (int i = 0; i < 6; ++i) { v[i] = fu(f.DAi[i]); v[6 + i] = fu(f.UA[i]); v[12 + i] = fu(f.LA[i]); }
        v[18] = 0u; v[19] = 0u;
        tm_st16(base + kTmCR, v);
        tm_st4(base + kTmCR + 16, v + 16);
#pragma unroll
        for (int i = 0; i < 9; ++i) v[i] = fu(f.last[i]);
#pragma unroll
        for (int i = 0; i < 6; ++i) v[9 + i] = fu(f.Dinv[i]);
        v[15] = 0u;
        tm_st16(base + kTmEnd, v);
        tm_wait_st();
    }
    // a, c, e of the lane's two stages (what A x and A'y need)
    __device__ __forceinline__ void load_ace(Stage2& s) const {
        uint32_t v[32];
        tm_ld16(base + kTmA, v);
        tm_ld16(base + kTmCE, v + 16);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 8; ++i) s.a[i] = mk(uf(v[2 * i]), uf(v[2 * i + 1]));
#pragma unroll
        for (int i = 0; i < 3; ++i) s.c[i] = mk(uf(v[16 + 2 * i]), uf(v[17 + 2 * i]));
        s.e[0] = mk(uf(v[22]), uf(v[23])); s.e[3] = mk(uf(v[24]), uf(v[25])); s.e[4] = mk(uf(v[26]), uf(v[27]));
        s.e[1] = mk(0.f, 0.f); s.e[2] = mk(0.f, 0.f);
    }
    __device__ __forceinline__ void load_P(f2 (&P)[5]) const {
        uint32_t v[16];
        tm_ld16(base + kTmP, v);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 5; ++i) P[i] = mk(uf(v[2 * i]), uf(v[2 * i + 1]));
    }
    __device__ __forceinline__ void load_bounds(f2 (&lo)[5], f2 (&hi)[5]) const {
        uint32_t v[16];
        tm_ld16(base + kTmLH, v);
        tm_wait_ld();
        lo[0] = mk(uf(v[0]), uf(v[1])); lo[3] = mk(uf(v[2]), uf(v[3])); lo[4] = mk(uf(v[4]), uf(v[5]));
        hi[0] = mk(uf(v[6]), uf(v[7])); hi[3] = mk(uf(v[8]), uf(v[9])); hi[4] = mk(uf(v[10]), uf(v[11]));
        lo[1] = lo[2] = mk(-1e30f, -1e30f); hi[1] = hi[2] = mk(1e30f, 1e30f);
    }
};

// x = S^-1 b for both stages of the lane: admm_pair.cuh::kkt_solve2 with the factor read from tensor memory where it is used
template <int LPS>
__device__ __forceinline__ void kkt_solve_tm(const GroupComm<LPS>& cm, const TmStore tm, const f2 b[5], f2 x[5], const float4* cf) {
    constexpr int NLEV = PairFactor<LPS>::NLEV;
    f2 iv, ik, nsxv0, nsxv2, nsxk0, nsxk1, nfv, nfk;
    {
        uint32_t v[16];
        tm_ld16(tm.base + kTmEl, v);
        tm_wait_ld();
        iv = mk(TmStore::uf(v[0]), TmStore::uf(v[1])); ik = mk(TmStore::uf(v[2]), TmStore::uf(v[3]));
        nsxv0 = mk(TmStore::uf(v[4]), TmStore::uf(v[5])); nsxv2 = mk(TmStore::uf(v[6]), TmStore::uf(v[7]));
        nsxk0 = mk(TmStore::uf(v[8]), TmStore::uf(v[9])); nsxk1 = mk(TmStore::uf(v[10]), TmStore::uf(v[11]));
        nfv = mk(TmStore::uf(v[12]), TmStore::uf(v[13])); nfk = mk(TmStore::uf(v[14]), TmStore::uf(v[15]));
    }
    const f2 bv = pmul(iv, b[3]), bk = pmul(ik, b[4]);
    f2 bx0 = pfma(bk, nsxk0, pfma(bv, nsxv0, b[0]));
    f2 bx1 = pfma(bk, nsxk1, b[1]);
    f2 bx2 = pfma(bv, nsxv2, b[2]);
    bx1 = padd(bx1, cm.to_next(pmul(bk, nfk)));
    bx2 = padd(bx2, cm.to_next(pmul(bv, nfv)));
    float DAi[6], UA[6], LA[6];
    {
        uint32_t v[20];
        tm_ld16(tm.base + kTmCR, v);
        tm_ld4(tm.base + kTmCR + 16, v + 16);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 6; ++i) { DAi[i] = TmStore::uf(v[i]); UA[i] = TmStore::uf(v[6 + i]); LA[i] = TmStore::uf(v[12 + i]); }
    }
    // in-lane cyclic-reduction level: t = DA^-1 b_A,  b_B' = b_B - U_A' t - [Lo_A' t](l+1)
    const float bA0 = bx0.x, bA1 = bx1.x, bA2 = bx2.x;
    const float t0 = fmaf(DAi[2], bA2, fmaf(DAi[1], bA1, DAi[0] * bA0));
    const float t1 = fmaf(DAi[4], bA2, fmaf(DAi[3], bA1, DAi[1] * bA0));
    const float t2 = fmaf(DAi[5], bA2, fmaf(DAi[4], bA1, DAi[2] * bA0));
    const float h0 = fmaf(LA[2], t2, fmaf(LA[1], t1, LA[0] * t0));
    const float h1 = fmaf(LA[4], t1, LA[3] * t0);
    const float h2 = LA[5] * t2;
    f2 R0 = mk(fmaf(-UA[3], t1, fmaf(-UA[0], t0, bx0.y)), 0.0f);
    f2 R1 = mk(fmaf(-UA[4], t1, fmaf(-UA[1], t0, bx1.y)), 0.0f);
    f2 R2 = mk(fmaf(-UA[5], t2, fmaf(-UA[2], t0, bx2.y)), 0.0f);
    R0.x -= cm.next(h0); R1.x -= cm.next(h1); R2.x -= cm.next(h2);
    // PCR (coefficients of a level are read from shared memory when the level starts)
#pragma unroll
    for (int lev = 0; lev < NLEV - 1; ++lev) {
        const int sft = 1 << lev;
        const f2 n0 = mk(cm.up(R0.x, sft), cm.dn(R0.x, sft));
        const f2 n1 = mk(cm.up(R1.x, sft), cm.dn(R1.x, sft));
        const f2 n2 = mk(cm.up(R2.x, sft), cm.dn(R2.x, sft));
        f2 nab[9];
        float4 cq[6];
        const int j0 = (9 * lev) >> 1;
#pragma unroll
        for (int k = 0; k < 6; ++k)
            if (j0 + k < PcrCoef<LPS>::kF4) cq[k] = lds128v(cf + (j0 + k) * 32);
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            const int j = 9 * lev + i, c = (j >> 1) - j0;
            nab[i] = (j & 1) ? mk(cq[c].z, cq[c].w) : mk(cq[c].x, cq[c].y);
        }
        const f2 s0 = pfma(nab[2], n2, pfma(nab[1], n1, pfma(nab[0], n0, R0)));
        const f2 s1 = pfma(nab[5], n2, pfma(nab[4], n1, pfma(nab[3], n0, R1)));
        const f2 s2 = pfma(nab[8], n2, pfma(nab[7], n1, pfma(nab[6], n0, R2)));
        R0.x = s0.x + s0.y; R1.x = s1.x + s1.y; R2.x = s2.x + s2.y;
    }
    float r0 = R0.x, r1 = R1.x, r2 = R2.x;
    float xB0, xB1, xB2;
    {
        uint32_t v[16];
        tm_ld16(tm.base + kTmEnd, v);
        tm_wait_ld();
        float last[9], Dinv[6];
#pragma unroll
        for (int i = 0; i < 9; ++i) last[i] = TmStore::uf(v[i]);
#pragma unroll
        for (int i = 0; i < 6; ++i) Dinv[i] = TmStore::uf(v[9 + i]);
        const int sft = LPS / 2;
        const float n0 = cm.bfly(r0, sft), n1 = cm.bfly(r1, sft), n2 = cm.bfly(r2, sft);
        r0 = fmaf(-last[2], n2, fmaf(-last[1], n1, fmaf(-last[0], n0, r0)));
        r1 = fmaf(-last[5], n2, fmaf(-last[4], n1, fmaf(-last[3], n0, r1)));
        r2 = fmaf(-last[8], n2, fmaf(-last[7], n1, fmaf(-last[6], n0, r2)));
        xB0 = fmaf(Dinv[2], r2, fmaf(Dinv[1], r1, Dinv[0] * r0));
        xB1 = fmaf(Dinv[4], r2, fmaf(Dinv[3], r1, Dinv[1] * r0));
        xB2 = fmaf(Dinv[5], r2, fmaf(Dinv[4], r1, Dinv[2] * r0));
    }
    // back-substitution: x_A = t - DA^-1 (U_A x_B + Lo_A x_B(l-1))   (coefficients re-read)
    {
        uint32_t v[20];
        tm_ld16(tm.base + kTmCR, v);
        tm_ld4(tm.base + kTmCR + 16, v + 16);
        tm_wait_ld();
#pragma unroll
        for (int i = 0; i < 6; ++i) { DAi[i] = TmStore::uf(v[i]); UA[i] = TmStore::uf(v[6 + i]); LA[i] = TmStore::uf(v[12 + i]); }
    }
    const float p0 = cm.prev(xB0), p1 = cm.prev(xB1), p2 = cm.prev(xB2);
    const float w0 = fmaf(LA[3], p1, fmaf(LA[0], p0, fmaf(UA[2], xB2, fmaf(UA[1], xB1, UA[0] * xB0))));
    const float w1 = fmaf(LA[4], p1, fmaf(LA[1], p0, fmaf(UA[4], xB1, UA[3] * xB0)));
    const float w2 = fmaf(LA[5], p2, fmaf(LA[2], p0, UA[5] * xB2));
    const float xA0 = fmaf(-DAi[2], w2, fmaf(-DAi[1], w1, fmaf(-DAi[0], w0, t0)));
    const float xA1 = fmaf(-DAi[4], w2, fmaf(-DAi[3], w1, fmaf(-DAi[1], w0, t1)));
    const float xA2 = fmaf(-DAi[5], w2, fmaf(-DAi[4], w1, fmaf(-DAi[2], w0, t2)));
    x[0] = mk(xA0, xB0); x[1] = mk(xA1, xB1); x[2] = mk(xA2, xB2);
    const f2 xn1 = cm.from_next(x[1]), xn2 = cm.from_next(x[2]);  // fv, fk are 0 where there is no successor
    {
        uint32_t v[16];
        tm_ld16(tm.base + kTmEl, v);
        tm_wait_ld();
        iv = mk(TmStore::uf(v[0]), TmStore::uf(v[1])); ik = mk(TmStore::uf(v[2]), TmStore::uf(v[3]));
        nsxv0 = mk(TmStore::uf(v[4]), TmStore::uf(v[5])); nsxv2 = mk(TmStore::uf(v[6]), TmStore::uf(v[7]));
        nsxk0 = mk(TmStore::uf(v[8]), TmStore::uf(v[9])); nsxk1 = mk(TmStore::uf(v[10]), TmStore::uf(v[11]));
        nfv = mk(TmStore::uf(v[12]), TmStore::uf(v[13])); nfk = mk(TmStore::uf(v[14]), TmStore::uf(v[15]));
    }
    x[3] = pmul(iv, pfma(nfv, xn2, pfma(nsxv2, x[2], pfma(nsxv0, x[0], b[3]))));
    x[4] = pmul(ik, pfma(nfk, xn1, pfma(nsxk1, x[1], pfma(nsxk0, x[0], b[4]))));
}

// The OSQP loop: admm_pair.cuh::admm_solve2 for the reference's unbounded e_psi / t rows, with `sm` (the check-only rows) in
// GLOBAL memory and every constant of the pass fetched from tensor memory where it is used.
template <int LPS, typename Emit>
__device__ __forceinline__ void admm_solve_tm(const GroupComm<LPS>& cm, Stage2& s, const AdmmSettings& st, const f2 al2,
                                              const f2 nal2, int nvar, f2* sm, float4* cf, const TmStore tm, bool live, Emit emit) {
    constexpr bool LOOSE = true;
    typedef GroupComm<LPS> GC;
    const int gl = cm.gl;
    if (st.scaling > 0) ruiz_scale2<LPS>(cm, s, st.scaling, nvar);
    else {
#pragma unroll
        for (int i = 0; i < 5; ++i) { s.D[i] = bc(1.0f); s.Eb[i] = bc(1.0f); }
#pragma unroll
        for (int i = 0; i < 3; ++i) s.Ed[i] = bc(1.0f);
        s.cs = 1.0f;
    }
    const float thr = (float)(kOsqpInfty * kMinScaling);
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
    {
        PairFactor<LPS> f;
        factorize2<LPS, LOOSE>(cm, s, f, sigma, rdf, rb, sm, cf);
        tm.store_stage(s);
        tm.store_factor(f);
    }
    float nq_s = 0.0f, nq_u = 0.0f;
#pragma unroll
    for (int i = 0; i < 5; ++i) { amax(nq_s, s.q[i]); amax(nq_u, pmul(s.q[i], sm[(16 + i) * LPS + gl])); }
    sm[54 * LPS + gl] = mk(cm.max(nq_s), cm.max(nq_u));
    sm[55 * LPS + gl] = mk(s.cs, 1.0f / s.cs);
    const f2 zero = bc(0.0f);
    static_assert(!MPC_Y_FORM, "the tensor-memory variant implements the u-form only");
    f2 x[5], u[5], yd[3], vb[5], rbd[5], rdy[3];   // z = clip(v) is not carried: recomputed where it is needed (8 registers)
    f2 vl4 = zero;  // low word of v of the curvature row (MPC_COMPENSATED_V)
#pragma unroll
    for (int i = 0; i < 5; ++i) { x[i] = zero; u[i] = s.q[i]; vb[i] = zero; rbd[i] = zero; }
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
        emit(w, r);
        done = true;
    };
    // one ADMM pass
    auto pass = [&](const bool first, const bool keep) __attribute__((always_inline)) {
        f2 td[3], tb[5], rhs[5], s1d[3], s1b[5], dl[5], ed[3], eb[5], Pv[5];
        f2 zb[5] = {zero, zero, zero, zero, zero}, zl4 = zero;  // (names the dead y-form branches refer to)
        tm.load_P(Pv);
#pragma unroll
        for (int i = 0; i < 3; ++i) td[i] = MPC_Y_FORM ? pfma(rd, rdy[i], yd[i]) : pmul(rd, rdy[i]);
#pragma unroll
        for (int i = 0; i < 5; ++i) {
            const f2 lin = MPC_Y_FORM ? (MPC_PASS_Q_REG ? s.q[i] : ldsv(&sm[(29 + i) * LPS + gl])) : u[i];
            rhs[i] = pfma(Pv[i], x[i], lin);
            if (LOOSE && (i == 1 || i == 2)) continue;
            if (MPC_Y_FORM) {  // y + rho r = rho ((v - z) + r)
                f2 vz = psub(vb[i], zb[i]);
                if (MPC_COMPENSATED_V && i == 4) vz = padd(vz, psub(vl4, zl4));
                tb[i] = pmul(rb[i], padd(vz, rbd[i]));
            } else {
                tb[i] = pmul(rb[i], rbd[i]);
            }
        }
        {
            Stage2 sa;
            tm.load_ace(sa);
            At_apply2<LPS, LOOSE>(cm, sa, td, tb, rhs, rhs);  // rhs = P x + q + A'(y + rho r);  S D = -rhs
        }
        kkt_solve_tm<LPS>(cm, tm, rhs, dl, cf);
#pragma unroll
        for (int i = 0; i < 5; ++i) { dl[i] = pmul(dl[i], nal2); x[i] = padd(x[i], dl[i]); }  // dl = alpha D
        {
            Stage2 sa;
            tm.load_ace(sa);
            A_apply2<LPS, LOOSE>(cm, sa, dl, s1d, s1b);
        }
        f2 lo3[5], hi3[5];
        tm.load_bounds(lo3, hi3);
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
            // z of the previous pass = clip(v); the cold start is z = 0 whatever the bounds say (iteration 1)
            const f2 zc = pmin(pmax(vb[i], lo3[i]), hi3[i]);
            const f2 zo = first ? zero : zc;
            f2 zlo = zero;
            if (MPC_COMPENSATED_V && i == 4) {
                zlo = mk(zo.x == vb[i].x ? vl4.x : 0.0f, zo.y == vb[i].y ? vl4.y : 0.0f);
                const f2 vs = padd(vb[i], wv), bb = psub(vs, vb[i]);             // TwoSum(v, w)
                vl4 = padd(vl4, padd(psub(vb[i], psub(vs, bb)), psub(wv, bb)));
                vb[i] = vs;
            } else {
                vb[i] = padd(vb[i], wv);
            }
            const f2 zn = pmin(pmax(vb[i], lo3[i]), hi3[i]);
            const f2 step = psub(zn, zo);
            if (MPC_COMPENSATED_V && i == 4) {
                const f2 zln = mk(zn.x == vb[i].x ? vl4.x : 0.0f, zn.y == vb[i].y ? vl4.y : 0.0f);  // inactive row: z = v
                const f2 stl = psub(zln, zlo);
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
            Stage2 sa;
            tm.load_ace(sa);
            At_apply2<LPS, LOOSE>(cm, sa, ed, eb, u, u);
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
            Stage2 sa;
            tm.load_ace(sa);
            A_apply2<LPS, LOOSE>(cm, sa, x, axd, axb);
            f2 zb[5], zl4 = zero;
#pragma unroll
            for (int i = 0; i < 5; ++i) {
                if (i == 1 || i == 2) continue;
                zb[i] = pmin(pmax(vb[i], ldsv(&sm[(39 + i) * LPS + gl])), ldsv(&sm[(44 + i) * LPS + gl]));
                if (MPC_COMPENSATED_V && i == 4) zl4 = mk(zb[i].x == vb[i].x ? vl4.x : 0.0f, zb[i].y == vb[i].y ? vl4.y : 0.0f);
            }
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
                At_apply2<LPS, LOOSE>(cm, sa, yd, yb5, z5, aty5);
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
                        At_apply2<LPS, LOOSE>(cm, sa, ed, pyb, z5, atdy);
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
                        A_apply2<LPS, LOOSE>(cm, sa, dl, adxd, adxb);
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
                    PairFactor<LPS> f;
                    factorize2<LPS, LOOSE>(cm, sa, f, sigma, rdf, rb, sm, cf);
                    tm.store_factor(f);
                }
            }
        }
        return false;
    };
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
#ifdef MPC_QUAD_MARK
        asm volatile("pmevent 1;");
#endif
        if (phase == 0) pass(iter == 1, chk == 1 || iter >= st.max_iter);
#ifdef MPC_QUAD_MARK
        asm volatile("pmevent 2;");
#endif
        if (after_pass()) break;   // phase 2 always ends here: every scenario still open is finished with -2
        if (phase == 1 || (phase == 0 && iter >= st.max_iter)) {
            // the last pass was a check pass iff check_termination divides max_iter
            const bool checked = phase == 0 && st.check_termination > 0 && (st.max_iter % st.check_termination == 0);
            phase = (phase == 1 || checked) ? 2 : 1;
            if (phase == 2) tol = 10.0f;
        }
    }
}

}  // namespace mpcb
