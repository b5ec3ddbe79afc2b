// geometry.cu -- K3 (dynamic raycast), K3b (static width), obstacle rasteriser, K4 (localise + t2s,
// rollout) for sm_100a.  Compiled with -fmad=false: every fp64 expression here must round exactly
// like the reference's numpy / CPython arithmetic (no FMA contraction), because grid cells and
// drivable widths are compared bit-for-bit.
//
// Reference: src/reference_path.py:206-287 (static width), 466-648 (update_path_constraints),
// src/map.py:77-137 (w2m, m2w, add_obstacles), src/spatial_bicycle_models.py:183-279 (t2s, drive,
// get_current_waypoint); skimage.draw.line_aa cell order restated from skimage/draw/_draw.pyx.
#include "launch_util.h"
#include "engine.h"
#include <cstdlib>

namespace mpcb {

__device__ __forceinline__ double np_mod(double a, double b) {  // numpy floor-mod, b > 0
    double m = fmod(a, b);
    if (m != 0.0) { if (m < 0) m += b; } else m = 0.0;
    return m;
}
__device__ __forceinline__ double sq(double x) { return x * x; }

__device__ __forceinline__ void w2m(const GridView& g, double x, double y, int& dx, int& dy) {
    dx = (int)floor((x - g.ox) / g.res);  // map.py:85  (IEEE divide, not reciprocal-multiply)
    dy = (int)floor((y - g.oy) / g.res);  // map.py:86
}
__device__ __forceinline__ void m2w(const GridView& g, int dx, int dy, double& x, double& y) {
    x = ((double)dx + 0.5) * g.res + g.ox;  // map.py:98
    y = ((double)dy + 0.5) * g.res + g.oy;  // map.py:99
}

// skimage.draw.line_aa(r0, c0, r1, c1) cell sequence; the reference passes (x, y) as (r, c).
// Usage:  LineAA it(x0,y0,x1,y1); do { visit(it.x, it.y) for each emitted cell } -- see walk().
template <typename Visit> __device__ __forceinline__ void line_aa_walk(int r0, int c0, int r1, int c1, Visit&& visit) {
    const int dc = abs(c0 - c1), dr = abs(r0 - r1);
    float err = (float)(dc - dr);
    const int sign_c = (c0 < c1) ? 1 : -1, sign_r = (r0 < r1) ? 1 : -1;
    const float ed = (dc + dr == 0) ? 1.0f : (float)sqrt((double)(dc * dc + dr * dr));
    int c = c0, r = r0;
    for (;;) {
        if (!visit(r, c)) return;
        const float err_prime = err;
        const int c_prime = c;
        if (2 * err_prime >= -dc) {
            if (c == c1) break;
            if (err_prime + dr < ed) { if (!visit(r + sign_r, c)) return; }
            err -= dr;
            c += sign_c;
        }
        if (2 * err_prime <= dr) {
            if (r == r1) break;
            if (dc - err_prime < ed) { if (!visit(r, c_prime + sign_c)) return; }
            err += dc;
            r += sign_r;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// obstacle rasteriser: Map.add_obstacles (map.py:126-137) into per-scenario bit-packed grids
// ------------------------------------------------------------------------------------------------
// one CTA per scenario; grid words staged in shared memory, discs applied with atomicAnd.
__global__ void rasterize_kernel(const uint32_t* __restrict__ base, uint32_t* __restrict__ grids, int words, GridView g,
                                 const int* __restrict__ obs_px /*[n][3] cx,cy,r*/, const int* __restrict__ offsets) {
    extern __shared__ uint32_t sm[];
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < words; i += blockDim.x) sm[i] = base[i];
    __syncthreads();
    for (int o = offsets[b]; o < offsets[b + 1]; ++o) {
        const int cx = obs_px[3 * o], cy = obs_px[3 * o + 1], r = obs_px[3 * o + 2];
        const int side = 2 * r;  // window [-r, r) (np.ogrid[-r:r], quirk Q5)
        for (int t = threadIdx.x; t < side * side; t += blockDim.x) {
            const int dy = t / side - r, dx = t % side - r;
            if (dx * dx + dy * dy <= r * r) {
                const int x = cx + dx, y = cy + dy;
                if (x >= 0 && y >= 0 && x < g.W && y < g.H) atomicAnd(&sm[y * g.pitch_words + (x >> 5)], ~(1u << (x & 31)));
            }
        }
    }
    __syncthreads();
    uint32_t* out = grids + (size_t)b * words;
    for (int i = threadIdx.x; i < words; i += blockDim.x) out[i] = sm[i];
}

void launch_rasterize(const uint32_t* base, uint32_t* grids, int words, const GridView& g, const int* obs_px,
                      const int* offsets, int B, cudaStream_t st) {
    NvtxRange nvtx_("mpc:rasterize");
    { static int have_ = 0; ensure_dynamic_smem(rasterize_kernel, have_, (size_t)words * 4); }
    rasterize_kernel<<<B, 256, words * 4, st>>>(base, grids, words, g, obs_px, offsets);
}

// ------------------------------------------------------------------------------------------------
// K3b: ReferencePath._compute_width / _get_min_width (rp.py:206-287)
// one warp per (waypoint, side); lanes 0..8 walk the 9 anti-aliased rays (target cell +-1).
// ------------------------------------------------------------------------------------------------
// Batched over tracks (blockIdx.y): track t has its own bit grid (grid_stride_words apart) and its own waypoint table
// (double[12][n_max] rows as mpc_set_path takes them, table_stride doubles apart) with n_wp_arr[t] valid waypoints; the
// single-track call is T = 1 with the engine's own tables.
__global__ void compute_width_kernel(const uint32_t* __restrict__ grids, size_t grid_stride_words, GridView g,
                                     const double* __restrict__ tables, size_t table_stride, int n_max,
                                     const int* __restrict__ n_wp_arr, int n_wp_single, double max_width,
                                     double* __restrict__ out_ub, double* __restrict__ out_lb,
                                     double* __restrict__ out_border, int* __restrict__ err_flag) {
    const int t = blockIdx.y;
    const int n_wp = n_wp_arr ? n_wp_arr[t] : n_wp_single;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= 2 * n_wp) return;
    const uint32_t* grid = grids + (size_t)t * grid_stride_words;
    const double* tab = tables + (size_t)t * table_stride;
    out_ub += (size_t)t * n_max; out_lb += (size_t)t * n_max; out_border += (size_t)t * n_max * 4;
    const int k = warp >> 1, side = warp & 1;
    const double wx = tab[k], wy = tab[(size_t)n_max + k];
    // angle = mod(psi +- pi/2 + pi, 2pi) - pi (rp.py:221-225): rows 8 / 9 (upper) and 10 / 11 (lower) of the table
    const double ca = tab[(size_t)(side == 0 ? 8 : 10) * n_max + k];
    const double sa = tab[(size_t)(side == 0 ? 9 : 11) * n_max + k];
    int tx, ty, px, py;
    w2m(g, wx + max_width * ca, wy + max_width * sa, tx, ty);  // rp.py:227
    w2m(g, wx, wy, px, py);                                    // rp.py:263
    double best = max_width;
    int bx = 0, by = 0, found = 0, bad = 0;
    if (lane < 9) {
        const int i = lane / 3 - 1, j = lane % 3 - 1;  // tn_x = t_x + i (outer), tn_y = t_y + j (inner), rp.py:257-260
        line_aa_walk(px, py, tx + i, ty + j, [&](int x, int y) {
            int xx = x < 0 ? x + g.W : x, yy = y < 0 ? y + g.H : y;  // numpy negative index wrap
            if (xx < 0 || yy < 0 || xx >= g.W || yy >= g.H) { bad = 1; return false; }
            const uint32_t wd = grid[yy * g.pitch_words + (xx >> 5)];
            if (!((wd >> (xx & 31)) & 1u)) {  // occupied (rp.py:279)
                double cx, cy;
                m2w(g, x, y, cx, cy);
                const double d = sqrt(sq(wx - cx) + sq(wy - cy));  // rp.py:282
                if (d < best) { best = d; bx = x; by = y; found = 1; }
            }
            return true;
        });
    }
    if (__any_sync(0xffffffffu, bad)) { if (lane == 0) atomicOr(&err_flag[t], MPC_ST_INDEX_ERROR); }
    // sequential `<` over paths in order: the smallest distance wins, ties go to the earliest path
    double wbest = best;
    int wl = found ? lane : 64;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ob = __shfl_xor_sync(0xffffffffu, wbest, s);
        const int ol = __shfl_xor_sync(0xffffffffu, wl, s);
        if (ob < wbest || (ob == wbest && ol < wl)) { wbest = ob; wl = ol; }
    }
    const int src = wl < 32 ? wl : 0;
    bx = __shfl_sync(0xffffffffu, bx, src);
    by = __shfl_sync(0xffffffffu, by, src);
    if (lane == 0) {
        double cx, cy;
        if (wl < 32) m2w(g, bx, by, cx, cy);
        else m2w(g, tx + 1, ty + 1, cx, cy);  // rp.py:274 with the loop-leaked (t_x+1, t_y+1) (quirk Q3)
        if (side == 0) out_ub[k] = wbest; else out_lb[k] = -1 * wbest;  // rp.py:236-237
        out_border[4 * k + 2 * side] = cx;
        out_border[4 * k + 2 * side + 1] = cy;
    }
}

void launch_compute_width(const uint32_t* grid, const GridView& g, const PathView& pv, double max_width, double* ub,
                          double* lb, double* border, int* err, cudaStream_t st) {
    NvtxRange nvtx_("mpc:K3b compute_width");
    const int warps = 2 * pv.n_wp;
    // the engine's path rows are contiguous: double[13][n_wp] starting at pv.x (engine.cu::bind_path)
    compute_width_kernel<<<dim3((warps * 32 + 127) / 128, 1), 128, 0, st>>>(grid, 0, g, pv.x, 0, pv.n_wp, nullptr, pv.n_wp, max_width,
                                                                         ub, lb, border, err);
}

void launch_compute_width_batch(const uint32_t* grids, size_t grid_stride_words, const GridView& g, const double* tables,
                                int n_max, const int* n_wp, int T, double max_width, double* ub, double* lb, double* border,
                                int* err, cudaStream_t st) {
    NvtxRange nvtx_("mpc:K3b compute_width (batched over tracks)");
    const int warps = 2 * n_max;
    compute_width_kernel<<<dim3((warps * 32 + 127) / 128, T), 128, 0, st>>>(grids, grid_stride_words, g, tables, (size_t)12 * n_max,
                                                                         n_max, n_wp, 0, max_width, ub, lb, border, err);
}

// ------------------------------------------------------------------------------------------------
// K4 front: SpatialBicycleModel.get_current_waypoint + t2s (sbm.py:256-279, 183-219)
// ------------------------------------------------------------------------------------------------
// returns the waypoint index, or -1 when s is past the end of the path (simulation.py:134 loop condition)
__device__ __forceinline__ int localize_one(const double* __restrict__ state, double* __restrict__ spatial,
                                            const PathView& pv, double length, int b, int B) {
    const double x = state[b], y = state[(size_t)B + b], psi = state[2 * (size_t)B + b], s = state[3 * (size_t)B + b];
    // first index with length_cum > s  (sbm.py:265-266); all-False -> IndexError in the reference
    int lo = 0, hi = pv.n_wp;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (pv.length_cum[mid] > s) hi = mid; else lo = mid + 1;
    }
    if (lo >= pv.n_wp || !(s < length)) return -1;
    const int next = lo, prev = next > 0 ? next - 1 : pv.n_wp - 1;  // index -1 wraps in numpy
    const double s_next = pv.length_cum[next], s_prev = pv.length_cum[prev];
    const int w = (fabs(s - s_next) < fabs(s - s_prev)) ? next : prev;  // strict <: ties -> prev (quirk Q8)
    const double e_y = pv.cos_psi[w] * (y - pv.y[w]) - pv.sin_psi[w] * (x - pv.x[w]);  // sbm.py:202-205
    const double PI = 3.141592653589793;
    double e_psi = psi - pv.psi[w];
    e_psi = np_mod(e_psi + PI, 2 * PI) - PI;  // sbm.py:209
    spatial[b] = e_y;
    spatial[(size_t)B + b] = e_psi;
    return w;
}

// The same for a whole warp working on one car: the search over the cumulative lengths is spread over the lanes
// (independent loads instead of a chain of dependent ones; length_cum is non-decreasing, so the first index with
// length_cum > s is the number of entries <= s).  Returns the waypoint in every lane; lane 0 writes `spatial`.
__device__ __forceinline__ int localize_warp_values(double x, double y, double psi, double s, double* __restrict__ spatial,
                                                    const PathView& pv, double length, int b, int B, int lane) {
    int cnt = 0;
    for (int i = lane; i < pv.n_wp; i += 32) cnt += (pv.length_cum[i] > s) ? 0 : 1;
    const int lo = __reduce_add_sync(0xffffffffu, cnt);
    if (lo >= pv.n_wp || !(s < length)) return -1;
    const int next = lo, prev = next > 0 ? next - 1 : pv.n_wp - 1;  // index -1 wraps in numpy
    int w = 0;
    if (lane == 0) {
        const double s_next = pv.length_cum[next], s_prev = pv.length_cum[prev];
        w = (fabs(s - s_next) < fabs(s - s_prev)) ? next : prev;  // strict <: ties -> prev (quirk Q8)
        const double e_y = pv.cos_psi[w] * (y - pv.y[w]) - pv.sin_psi[w] * (x - pv.x[w]);  // sbm.py:202-205
        const double PI = 3.141592653589793;
        double e_psi = psi - pv.psi[w];
        e_psi = np_mod(e_psi + PI, 2 * PI) - PI;  // sbm.py:209
        spatial[b] = e_y;
        spatial[(size_t)B + b] = e_psi;
    }
    return __shfl_sync(0xffffffffu, w, 0);
}
__device__ __forceinline__ int localize_warp(const double* __restrict__ state, double* __restrict__ spatial,
                                             const PathView& pv, double length, int b, int B, int lane) {
    const double s = state[3 * (size_t)B + b];
    double x = 0.0, y = 0.0, psi = 0.0;  // used by lane 0 only
    if (lane == 0) { x = state[b]; y = state[(size_t)B + b]; psi = state[2 * (size_t)B + b]; }
    return localize_warp_values(x, y, psi, s, spatial, pv, length, b, B, lane);
}

__global__ void localize_t2s_kernel(const double* __restrict__ state, int* __restrict__ wp_id,
                                    double* __restrict__ spatial, int* __restrict__ flags, PathView pv, double length,
                                    int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (flags && (flags[b] & (MPC_ST_DEAD | MPC_ST_FINISHED))) return;
    const int w = localize_one(state, spatial, pv, length, b, B);
    if (w < 0) {
        if (flags) atomicOr(&flags[b], MPC_ST_FINISHED);
        return;
    }
    wp_id[b] = w;
}

// ------------------------------------------------------------------------------------------------
// K3: ReferencePath.update_path_constraints (rp.py:522-648) on per-scenario bit-packed grids.
// One warp per scenario.  The rows of the grid that the horizon's rays can touch are one
// contiguous byte range (full 64 B-pitch rows): a single cp.async.bulk (TMA bulk copy, UBLKCP)
// stages it in shared memory, signalled through an mbarrier.  Phase 1: lane n replays the
// precomputed cell sequence of horizon waypoint n's ray (ray table) against the scenario's grid and
// records its free segments (rp.py:466-520); phase 2: every waypoint with <= 1 candidate (or n == 0)
// is finalised independently; phase 3: the few waypoints with >= 2 candidates are resolved in
// order (nearest to the projected previous pick, rp.py:552-586).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxSeg = 8;

struct RayOut {
    double ub, lb, cells_sm[4], cells[4];
};

__device__ __forceinline__ RayOut finalize_wp(const PathView& pv, int k, double ubx, double uby, double lbx, double lby,
                                              double sm) {
    RayOut o;
    const double wx = pv.x[k], wy = pv.y[k], psi = pv.psi[k];
    const double PI = 3.141592653589793;
    const double angle_ub = np_mod(atan2(uby - wy, ubx - wx) - psi + PI, 2 * PI) - PI;  // rp.py:598
    const double angle_lb = np_mod(atan2(lby - wy, lbx - wx) - psi + PI, 2 * PI) - PI;  // rp.py:600
    const double sgu = (double)((angle_ub > 0) - (angle_ub < 0)), sgl = (double)((angle_lb > 0) - (angle_lb < 0));
    double ub = sgu * sqrt(sq(ubx - wx) + sq(uby - wy));  // rp.py:606
    double lb = sgl * sqrt(sq(lbx - wx) + sq(lby - wy));  // rp.py:608
    ub -= sm;
    lb += sm;
    if (ub < lb) { ub = 0.0; lb = 0.0; }  // rp.py:616-619
    const double cu = pv.cos_ub[k], su = pv.sin_ub[k], cl = pv.cos_lb[k], sl = pv.sin_lb[k];
    o.cells_sm[0] = wx + ub * cu; o.cells_sm[1] = wy + ub * su;  // rp.py:627
    o.cells_sm[2] = wx - lb * cl; o.cells_sm[3] = wy - lb * sl;  // rp.py:629
    o.cells[0] = wx + (ub + sm) * cu; o.cells[1] = wy + (ub + sm) * su;  // rp.py:633
    o.cells[2] = wx - (lb - sm) * cl; o.cells[3] = wy - (lb - sm) * sl;  // rp.py:635
    o.ub = ub;
    o.lb = lb;
    return o;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// TMA bulk copy global -> shared (cp.async.bulk, SASS UBLKCP) signalled through an mbarrier.
__device__ __forceinline__ void tma_bulk_load(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(mbar))
                 : "memory");
}
__device__ __forceinline__ void mbar_init(uint64_t* mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* mbar, uint32_t parity) {
    uint32_t done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(mbar)), "r"(parity)
            : "memory");
    }
}

// closing a free segment (rp.py:503-515) is rare: keep its fp64 arithmetic out of line so that the replay loop
// stays small
__device__ __noinline__ int close_segment(double ox, double oy, double res, int uo_x, int uo_y, int x, int y,
                                          double min_width, short4* segs, int nseg) {
    const double ux = ((double)uo_x + 0.5) * res + ox, uy = ((double)uo_y + 0.5) * res + oy;  // map.py:98-99
    const double lx = ((double)x + 0.5) * res + ox, ly = ((double)y + 0.5) * res + oy;
    if (sqrt(sq(ux - lx) + sq(uy - ly)) > min_width) {  // rp.py:510
        if (nseg < kMaxSeg) segs[nseg] = make_short4((short)uo_x, (short)uo_y, (short)x, (short)y);
        ++nseg;
    }
    return nseg;
}

// ------------------------------------------------------------------------------------------------
// Ray table: the cell sequence of every waypoint's ray (static border cell -> static border cell) is a
// property of the path and the grid geometry, not of the scenario.  It is built once (per set_path /
// set_base_grid / compute_width) by walking skimage's line_aa order on the device, and the per-step kernel
// replays it: entry = word offset in the grid (bits 6..31) | at-end flag (bit 5) | bit (0..4), layout
// [cell][waypoint] so that the lanes of a warp (consecutive waypoints) read consecutive words.  Entries past a ray's
// end repeat its last cell: replaying them is inert (an occupied cell cannot close a segment when no free cell is
// pending, a free end cell closes a zero-length one that rp.py:510 discards), so the replay needs no per-cell
// "is this lane still inside its ray" predicate.
// ------------------------------------------------------------------------------------------------
constexpr uint32_t kRayEnd = 1u << 5;

__global__ void build_ray_table_kernel(GridView g, PathView pv, uint32_t* __restrict__ cells, int* __restrict__ len,
                                       int max_len) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= pv.n_wp) return;
    const double* bc = pv.border + 4 * k;
    int ubx, uby, lbx, lby;
    w2m(g, bc[0], bc[1], ubx, uby);  // rp.py:478
    w2m(g, bc[2], bc[3], lbx, lby);  // rp.py:480
    int n = 0, bad = 0, first = 1;
    uint32_t last = 0u;
    line_aa_walk(ubx, uby, lbx, lby, [&](int x, int y) {
        if (first) { first = 0; return true; }  // rp.py:494 skips the first emitted cell (quirk Q4)
        if (x < 0 || y < 0 || x >= g.W || y >= g.H || n >= max_len) { bad = 1; return false; }
        const uint32_t word = (uint32_t)(y * g.pitch_words + (x >> 5));
        last = (word << 6) | (uint32_t)(x & 31) | ((x == lbx && y == lby) ? kRayEnd : 0u);
        cells[(size_t)n * pv.n_wp + k] = last;
        ++n;
        return true;
    });
    for (int i = n; i < max_len; ++i) cells[(size_t)i * pv.n_wp + k] = last;
    len[k] = bad ? -1 : n;  // -1: the ray leaves the grid (IndexError in the reference, rp.py:496)
}

void launch_build_ray_table(const GridView& g, const PathView& pv, uint32_t* cells, int* len, int max_len,
                            cudaStream_t st) {
    NvtxRange nvtx_("mpc:build_ray_table");
    build_ray_table_kernel<<<(pv.n_wp + 127) / 128, 128, 0, st>>>(g, pv, cells, len, max_len);
}

// Replay one waypoint's ray over a bit grid and record its free segments (rp.py:494-518).
// base[word] is the grid word (staged rows: base is pre-offset by the first staged row).
// Per cell: v = its bit; term = occupied or at-end; a segment closes at a term cell when a free cell is pending
// (rp.py:503-515); `uo` (the segment's upper end) moves to every term cell (rp.py:514 / 516-518).
__device__ __forceinline__ int replay_ray(const uint32_t* __restrict__ base, const uint32_t* __restrict__ cells, int n_wp,
                                          int k, int max_len_warp, int ubx, int uby, int pitch, double ox, double oy,
                                          double res, double min_width, short4* segs) {
    uint32_t uo = 0xffffffffu;  // packed cell of the segment's upper end; all ones = the ray's start cell
    uint32_t pending = 0u;      // bit 0: a free cell has been seen since the last close
    int nseg = 0;
    const uint32_t* p = cells + k;
    constexpr int U = 8;
    // all lanes run the warp's longest ray so that the loads stay converged (shorter rays replay their padding).
    // Software pipeline: the table entries of batch i+1 are in flight (L2 latency) while batch i's grid words are
    // fetched from shared memory and its (serial) segment state machine runs.
    uint32_t en[U];
#pragma unroll
    for (int j = 0; j < U; ++j) en[j] = __ldg(p + (size_t)j * n_wp);
    for (int i0 = 0; i0 < max_len_warp; i0 += U) {
        uint32_t e[U], wv[U];
#pragma unroll
        for (int j = 0; j < U; ++j) e[j] = en[j];
        p += (size_t)U * n_wp;
        if (i0 + U < max_len_warp) {
#pragma unroll
            for (int j = 0; j < U; ++j) en[j] = __ldg(p + (size_t)j * n_wp);
        }
#pragma unroll
        for (int j = 0; j < U; ++j) wv[j] = base[e[j] >> 6];
#pragma unroll
        for (int j = 0; j < U; ++j) {
            const uint32_t x = __funnelshift_r(wv[j], 0u, e[j]);  // bit 0 = the cell's bit (shift count = e & 31)
            const uint32_t term = (~x | (e[j] >> 5)) & 1u;         // occupied, or the ray's end cell
            pending |= x;
            if (term & pending) {  // rp.py:503-515
                const uint32_t w = e[j] >> 6;
                const int y = (int)(w / (uint32_t)pitch), xx = (int)(w % (uint32_t)pitch) * 32 + (int)(e[j] & 31u);
                int ux = ubx, uy = uby;
                if (uo != 0xffffffffu) {
                    const uint32_t wu = uo >> 6;
                    uy = (int)(wu / (uint32_t)pitch); ux = (int)(wu % (uint32_t)pitch) * 32 + (int)(uo & 31u);
                }
                nseg = close_segment(ox, oy, res, ux, uy, xx, y, min_width, segs, nseg);
                pending = 0u;
            }
            uo = term ? e[j] : uo;
        }
    }
    return nseg;
}

// The slow exact path of a ray with more free segments than its scratch holds (kMaxSeg; the reference's list is unbounded,
// rp.py:466-520): ONE lane walks ray k again and stores the segments number first .. first + kMaxSeg - 1.
__device__ __noinline__ void refill_segments(const uint32_t* __restrict__ base, const uint32_t* __restrict__ cells, int n_wp,
                                             int k, int len, int ubx, int uby, int pitch, double ox, double oy, double res,
                                             double min_width, short4* segs, int first) {
    uint32_t uo = 0xffffffffu, pending = 0u;
    int nseg = 0;
    for (int i = 0; i < len; ++i) {
        const uint32_t e = __ldg(cells + (size_t)i * n_wp + k);
        const uint32_t x = (base[e >> 6] >> (e & 31u)) & 1u;
        const uint32_t term = (~x | (e >> 5)) & 1u;
        pending |= x;
        if (term & pending) {
            const uint32_t w = e >> 6;
            const int y = (int)(w / (uint32_t)pitch), xx = (int)(w % (uint32_t)pitch) * 32 + (int)(e & 31u);
            int ux = ubx, uy = uby;
            if (uo != 0xffffffffu) {
                const uint32_t wu = uo >> 6;
                uy = (int)(wu / (uint32_t)pitch); ux = (int)(wu % (uint32_t)pitch) * 32 + (int)(uo & 31u);
            }
            const double uxw = ((double)ux + 0.5) * res + ox, uyw = ((double)uy + 0.5) * res + oy;
            const double lxw = ((double)xx + 0.5) * res + ox, lyw = ((double)y + 0.5) * res + oy;
            if (sqrt(sq(uxw - lxw) + sq(uyw - lyw)) > min_width) {
                if (nseg >= first && nseg < first + kMaxSeg) segs[nseg - first] = make_short4((short)ux, (short)uy, (short)xx, (short)y);
                ++nseg;
            }
            pending = 0u;
        }
        uo = term ? e : uo;
    }
}

// The two out-of-line routines below take scalars only (a reference to the kernel's GridView / PathView would force a copy of
// the structs into local memory on every path).
struct WinGeom { int pitch, n_wp; double ox, oy, res, min_width; };
__device__ __forceinline__ void m2w_s(const WinGeom& w, int dx, int dy, double& x, double& y) {
    x = ((double)dx + 0.5) * w.res + w.ox;  // map.py:98
    y = ((double)dy + 0.5) * w.res + w.oy;  // map.py:99
}
// rp.py:552-586 for a ray with more than kMaxSeg free segments: the whole warp takes the candidates kMaxSeg at a time (lane 0
// re-walks the ray for every window after the first) and keeps the first minimum of the mean end-point distance.
__device__ __noinline__ void pick_nearest_windowed(int pitch, int n_wp, double ox, double oy, double res, double min_width,
                                                   const uint32_t* base, const uint32_t* cells, int len, int k, int nseg, int sx,
                                                   int sy, short4* segs, double upx, double upy, double lpx, double lpy, int lane,
                                                   double* out4) {
    const WinGeom w{pitch, n_wp, ox, oy, res, min_width};
    double best_md = INFINITY, bux = 0, buy = 0, blx = 0, bly = 0;
    for (int w0 = 0; w0 < nseg; w0 += kMaxSeg) {
        if (w0 > 0) {
            if (lane == 0) refill_segments(base, cells, n_wp, k, len, sx, sy, pitch, ox, oy, res, min_width, segs, w0);
            __syncwarp();
        }
        const int cnt = nseg - w0 < kMaxSeg ? nseg - w0 : kMaxSeg;
        double md = INFINITY, ubx = 0, uby = 0, lbx = 0, lby = 0;
        if (lane < cnt) {
            const short4 s4 = segs[lane];
            m2w_s(w, s4.x, s4.y, ubx, uby);
            m2w_s(w, s4.z, s4.w, lbx, lby);
            md = (sqrt(sq(ubx - upx) + sq(uby - upy)) + sqrt(sq(lbx - lpx) + sq(lby - lpy))) / 2;  // rp.py:576-578
        }
        double wmd = md;
        int wl = lane;
#pragma unroll
        for (int s = 4; s > 0; s >>= 1) {
            const double o = __shfl_xor_sync(0xffffffffu, wmd, s);
            const int ol = __shfl_xor_sync(0xffffffffu, wl, s);
            if (o < wmd || (o == wmd && ol < wl)) { wmd = o; wl = ol; }
        }
        wl = __shfl_sync(0xffffffffu, wl, 0);
        wmd = __shfl_sync(0xffffffffu, wmd, 0);
        const double cux = __shfl_sync(0xffffffffu, ubx, wl), cuy = __shfl_sync(0xffffffffu, uby, wl);
        const double clx = __shfl_sync(0xffffffffu, lbx, wl), cly = __shfl_sync(0xffffffffu, lby, wl);
        if (wmd < best_md) { best_md = wmd; bux = cux; buy = cuy; blx = clx; bly = cly; }  // first minimum over the windows
        __syncwarp();
    }
    if (lane == 0) { out4[0] = bux; out4[1] = buy; out4[2] = blx; out4[3] = bly; }
    __syncwarp();
}

// rp.py:545-548 (largest free segment of the first waypoint, first maximum) for a ray with more than kMaxSeg free segments;
// leaves the winner in segs[0]
__device__ __noinline__ void pick_largest_windowed(int pitch, int n_wp, double ox, double oy, double res, double min_width,
                                                   const uint32_t* base, const uint32_t* cells, int len, int k, int nseg, int sx,
                                                   int sy, short4* segs) {
    const WinGeom w{pitch, n_wp, ox, oy, res, min_width};
    double bl = -1.0;
    short4 best = segs[0];
    for (int w0 = 0; w0 < nseg; w0 += kMaxSeg) {
        if (w0 > 0) refill_segments(base, cells, n_wp, k, len, sx, sy, pitch, ox, oy, res, min_width, segs, w0);
        const int cnt = nseg - w0 < kMaxSeg ? nseg - w0 : kMaxSeg;
        for (int i = 0; i < cnt; ++i) {
            const short4 c4 = segs[i];
            double ux, uy, lx, ly;
            m2w_s(w, c4.x, c4.y, ux, uy);
            m2w_s(w, c4.z, c4.w, lx, ly);
            const double l = sqrt(sq(ux - lx) + sq(uy - ly));
            if (l > bl) { bl = l; best = c4; }
        }
    }
    segs[0] = best;
}

struct RaycastArgs {
    const uint32_t* grids;
    size_t grid_stride_words;  // 0: every scenario uses the same grid
    GridView g;
    PathView pv;
    const int2* rowspan;
    const uint32_t* ray_cells;  // [max_len][n_wp]
    const int* ray_len;         // [n_wp], -1 = ray leaves the grid
    const int* wp_id;
    int first_offset, N;
    double min_width, sm;
    double *ub_out, *lb_out, *cells_sm_out;
    int* flags;
    int B;
    int stage_rows;  // rows of shared memory reserved per staging slab (0: read the grid from global memory)
    // fused K4a (closed-loop path): when state != nullptr lane 0 first localises the car and writes wp_id / spatial
    const double* state;
    int* wp_id_out;
    double* spatial_out;
    double length;
    // solve order for this step (closed-loop path; nullptr = off): one extra CTA of the launch buckets the scenarios by
    // the ADMM iteration count of their previous solve, longest first (see plan_solve_order)
    const int* prev_iters;
    int* order_out;
    int* long_out;  // host-mapped: number of live scenarios whose previous solve took >= kLongSolve iterations
    unsigned char* bucket_of;  // [B] scratch of the planner CTA (see plan_solve_order)
    // mpc_step_host on page-locked buffers: `state` then points into the CALLER's host memory (device-mapped); the kernel
    // leaves a device copy for the solve kernel's rollout and mirrors the flags it sets into the caller's flags
    HostMirror hm;
};

// lane 0: scenario b does not reach the solve kernel this step -- hand the caller its flags and the control on record
__device__ __forceinline__ void mirror_skipped(const RaycastArgs& a, int b, int fl) {
    if (a.hm.flags_host) a.hm.flags_host[b] = fl;
    if (a.hm.u_host) {
        a.hm.u_host[2 * (size_t)b] = a.hm.u_dev[2 * (size_t)b];
        a.hm.u_host[2 * (size_t)b + 1] = a.hm.u_dev[2 * (size_t)b + 1];
    }
}
__device__ __forceinline__ void flag_scenario(const RaycastArgs& a, int b, int bits) {  // lane 0
    if (!a.flags) return;
    const int old = atomicOr(&a.flags[b], bits);
    mirror_skipped(a, b, old | bits);
}

// Solve order for the paired ADMM kernel.  Its warps hold 2-4 scenarios that run in lockstep until the slowest is
// done, and CTAs are dispatched in index order; the iteration count of a car's previous solve predicts the next one
// well (the hard stretches of the track stay hard for several steps), so the scenarios are bucketed by it
// (25 iterations per bucket = one termination check), longest first: warp-mates get similar iteration counts and
// the long solves start first (shorter tail).  One CTA, shared-memory counting sort; the order inside a bucket is
// arbitrary, which cannot change any result: a scenario's arithmetic never depends on its warp-mates.
constexpr int kLongSolve = 300;

__device__ void plan_solve_order(const int* __restrict__ prev_iters, const int* __restrict__ flags, int* __restrict__ order,
                                 int* long_out, unsigned char* __restrict__ bucket_of, int B) {
    constexpr int NB = 192;  // 25 iterations per bucket: covers OSQP's default max_iter = 4000
    __shared__ int hist[NB], cursor[NB];
    for (int i = threadIdx.x; i < NB; i += blockDim.x) hist[i] = 0;
    __syncthreads();
    // The ray CTAs of the same launch set MPC_ST_FINISHED / MPC_ST_DEAD in `flags` while this CTA runs, so a scenario's
    // bucket is decided ONCE (histogram pass) and remembered in bucket_of[]; the scatter pass re-reads the remembered
    // value, never the flags, so histogram and cursors always describe the same assignment and `order` is a permutation
    // of 0..B-1.  A scenario that finishes during this launch merely keeps an early slot; the solve kernel re-reads its
    // flags and skips it.
    // Most scenarios of a step share two or three buckets, so per-lane shared-memory atomics would serialise on those
    // addresses.  Lanes with the same bucket are grouped with match.any and one of them adds the group's size: a handful
    // of atomics per warp.
    // The loads of four rounds are issued together: the CTA is on the step's critical path and would otherwise pay one
    // global-memory latency per round and operand.
    const int lane = threadIdx.x & 31;
    constexpr int U = 4;
    for (int b0 = threadIdx.x - lane; b0 < B; b0 += U * blockDim.x) {  // warp-uniform trip count
        int fl[U], it[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * blockDim.x + lane;
            fl[u] = (b < B && flags) ? flags[b] : 0;
            it[u] = b < B ? prev_iters[b] : 0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * blockDim.x + lane;
            int k = -1;
            if (b < B) {
                k = NB - 1;  // skipped by the solve: last
                if (!(fl[u] & (MPC_ST_DEAD | MPC_ST_FINISHED))) {
                    const int q = it[u] / 25;
                    k = NB - 1 - (q < 0 ? 0 : (q < NB - 1 ? q : NB - 1));
                }
                bucket_of[b] = (unsigned char)k;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, k);
            if (k >= 0 && lane == __ffs(peers) - 1) atomicAdd(&hist[k], __popc(peers));
        }
    }
    __syncthreads();
    __shared__ int wsum[NB / 32], nlong_s;
    if (threadIdx.x == 0) nlong_s = 0;
    if (blockDim.x >= NB) {
        // exclusive scan of the histogram by the first six warps (NB = 192 = 6 x 32 lanes): warp scan + carry
        if (threadIdx.x < NB) {
            const int w = threadIdx.x >> 5;
            int v = hist[threadIdx.x];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += t;
            }
            if (lane == 31) wsum[w] = v;
            cursor[threadIdx.x] = v - hist[threadIdx.x];  // exclusive within the warp
        }
        __syncthreads();
        if (threadIdx.x < NB) {
            int carry = 0;
            for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) carry += wsum[w];
            cursor[threadIdx.x] += carry;
            if (threadIdx.x < NB - 1 && (NB - 1 - (int)threadIdx.x) * 25 >= kLongSolve && hist[threadIdx.x])
                atomicAdd(&nlong_s, hist[threadIdx.x]);
        }
    } else if (threadIdx.x == 0) {  // small CTAs (the TMA-staged raycast variant): serial scan
        int run = 0, nl = 0;
        for (int i = 0; i < NB; ++i) {
            cursor[i] = run; run += hist[i];
            if (i < NB - 1 && (NB - 1 - i) * 25 >= kLongSolve) nl += hist[i];
        }
        nlong_s = nl;
    }
    __syncthreads();
    // feeds the host's choice between the paired and the lane-per-stage solve kernel for a LATER step (engine.cu)
    if (threadIdx.x == 0 && long_out) *long_out = nlong_s;
    for (int b0 = threadIdx.x - lane; b0 < B; b0 += U * blockDim.x) {
        int kk[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * blockDim.x + lane;
            kk[u] = b < B ? (int)bucket_of[b] : -1;  // written by this very thread in the histogram pass
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * blockDim.x + lane, k = kk[u];
            const unsigned peers = __match_any_sync(0xffffffffu, k);
            const int leader = __ffs(peers) - 1;
            int base = 0;
            if (k >= 0 && lane == leader) base = atomicAdd(&cursor[k], __popc(peers));
            base = __shfl_sync(0xffffffffu, base, leader);
            const int slot = base + __popc(peers & ((1u << lane) - 1u));
            if (k >= 0 && slot < B) order[slot] = b;  // always true for a consistent histogram; never write past the array
        }
    }
}

// MODE 0: no staging (global / L1 reads).  MODE 1: one grid shared by all scenarios, staged once per CTA
// (whole grid, one TMA bulk copy) and reused by all warps and all scenarios the CTA loops over.
// MODE 2: per-scenario grids, each warp stages the row span of its own scenario (A/B switch MPC_RAYCAST_MODE=2; the
// default for per-scenario grids is MODE 0 at 32 warps per SM, which is faster -- see raycast_plan).
#ifndef MPC_RAYCAST_MINB0
#define MPC_RAYCAST_MINB0 4
#endif
template <int MODE>
__global__ void __launch_bounds__(MODE == 1 ? 896 : 256, MODE == 0 ? MPC_RAYCAST_MINB0 : 1)
raycast_kernel(RaycastArgs a) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const GridView& g = a.g;
    const PathView& pv = a.pv;
    const int N = a.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ray_ctas = a.order_out ? (int)gridDim.x - 1 : (int)gridDim.x;
    if ((int)blockIdx.x == ray_ctas) {  // the extra CTA: plans the solve order while the others walk rays
        plan_solve_order(a.prev_iters, a.flags, a.order_out, a.long_out, a.bucket_of, a.B);
        return;
    }
    // shared layout: [mbarriers: 8 x u64][staging slabs][per-warp scratch: segs, prev_cells, nsegs]
    uint64_t* mbars = reinterpret_cast<uint64_t*>(smem_raw);
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + 128);
    const size_t slab_words = (size_t)a.stage_rows * g.pitch_words;
    const size_t stage_bytes = MODE == 1 ? slab_words * 4 : (MODE == 2 ? slab_words * 4 * nwarps : 0);
    const size_t scratch_per_warp = (size_t)N * kMaxSeg * sizeof(short4) + (size_t)N * 4 * sizeof(double) + (size_t)((N + 3) & ~3) * sizeof(int);
    unsigned char* scratch = smem_raw + 128 + stage_bytes + warp * scratch_per_warp;
    short4* segs = reinterpret_cast<short4*>(scratch);
    double* prev_cells = reinterpret_cast<double*>(segs + (size_t)N * kMaxSeg);
    int* nsegs = reinterpret_cast<int*>(prev_cells + (size_t)N * 4);
    uint32_t* srow = MODE == 2 ? stage + warp * slab_words : stage;

    if (MODE == 1) {
        if (threadIdx.x == 0) {
            mbar_init(&mbars[0], 1);
            tma_bulk_load(stage, a.grids, (uint32_t)(slab_words * 4), &mbars[0]);
        }
        __syncthreads();
        mbar_wait(&mbars[0], 0);
    } else if (MODE == 2) {
        if (lane == 0) mbar_init(&mbars[warp], 1);
        __syncwarp();
    }
    uint32_t phase = 0;
    for (int b = blockIdx.x * nwarps + warp; b < a.B; b += ray_ctas * nwarps) {
        const int fl = a.flags ? a.flags[b] : 0;
        if (fl & (MPC_ST_DEAD | MPC_ST_FINISHED)) {
            if (lane == 0) mirror_skipped(a, b, fl);
            continue;
        }
        int wp_now;
        if (a.state) {  // get_current_waypoint + t2s (sbm.py:256-279, 183-219), same arithmetic as localize_t2s_kernel
            const int w = localize_warp(a.state, a.spatial_out, pv, a.length, b, a.B, lane);
            if (w < 0) {
                if (lane == 0) flag_scenario(a, b, MPC_ST_FINISHED);
                continue;
            }
            if (lane == 0) a.wp_id_out[b] = w;
            wp_now = w;
        } else {
            wp_now = a.wp_id[b];
        }
        const long first = (long)wp_now + a.first_offset;
        int status = 0;
        if (!pv.circular && first + N - 1 >= pv.n_wp) status |= MPC_ST_END_OF_PATH;  // rp.py:367-369
        const int first_w = (int)(first % pv.n_wp);
        const uint32_t* gsrc = a.grids + (size_t)b * a.grid_stride_words;
        int row0 = 0;
        if (MODE == 2) {
            const int2 rs = a.rowspan[first_w];
            row0 = rs.x;
            __syncwarp();  // every lane is done reading the previous scenario's rows
            if (lane == 0)
                tma_bulk_load(srow, gsrc + (size_t)row0 * g.pitch_words, (uint32_t)(rs.y - rs.x + 1) * g.pitch_words * 4,
                              &mbars[warp]);
            mbar_wait(&mbars[warp], phase);
            phase ^= 1;
        }
        // ---- phase 1: free segments per horizon waypoint (rp.py:466-520), one lane per waypoint ----
        // MODE 0 reads the grid in global memory, MODES 1/2 the staged rows (row0 = first staged row)
        const uint32_t* base = MODE == 0 ? gsrc : srow - (size_t)row0 * g.pitch_words;
        for (int n0 = 0; n0 < N; n0 += 32) {
            const int n = n0 + lane;
            const int k = (first_w + (n < N ? n : 0)) % pv.n_wp;
            int len = n < N ? a.ray_len[k] : 0;
            if (len < 0) { status |= MPC_ST_INDEX_ERROR; len = 0; }
            const int max_len_warp = __reduce_max_sync(0xffffffffu, len);
            int ubx, uby;
            w2m(g, pv.border[4 * k], pv.border[4 * k + 1], ubx, uby);  // the ray's start cell (rp.py:478, 488)
            const int nseg = replay_ray(base, a.ray_cells, pv.n_wp, k, max_len_warp, ubx, uby, g.pitch_words, g.ox, g.oy,
                                        g.res, a.min_width, segs + (n < N ? n : 0) * kMaxSeg);
            if (n < N) nsegs[n] = nseg;  // may exceed kMaxSeg: such a ray is re-walked window by window below
        }
        __syncwarp();
        if (nsegs[0] == 0) status |= MPC_ST_NO_SEGMENT;  // rp.py:547 max([]) -> ValueError
        status = __reduce_or_sync(0xffffffffu, status);
        if (status) {
            if (lane == 0) flag_scenario(a, b, status | MPC_ST_DEAD);
            continue;
        }
        double* ub_o = a.ub_out + (size_t)b * N;
        double* lb_o = a.lb_out + (size_t)b * N;
        double* cs_o = a.cells_sm_out ? a.cells_sm_out + (size_t)b * N * 4 : nullptr;
        // ---- phase 2: waypoints whose pick does not depend on the previous one ----
        for (int n = lane; n < N; n += 32) {
            const int k = (first_w + n) % pv.n_wp;
            const int nseg = nsegs[n];
            if (n > 0 && nseg >= 2) continue;
            double ubx, uby, lbx, lby;
            if (nseg == 0) { ubx = pv.x[k]; uby = pv.y[k]; lbx = ubx; lby = uby; }  // rp.py:595
            else {
                short4 s4 = segs[n * kMaxSeg];
                if (n == 0 && nseg > kMaxSeg) {
                    int sx, sy;
                    w2m(g, pv.border[4 * k], pv.border[4 * k + 1], sx, sy);
                    pick_largest_windowed(g.pitch_words, pv.n_wp, g.ox, g.oy, g.res, a.min_width, base, a.ray_cells, a.ray_len[k], k,
                                          nseg, sx, sy, segs + n * kMaxSeg);
                    s4 = segs[n * kMaxSeg];
                } else if (n == 0 && nseg > 1) {  // largest segment, first maximum (rp.py:545-548)
                    double bl = -1.0;
                    for (int i = 0; i < nseg; ++i) {
                        const short4 c4 = segs[n * kMaxSeg + i];
                        double ux, uy, lx, ly;
                        m2w(g, c4.x, c4.y, ux, uy);
                        m2w(g, c4.z, c4.w, lx, ly);
                        const double l = sqrt(sq(ux - lx) + sq(uy - ly));
                        if (l > bl) { bl = l; s4 = c4; }
                    }
                }
                m2w(g, s4.x, s4.y, ubx, uby);
                m2w(g, s4.z, s4.w, lbx, lby);
            }
            const RayOut o = finalize_wp(pv, k, ubx, uby, lbx, lby, a.sm);
            ub_o[n] = o.ub;
            lb_o[n] = o.lb;
#pragma unroll
            for (int i = 0; i < 4; ++i) prev_cells[4 * n + i] = o.cells[i];
            if (cs_o)
#pragma unroll
                for (int i = 0; i < 4; ++i) cs_o[n * 4 + i] = o.cells_sm[i];
        }
        __syncwarp();
        // ---- phase 3: multi-candidate waypoints, in order (rp.py:552-586) ----
        for (int n = 1; n < N; ++n) {
            const int nseg = nsegs[n];  // warp-uniform
            if (nseg < 2) continue;
            const int k = (first_w + n) % pv.n_wp, kp = (first_w + n - 1) % pv.n_wp;
            const double ds = pv.ds_next[kp];  // wp_prev - wp (rp.py:558)
            const double upx = prev_cells[4 * (n - 1) + 0] + ds * pv.cos_psi[kp];  // rp.py:559
            const double upy = prev_cells[4 * (n - 1) + 1] + ds * pv.cos_psi[kp];  // rp.py:560 (quirk Q2)
            const double lpx = prev_cells[4 * (n - 1) + 2] + ds * pv.sin_psi[kp];  // rp.py:561
            const double lpy = prev_cells[4 * (n - 1) + 3] + ds * pv.sin_psi[kp];  // rp.py:562
            // candidates live in lanes 0 .. kMaxSeg-1
            double bux, buy, blx, bly;
            if (nseg <= kMaxSeg) {
                double md = INFINITY, ubx = 0, uby = 0, lbx = 0, lby = 0;
                if (lane < nseg) {
                    const short4 s4 = segs[n * kMaxSeg + lane];
                    m2w(g, s4.x, s4.y, ubx, uby);
                    m2w(g, s4.z, s4.w, lbx, lby);
                    const double d_ub = sqrt(sq(ubx - upx) + sq(uby - upy));  // rp.py:576
                    const double d_lb = sqrt(sq(lbx - lpx) + sq(lby - lpy));  // rp.py:577
                    md = (d_ub + d_lb) / 2;
                }
                double wmd = md;
                int wl = lane;
#pragma unroll
                for (int s = 4; s > 0; s >>= 1) {
                    const double o = __shfl_xor_sync(0xffffffffu, wmd, s);
                    const int ol = __shfl_xor_sync(0xffffffffu, wl, s);
                    if (o < wmd || (o == wmd && ol < wl)) { wmd = o; wl = ol; }  // list.index(min()) = first minimum
                }
                wl = __shfl_sync(0xffffffffu, wl, 0);
                bux = __shfl_sync(0xffffffffu, ubx, wl); buy = __shfl_sync(0xffffffffu, uby, wl);
                blx = __shfl_sync(0xffffffffu, lbx, wl); bly = __shfl_sync(0xffffffffu, lby, wl);
            } else {  // more free segments than the scratch holds: slow exact path, window by window
                int sx, sy;
                w2m(g, pv.border[4 * k], pv.border[4 * k + 1], sx, sy);
                double* out4 = prev_cells + 4 * n;  // scratch of this waypoint, overwritten with its cells right below
                pick_nearest_windowed(g.pitch_words, pv.n_wp, g.ox, g.oy, g.res, a.min_width, base, a.ray_cells, a.ray_len[k], k,
                                      nseg, sx, sy, segs + n * kMaxSeg, upx, upy, lpx, lpy, lane, out4);
                bux = out4[0]; buy = out4[1]; blx = out4[2]; bly = out4[3];
                __syncwarp();
            }
            if (lane == 0) {
                const RayOut o = finalize_wp(pv, k, bux, buy, blx, bly, a.sm);
                ub_o[n] = o.ub;
                lb_o[n] = o.lb;
#pragma unroll
                for (int i = 0; i < 4; ++i) prev_cells[4 * n + i] = o.cells[i];
                if (cs_o)
#pragma unroll
                    for (int i = 0; i < 4; ++i) cs_o[n * 4 + i] = o.cells_sm[i];
            }
            __syncwarp();
        }
        __syncwarp();
    }
}

// K4a + width-table replay.  With ONE grid shared by every scenario, update_path_constraints(wp_id, N, ...) is a pure
// function of the waypoint index (rp.py:522-648 reads the path, the grid and wp_id -- never the car's pose), so the engine
// ray-casts each of the n_wp horizons once per (path, grid, N, car width) with raycast_kernel itself (mpc_engine's
// width table) and a closed-loop step only localises the car and copies the row of its waypoint: bit-identical output,
// 200 distinct ray-casts instead of one per car per step.  Same launch shape as raycast_kernel: one warp per scenario,
// plus the planner CTA.  Per-scenario grids (obstacle scenarios) never take this path.
__global__ void __launch_bounds__(1024)
localize_gather_kernel(RaycastArgs a, const double* __restrict__ memo_ub, const double* __restrict__ memo_lb,
                       const int* __restrict__ memo_flags) {
    const PathView& pv = a.pv;
    const int N = a.N, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int ray_ctas = a.order_out ? (int)gridDim.x - 1 : (int)gridDim.x;
    if ((int)blockIdx.x == ray_ctas) {
        plan_solve_order(a.prev_iters, a.flags, a.order_out, a.long_out, a.bucket_of, a.B);
        return;
    }
    // With the caller's page-locked state as input (hm.state_dev_out set) the 32 scenarios of one round of the CTA are
    // fetched over PCIe by 128 threads as four coalesced 256 B reads, left in HBM for the solve kernel's rollout, and
    // handed to the warps through shared memory (two buffers: one barrier per round).
    __shared__ double staged[2][4][32];
    const bool from_host = a.state && a.hm.state_dev_out;
    int round = 0;
    for (int base = blockIdx.x * nwarps; base < a.B; base += ray_ctas * nwarps, ++round) {  // CTA-uniform trip count
        const int b = base + warp;
        if (from_host) {
            const int k = threadIdx.x / nwarps, j = threadIdx.x - k * nwarps;
            if (k < 4 && base + j < a.B) {
                const double v = a.state[(size_t)k * a.B + base + j];
                staged[round & 1][k][j] = v;
                a.hm.state_dev_out[(size_t)k * a.B + base + j] = v;
            }
            __syncthreads();
        }
        if (b >= a.B) continue;
        const int fl = a.flags ? a.flags[b] : 0;
        if (fl & (MPC_ST_DEAD | MPC_ST_FINISHED)) {
            if (lane == 0) mirror_skipped(a, b, fl);
            continue;
        }
        int w;
        if (a.state) {
            if (from_host) {
                const double(*sv)[32] = staged[round & 1];
                w = localize_warp_values(sv[0][warp], sv[1][warp], sv[2][warp], sv[3][warp], a.spatial_out, pv, a.length, b,
                                         a.B, lane);
            } else {
                w = localize_warp(a.state, a.spatial_out, pv, a.length, b, a.B, lane);
            }
            if (w < 0) {
                if (lane == 0) flag_scenario(a, b, MPC_ST_FINISHED);
                continue;
            }
            if (lane == 0) a.wp_id_out[b] = w;
        } else {
            w = a.wp_id[b];
        }
        const int st = memo_flags[w];  // what the ray-cast of this horizon reported (incl. MPC_ST_DEAD), 0 = fine
        if (st) {
            if (lane == 0) flag_scenario(a, b, st);
            continue;
        }
        for (int n = lane; n < N; n += 32) {
            a.ub_out[(size_t)b * N + n] = memo_ub[(size_t)w * N + n];
            a.lb_out[(size_t)b * N + n] = memo_lb[(size_t)w * N + n];
        }
    }
}

void launch_localize_gather(const PathView& pv, const double* memo_ub, const double* memo_lb, const int* memo_flags,
                            const int* wp_id, int N, double* ub, double* lb, int* flags, int B, cudaStream_t st,
                            const double* state, int* wp_id_out, double* spatial_out, double length, const int* prev_iters,
                            int* order_out, int* long_out, unsigned char* bucket_of, const HostMirror* mirror) {
    NvtxRange nvtx_("mpc:K4a+K3 localize + width-table replay");
    RaycastArgs a{};
    if (mirror) a.hm = *mirror;
    a.prev_iters = prev_iters; a.order_out = (order_out && bucket_of) ? order_out : nullptr; a.long_out = long_out; a.bucket_of = bucket_of;
    a.state = state; a.wp_id_out = wp_id_out; a.spatial_out = spatial_out; a.length = length;
    a.pv = pv; a.wp_id = wp_id; a.first_offset = 1; a.N = N; a.ub_out = ub; a.lb_out = lb; a.flags = flags; a.B = B;
    const int warps = 32;  // 1024 threads: the planner CTA of the launch sorts the batch four times faster than with 256
    const int need_ctas = (B + warps - 1) / warps, max_ctas = sm_count() * 2;
    const int grid = (need_ctas < max_ctas ? need_ctas : max_ctas) + (a.order_out ? 1 : 0);
    localize_gather_kernel<<<grid, warps * 32, 0, st>>>(a, memo_ub, memo_lb, memo_flags);
}

static size_t raycast_scratch_bytes(int N) {
    return (size_t)N * kMaxSeg * sizeof(short4) + (size_t)N * 4 * sizeof(double) + (size_t)((N + 3) & ~3) * sizeof(int);
}

// shared-memory bytes of one CTA for the given mode (stage_rows rows per slab)
size_t raycast_smem_bytes(const GridView& g, int N, int mode, int stage_rows, int warps) {
    size_t s = 128 + (size_t)warps * raycast_scratch_bytes(N);
    if (mode == 1) s += (size_t)stage_rows * g.pitch_words * 4;
    if (mode == 2) s += (size_t)warps * stage_rows * g.pitch_words * 4;
    return s;
}

// mode selection: 1 (shared grid, whole grid in shared memory) / 2 (per-scenario grids, row span per warp) when they
// fit, otherwise 0 (global reads).  Returns the launch geometry through the out parameters.
int raycast_plan(const GridView& g, int N, bool shared_grid, int max_rows, bool rowspan_ok, int* warps, int* stage_rows,
                 size_t* smem, int B) {
    const size_t kLimit = 200 * 1024;
    if (shared_grid) {
        // one big CTA per SM: a single staged copy of the grid serves 28 warps, which leaves most of the SM's
        // L1 / shared-memory array to the cache, where the ray table (the other per-cell operand) then lives
        // (measured at 4096 scenarios: 47 -> 36 us); sized so that the batch is one wave of warps when it can be
        static int n_sm = 0;
        if (!n_sm) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n_sm <= 0) n_sm = 148;
        }
        const char* e = getenv("MPC_RAYCAST_WARPS");
        int w = e ? atoi(e) : (B > 0 ? (B + n_sm - 1) / n_sm : 28);
        w = w < 8 ? 8 : (w > 28 ? 28 : w);
        *warps = w; *stage_rows = g.H;
        *smem = raycast_smem_bytes(g, N, 1, g.H, w);
        if (*smem <= kLimit) return 1;
        *warps = 8;
        *smem = raycast_smem_bytes(g, N, 1, g.H, 8);
        if (*smem <= kLimit / 2) return 1;
    } else if (rowspan_ok && getenv("MPC_RAYCAST_MODE") && getenv("MPC_RAYCAST_MODE")[0] == '2') {
        // Per-scenario grids, row span staged per warp by TMA: an A/B switch since round 2.  Measured at 8192 obstacle
        // scenarios (profiles/r2_raycast_modes.txt): 164 us staged against 114 us for the direct walk below.  The kernel is
        // a latency-bound instruction stream (5.9 k instructions per scenario, DRAM 7 % busy): staging serialises copy ->
        // wait -> replay inside a warp and its 23 KB slab per warp caps an SM at 8 warps, while the direct walk keeps eight
        // independent table entries per lane in flight on 32 warps per SM (64 registers) and hides the DRAM latency itself.
        for (int w = 4; w >= 1; w >>= 1) {
            *warps = w; *stage_rows = max_rows;
            *smem = raycast_smem_bytes(g, N, 2, max_rows, w);
            if (*smem <= kLimit / 2 || (w == 1 && *smem <= kLimit)) return 2;
        }
    }
    *warps = 8; *stage_rows = 0;
    *smem = raycast_smem_bytes(g, N, 0, 0, 8);
    return 0;
}

void launch_raycast(const uint32_t* grids, size_t grid_stride_words, const GridView& g, const PathView& pv,
                    const int2* rowspan, int max_rows, const uint32_t* ray_cells, const int* ray_len, const int* wp_id, int first_offset, int N, double min_width,
                    double sm, double* ub, double* lb, double* cells_sm, int* flags, int B, bool rowspan_ok,
                    cudaStream_t st, const double* state, int* wp_id_out, double* spatial_out, double length,
                    const int* prev_iters, int* order_out, int* long_out, unsigned char* bucket_of,
                    const HostMirror* mirror) {
    NvtxRange nvtx_("mpc:K3 raycast");
    RaycastArgs a;
    a.hm = mirror ? *mirror : HostMirror{nullptr, nullptr, nullptr, nullptr};
    a.prev_iters = prev_iters; a.order_out = (order_out && bucket_of) ? order_out : nullptr; a.long_out = long_out; a.bucket_of = bucket_of;
    a.state = state; a.wp_id_out = wp_id_out; a.spatial_out = spatial_out; a.length = length;
    a.grids = grids; a.grid_stride_words = grid_stride_words; a.g = g; a.pv = pv; a.rowspan = rowspan; a.wp_id = wp_id;
    a.ray_cells = ray_cells; a.ray_len = ray_len;
    a.first_offset = first_offset; a.N = N; a.min_width = min_width; a.sm = sm; a.ub_out = ub; a.lb_out = lb;
    a.cells_sm_out = cells_sm; a.flags = flags; a.B = B;
    int warps = 8, stage_rows = 0;
    size_t smem = 0;
    const int mode = raycast_plan(g, N, grid_stride_words == 0, max_rows, rowspan_ok, &warps, &stage_rows, &smem, B);
    a.stage_rows = stage_rows;
    const int ctas_needed = (B + warps - 1) / warps;
    // persistent-style grid: enough CTAs to fill the machine a few times over, each looping over scenarios
    const int max_ctas = sm_count() * 8;
    const int grid = (ctas_needed < max_ctas ? ctas_needed : max_ctas) + (a.order_out ? 1 : 0);
    if (mode == 1) {
        { static int have_ = 0; ensure_dynamic_smem(raycast_kernel<1>, have_, smem); }
        raycast_kernel<1><<<grid, warps * 32, smem, st>>>(a);
    } else if (mode == 2) {
        { static int have_ = 0; ensure_dynamic_smem(raycast_kernel<2>, have_, smem); }
        raycast_kernel<2><<<grid, warps * 32, smem, st>>>(a);
    } else {
        { static int have_ = 0; ensure_dynamic_smem(raycast_kernel<0>, have_, smem); }
        raycast_kernel<0><<<grid, warps * 32, smem, st>>>(a);
    }
}

// ------------------------------------------------------------------------------------------------
// K4 front: SpatialBicycleModel.get_current_waypoint + t2s (sbm.py:256-279, 183-219)
// ------------------------------------------------------------------------------------------------
void launch_localize(const double* state, int* wp_id, double* spatial, int* flags, const PathView& pv, double length,
                     int B, cudaStream_t st) {
    NvtxRange nvtx_("mpc:K4a localize_t2s");
    localize_t2s_kernel<<<(B + 255) / 256, 256, 0, st>>>(state, wp_id, spatial, flags, pv, length, B);
}

// ------------------------------------------------------------------------------------------------
// K4 back: SpatialBicycleModel.drive (sbm.py:221-244), explicit Euler, in place
// ------------------------------------------------------------------------------------------------
__global__ void rollout_kernel(double* __restrict__ state, const double* __restrict__ spatial,
                               const int* __restrict__ wp_id, const double* __restrict__ u, const int* __restrict__ flags,
                               PathView pv, double L, double Ts, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    if (flags && (flags[b] & (MPC_ST_DEAD | MPC_ST_FINISHED))) return;
    const double v = u[2 * (size_t)b], delta = u[2 * (size_t)b + 1];
    const double psi = state[2 * (size_t)B + b];
    const double x_dot = v * cos(psi);          // sbm.py:231
    const double y_dot = v * sin(psi);          // sbm.py:232
    const double psi_dot = v / L * tan(delta);  // sbm.py:233
    state[b] += x_dot * Ts;                     // sbm.py:237
    state[(size_t)B + b] += y_dot * Ts;
    state[2 * (size_t)B + b] = psi + psi_dot * Ts;
    const double e_y = spatial[b], e_psi = spatial[(size_t)B + b];
    const double s_dot = 1 / (1 - e_y * pv.kappa[wp_id[b]]) * v * cos(e_psi);  // sbm.py:240
    state[3 * (size_t)B + b] += s_dot * Ts;                                      // sbm.py:244
}

void launch_rollout(double* state, const double* spatial, const int* wp_id, const double* u, const int* flags,
                    const PathView& pv, double L, double Ts, int B, cudaStream_t st) {
    NvtxRange nvtx_("mpc:K4b rollout");
    rollout_kernel<<<(B + 255) / 256, 256, 0, st>>>(state, spatial, wp_id, u, flags, pv, L, Ts, B);
}

// ------------------------------------------------------------------------------------------------
// closed-loop statistics: per-scenario accumulation + final reduction
// ------------------------------------------------------------------------------------------------
// MPC.update_prediction (MPC.py:224-248) + SpatialBicycleModel.s2t (sbm.py:155-181) for every scenario: the predicted
// spatial states of stages 2 .. N-1 mapped back to world x / y about the horizon waypoints.  One thread per
// (scenario, stage); x_sol is the solver output in the reference's dec.x order.
__global__ void predict_xy_kernel(const double* __restrict__ x_sol, const int* __restrict__ wp_id, PathView pv, int N,
                                  double* __restrict__ xy, int B) {
    const int per = N - 2;
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (per <= 0 || i >= (long)B * per) return;
    const int b = (int)(i / per), n = 2 + (int)(i % per);
    const long w = (long)wp_id[b] + n;
    double x = NAN, y = NAN;
    if (pv.circular || w < pv.n_wp) {  // rp.py:356-371: a non-circular path ends (the reference exits there)
        const int k = (int)(w % pv.n_wp);
        const double e_y = x_sol[(size_t)b * (5 * N + 3) + 3 * n];
        x = pv.x[k] - e_y * pv.sin_psi[k];  // sbm.py:171
        y = pv.y[k] + e_y * pv.cos_psi[k];  // sbm.py:172
    }
    xy[2 * i] = x;
    xy[2 * i + 1] = y;
}

void launch_predict_xy(const double* x_sol, const int* wp_id, const PathView& pv, int N, double* xy, int B, cudaStream_t st) {
    NvtxRange nvtx_("mpc:predict_xy");
    const long n = (long)B * (N - 2);
    if (n <= 0) return;
    predict_xy_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x_sol, wp_id, pv, N, xy, B);
}

__global__ void accumulate_stats_kernel(const int* __restrict__ flags, const int* __restrict__ iters,
                                        const double* __restrict__ spatial, double* __restrict__ acc /*[4][B]*/, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const int f = flags[b];
    if (f & (MPC_ST_DEAD | MPC_ST_FINISHED)) return;
    acc[b] += 1.0;                                         // scenario-steps (= QP solves)
    acc[(size_t)B + b] += (double)iters[b];                // ADMM iterations
    acc[2 * (size_t)B + b] += (f & MPC_ST_QP_FALLBACK) ? 1.0 : 0.0;
    const double ey = fabs(spatial[b]);
    acc[3 * (size_t)B + b] += ey;
    acc[4 * (size_t)B + b] = fmax(acc[4 * (size_t)B + b], ey);
}

void launch_accumulate_stats(const int* flags, const int* iters, const double* spatial, double* acc, int B,
                             cudaStream_t st) {
    NvtxRange nvtx_("mpc:accumulate_stats");
    accumulate_stats_kernel<<<(B + 255) / 256, 256, 0, st>>>(flags, iters, spatial, acc, B);
}

}  // namespace mpcb
