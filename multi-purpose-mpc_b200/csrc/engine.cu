// engine.cu -- the C ABI of libmpc_b200.so (include/mpc_b200.h): handle, tables, launches, CUDA graph.
#include "engine.h"
#include <cstdlib>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace mpcb;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_OK(expr)                                                                                   \
    do {                                                                                                \
        cudaError_t e_ = (expr);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(MPC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));               \
    } while (0)

template <typename T> struct DevBuf {
    T* p = nullptr;
    size_t n = 0;
    bool owned = true;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc(&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    void view(T* ptr, size_t count) {  // a window into another allocation (not freed here)
        release();
        p = ptr; n = count; owned = false;
    }
    void release() {
        if (p && owned) cudaFree(p);
        p = nullptr;
        n = 0;
        owned = true;
    }
};

struct mpc_engine {
    mpc_config cfg;
    MpcParams mp;
    AdmmSettings st;
    cudaStream_t stream = nullptr;
    int64_t launches = 0;
    // path
    int n_wp = 0, circular = 1;
    double length = 0.0;
    DevBuf<double> d_wp;      // [13][n_wp]: 12 rows + length_cum
    DevBuf<double> d_border;  // [n_wp][4]
    DevBuf<double> d_stage_tab;  // [n_wp][kStageTab]: K1's per-waypoint coefficients (PathView::stage_tab)
    std::vector<double> h_wp, h_border;
    bool have_path = false, have_border = false;
    PathView pv{};
    // grid
    GridView g{};
    int words = 0;
    DevBuf<uint32_t> d_base, d_grids;
    DevBuf<int> d_obs_px, d_obs_off;
    std::vector<double> h_obs;     // per-scenario obstacle lists as given (world cx, cy, radius), kept so that a new
    std::vector<int> h_obs_off;    // base grid can be re-rasterised (apply_obstacles)
    int grids_B = 0;  // 0: shared base grid
    bool have_grid = false;
    DevBuf<int2> d_rowspan;
    DevBuf<uint32_t> d_ray_cells;  // ray table [max_len][n_wp] (geometry.cu)
    DevBuf<int> d_ray_len;
    int ray_max_len = 0;
    bool rowspan_valid = false;
    int max_rows = 0;
    DevBuf<int> d_err;
    // scenarios (engine-owned closed loop)
    int B = 0;
    DevBuf<double> s_io;  // [state 4B | u 2B | flags B (int)]: what mpc_step_host returns, one D2H copy
    DevBuf<double> s_state, s_spatial, s_control, s_ub, s_lb, s_u, s_acc;
    DevBuf<int> s_wp_id, s_iters, s_qp_status, s_flags, s_infeas;
    DevBuf<int> s_order;  // solve order of the closed-loop step (geometry.cu::plan_solve_order)
    DevBuf<unsigned char> s_bucket;  // its per-scenario scratch
    DevBuf<double> s_stats8;  // reduced statistics of mpc_run_closed_loop
    cudaGraphExec_t graph_exec[2] = {nullptr, nullptr};  // [0] paired solve kernel, [1] lane-per-stage solve kernel
    int* h_long = nullptr;   // host-mapped counter written by the solve-order planner (geometry.cu)
    int* d_long = nullptr;
    int graph_B = 0;
    // pinned staging for mpc_step_host
    double* pin_state = nullptr;
    double* pin_u = nullptr;
    int* pin_flags = nullptr;
    int pin_B = 0;
    // mpc_step_host on caller-pinned memory: H2D + the two step kernels + D2H as ONE graph launch, keyed by the pointers
    cudaGraphExec_t io_graph[2] = {nullptr, nullptr};
    const void* io_key[3] = {nullptr, nullptr, nullptr};
    int io_B = 0;
    const void* pinned_seen[3] = {nullptr, nullptr, nullptr};  // host pointers already found to be page-locked
    // profiling
    int profiling = 0;
    int no_solve_order = 0;  // MPC_SOLVE_ORDER=off: scenarios are solved in index order (A/B switch)
    // width table (shared grid only): update_path_constraints of every waypoint's horizon, ray-cast once per
    // (path, border cells, grid, N, car width) -- see geometry.cu::localize_gather_kernel
    DevBuf<double> memo_ub, memo_lb;   // [n_wp][N]
    DevBuf<int> memo_flags, memo_wp;   // [n_wp]
    bool memo_valid = false;
    int no_memo = 0;                   // MPC_WIDTH_MEMO=off: ray-cast per car per step (A/B switch, bench.py reports both)
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_launches[4] = {0, 0, 0, 0};
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

static void refresh_params(mpc_engine* h) {
    const mpc_config& c = h->cfg;
    MpcParams& m = h->mp;
    m.N = c.N;
    for (int i = 0; i < 3; ++i) { m.Q[i] = c.Q[i]; m.QN[i] = c.QN[i]; m.xmin[i] = c.xmin[i]; m.xmax[i] = c.xmax[i]; }
    for (int i = 0; i < 2; ++i) { m.R[i] = c.R[i]; m.umin[i] = c.umin[i]; m.umax[i] = c.umax[i]; }
    m.ay_max = c.ay_max;
    m.L = c.car_length;
    AdmmSettings& s = h->st;
    s.rho = c.rho; s.sigma = c.sigma; s.alpha = c.alpha; s.eps_abs = c.eps_abs; s.eps_rel = c.eps_rel;
    s.eps_prim_inf = c.eps_prim_inf; s.eps_dual_inf = c.eps_dual_inf;
    s.adaptive_rho_tolerance = c.adaptive_rho_tolerance;
    s.max_iter = c.max_iter; s.scaling = c.scaling; s.check_termination = c.check_termination;
    s.adaptive_rho_interval = c.adaptive_rho_interval;
}

static void drop_graph(mpc_engine* h) {
    h->memo_valid = false;  // every caller changes something the width table depends on (or B: rebuilt lazily, cheap)
    for (int i = 0; i < 2; ++i) {
        if (h->graph_exec[i]) cudaGraphExecDestroy(h->graph_exec[i]);
        h->graph_exec[i] = nullptr;
    }
    h->graph_B = 0;
    for (int i = 0; i < 2; ++i) {
        if (h->io_graph[i]) cudaGraphExecDestroy(h->io_graph[i]);
        h->io_graph[i] = nullptr;
    }
    h->io_B = 0;
    h->pinned_seen[0] = h->pinned_seen[1] = h->pinned_seen[2] = nullptr;
}

extern "C" int mpc_set_error_(int code, const char* msg) { return fail(code, msg); }

extern "C" {

void mpc_config_default(mpc_config* c) {
    memset(c, 0, sizeof(*c));
    c->N = 30;  // simulation.py:100-103
    c->Q[0] = 1.0; c->R[0] = 0.5; c->QN[0] = 1.0;
    for (int i = 0; i < 3; ++i) { c->xmin[i] = -INFINITY; c->xmax[i] = INFINITY; }  // simulation.py:110-111
    c->car_length = 0.12; c->car_width = 0.06; c->Ts = 0.05;                      // simulation.py:53-54
    const double kmax = std::tan(0.66) / c->car_length;                           // simulation.py:106-109
    c->umin[0] = 0.0; c->umin[1] = -kmax; c->umax[0] = 1.0; c->umax[1] = kmax;
    c->ay_max = 4.0;
    c->rho = 0.1; c->sigma = 1e-6; c->alpha = 1.6; c->eps_abs = 1e-3; c->eps_rel = 1e-3;
    c->eps_prim_inf = 1e-4; c->eps_dual_inf = 1e-4; c->max_iter = 4000; c->scaling = 10;
    c->check_termination = 25; c->adaptive_rho_interval = 25; c->adaptive_rho_tolerance = 5.0;
    c->precision = 0;
}

const char* mpc_last_error(void) { return g_err.c_str(); }
int mpc_abi_version(void) { return MPC_B200_ABI_VERSION; }

static int validate_cfg(const mpc_config* c) {
    if (c->N < 3 || c->N > 127) return fail(MPC_E_UNSUPPORTED, "horizon N must be in [3, 127]");
    if (!(c->car_length > 0) || !(c->Ts > 0)) return fail(MPC_E_INVALID, "car_length and Ts must be positive");
    if (c->precision != 0 && c->precision != 1) return fail(MPC_E_INVALID, "precision must be 0 (fp32) or 1 (fp64)");
    if (c->max_iter < 1 || c->check_termination < 0 || c->adaptive_rho_interval < 0 || c->scaling < 0)
        return fail(MPC_E_INVALID, "bad OSQP settings");
    return 0;
}

int mpc_engine_create(const mpc_config* cfg, mpc_engine** out) {
    if (!cfg || !out) return fail(MPC_E_INVALID, "null argument");
    if (int r = validate_cfg(cfg)) return r;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(MPC_E_CUDA, std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e));
    mpc_engine* h = new mpc_engine();
    h->cfg = *cfg;
    refresh_params(h);
    if (h->d_err.alloc(1) != cudaSuccess) { delete h; return fail(MPC_E_CUDA, "cudaMalloc failed"); }
    if (cudaHostAlloc(&h->h_long, sizeof(int), cudaHostAllocMapped) == cudaSuccess) {
        *h->h_long = 0;
        if (cudaHostGetDevicePointer(&h->d_long, h->h_long, 0) != cudaSuccess) h->d_long = nullptr;
    }
    {
        const char* e2 = getenv("MPC_SOLVE_ORDER");
        h->no_solve_order = (e2 && e2[0] == 'o' && e2[1] == 'f') ? 1 : 0;
        const char* e3 = getenv("MPC_WIDTH_MEMO");
        h->no_memo = (e3 && e3[0] == 'o' && e3[1] == 'f') ? 1 : 0;
    }
    cudaMemset(h->d_err.p, 0, sizeof(int));
    for (int i = 0; i < 5; ++i)
        if (cudaEventCreate(&h->ev[i]) != cudaSuccess) {
            mpc_engine_destroy(h);
            return fail(MPC_E_CUDA, "cudaEventCreate failed");
        }
    *out = h;
    return 0;
}

int mpc_engine_destroy(mpc_engine* h) {
    if (!h) return 0;
    cudaStreamSynchronize(h->stream);
    drop_graph(h);
    h->d_wp.release(); h->d_border.release(); h->d_base.release(); h->d_grids.release(); h->d_obs_px.release();
    h->d_obs_off.release(); h->d_rowspan.release(); h->d_err.release(); h->d_ray_cells.release(); h->d_ray_len.release();
    h->s_state.release(); h->s_spatial.release(); h->s_control.release(); h->s_ub.release(); h->s_lb.release();
    h->s_u.release(); h->s_acc.release(); h->s_wp_id.release(); h->s_iters.release(); h->s_qp_status.release();
    h->s_flags.release(); h->s_infeas.release(); h->s_order.release(); h->s_bucket.release(); h->s_io.release(); h->s_stats8.release();
    if (h->pin_state) cudaFreeHost(h->pin_state);  // pin_u / pin_flags point into it
    if (h->h_long) cudaFreeHost(h->h_long);
    for (int i = 0; i < 5; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    delete h;
    return 0;
}

int mpc_engine_set_stream(mpc_engine* h, void* s) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    h->stream = (cudaStream_t)s;
    drop_graph(h);
    return 0;
}

static int build_stage_table(mpc_engine* h);

int mpc_engine_update_config(mpc_engine* h, const mpc_config* cfg) {
    if (!h || !cfg) return fail(MPC_E_INVALID, "null argument");
    if (cfg->N != h->cfg.N) return fail(MPC_E_INVALID, "N cannot change after creation");
    if (int r = validate_cfg(cfg)) return r;
    h->cfg = *cfg;
    refresh_params(h);
    drop_graph(h);
    return build_stage_table(h);
}

int mpc_engine_sync(mpc_engine* h) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------
static void bind_path(mpc_engine* h) {
    const int n = h->n_wp;
    const double* p = h->d_wp.p;
    PathView& v = h->pv;
    v.n_wp = n; v.circular = h->circular;
    v.x = p; v.y = p + n; v.psi = p + 2 * n; v.kappa = p + 3 * n; v.v_ref = p + 4 * n; v.ds_next = p + 5 * n;
    v.cos_psi = p + 6 * n; v.sin_psi = p + 7 * n; v.cos_ub = p + 8 * n; v.sin_ub = p + 9 * n; v.cos_lb = p + 10 * n;
    v.sin_lb = p + 11 * n; v.length_cum = p + 12 * n;
    v.border = h->d_border.p;
    v.stage_tab = h->d_stage_tab.p;
}

// K1's per-waypoint coefficients depend on the path (ds, kappa), v_ref and R: rebuilt by whoever changes one of them
static int build_stage_table(mpc_engine* h) {
    if (!h->have_path) return 0;
    launch_build_stage_table(h->pv, h->mp, h->d_stage_tab.p, h->stream);
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// rows of the grid that the rays of a horizon starting at waypoint w (N waypoints) can touch
static int compute_rowspan(mpc_engine* h) {
    h->rowspan_valid = false;
    if (!h->have_path || !h->have_border || !h->have_grid) return 0;
    const int n = h->n_wp, N = h->cfg.N;
    std::vector<int> lo(n), hi(n);
    for (int k = 0; k < n; ++k) {
        const double* b = &h->h_border[4 * k];
        const int y0 = (int)std::floor((b[1] - h->g.oy) / h->g.res), y1 = (int)std::floor((b[3] - h->g.oy) / h->g.res);
        lo[k] = std::min(y0, y1) - 1;  // anti-aliasing side cells reach one row beyond the main chain
        hi[k] = std::max(y0, y1) + 1;
    }
    std::vector<int2> rs(n);
    int max_rows = 0;
    for (int w = 0; w < n; ++w) {
        int a = 1 << 30, b = -(1 << 30);
        for (int j = 0; j < N; ++j) {
            const int k = (w + j) % n;
            a = std::min(a, lo[k]);
            b = std::max(b, hi[k]);
        }
        a = std::max(a, 0);
        b = std::min(b, h->g.H - 1);
        if (b < a) { a = 0; b = 0; }
        rs[w] = make_int2(a, b);
        max_rows = std::max(max_rows, b - a + 1);
    }
    CUDA_OK(h->d_rowspan.alloc(n));
    CUDA_OK(cudaMemcpyAsync(h->d_rowspan.p, rs.data(), n * sizeof(int2), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->max_rows = max_rows;
    // ray table: one entry per emitted cell of every waypoint's ray; 3 cells per chain step bounds the length
    int max_len = 0;
    for (int k = 0; k < n; ++k) {
        const double* b = &h->h_border[4 * k];
        const long dx = std::labs((long)std::floor((b[0] - h->g.ox) / h->g.res) - (long)std::floor((b[2] - h->g.ox) / h->g.res));
        const long dy = std::labs((long)std::floor((b[1] - h->g.oy) / h->g.res) - (long)std::floor((b[3] - h->g.oy) / h->g.res));
        max_len = std::max(max_len, (int)std::min<long>(3 * (dx + dy) + 4, 1 << 20));
    }
    if ((size_t)h->words >= (1u << 26)) return fail(MPC_E_UNSUPPORTED, "grid too large for the packed ray table");
    max_len = (max_len + 7) & ~7;  // the replay fetches 8 entries at a time and reads the padding of shorter rays
    h->ray_max_len = max_len;
    CUDA_OK(h->d_ray_cells.alloc((size_t)max_len * n));
    CUDA_OK(h->d_ray_len.alloc(n));
    launch_build_ray_table(h->g, h->pv, h->d_ray_cells.p, h->d_ray_len.p, max_len, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->rowspan_valid = true;
    return 0;
}

int mpc_set_path(mpc_engine* h, const double* h_wp, const double* h_length_cum, const double* h_border, int32_t n_wp,
                 int32_t circular) {
    if (!h || !h_wp || !h_length_cum || n_wp < 2) return fail(MPC_E_INVALID, "bad path arguments");
    h->n_wp = n_wp;
    h->circular = circular ? 1 : 0;
    h->h_wp.assign(h_wp, h_wp + 12 * (size_t)n_wp);
    h->h_wp.insert(h->h_wp.end(), h_length_cum, h_length_cum + n_wp);
    h->length = h_length_cum[n_wp - 1];
    CUDA_OK(h->d_wp.alloc(13 * (size_t)n_wp));
    CUDA_OK(cudaMemcpyAsync(h->d_wp.p, h->h_wp.data(), 13 * (size_t)n_wp * sizeof(double), cudaMemcpyHostToDevice,
                            h->stream));
    CUDA_OK(h->d_border.alloc(4 * (size_t)n_wp));
    CUDA_OK(h->d_stage_tab.alloc(kStageTab * (size_t)n_wp));
    h->have_border = false;
    if (h_border) {
        h->h_border.assign(h_border, h_border + 4 * (size_t)n_wp);
        CUDA_OK(cudaMemcpyAsync(h->d_border.p, h_border, 4 * (size_t)n_wp * sizeof(double), cudaMemcpyHostToDevice,
                                h->stream));
        h->have_border = true;
    }
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->have_path = true;
    bind_path(h);
    drop_graph(h);
    if (int r = build_stage_table(h)) return r;
    return compute_rowspan(h);
}

int mpc_set_vref(mpc_engine* h, const double* h_vref, int32_t n_wp) {
    if (!h || !h_vref) return fail(MPC_E_INVALID, "null argument");
    if (!h->have_path || n_wp != h->n_wp) return fail(MPC_E_STATE, "mpc_set_path first (same n_wp)");
    memcpy(&h->h_wp[4 * (size_t)n_wp], h_vref, n_wp * sizeof(double));
    CUDA_OK(cudaMemcpyAsync(h->d_wp.p + 4 * (size_t)n_wp, h_vref, n_wp * sizeof(double), cudaMemcpyHostToDevice,
                            h->stream));
    return build_stage_table(h);  // synchronises
}

static int apply_obstacles(mpc_engine* h);

int mpc_set_base_grid(mpc_engine* h, const int8_t* data, int32_t H, int32_t W, double ox, double oy, double res) {
    if (!h || !data || H <= 0 || W <= 0 || !(res > 0)) return fail(MPC_E_INVALID, "bad grid arguments");
    if (H > 32767 || W > 32767) return fail(MPC_E_UNSUPPORTED, "grid dimensions above 32767 are not supported");
    GridView& g = h->g;
    g.H = H; g.W = W; g.ox = ox; g.oy = oy; g.res = res;
    g.pitch_words = ((W + 511) / 512) * 16;  // 64-byte multiple
    h->words = H * g.pitch_words;
    std::vector<uint32_t> bits((size_t)h->words, 0u);
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x)
            if (data[(size_t)y * W + x] == 1) bits[(size_t)y * g.pitch_words + (x >> 5)] |= 1u << (x & 31);
    CUDA_OK(h->d_base.alloc(h->words));
    CUDA_OK(cudaMemcpyAsync(h->d_base.p, bits.data(), (size_t)h->words * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->have_grid = true;
    h->grids_B = 0;
    drop_graph(h);
    if (int r = compute_rowspan(h)) return r;
    return apply_obstacles(h);  // per-scenario obstacle sets survive a change of the base map
}

// Map.add_obstacles per scenario (map.py:116-137): rasterise the remembered obstacle lists onto copies of the CURRENT base
// grid.  Called by mpc_set_obstacles and again by mpc_set_base_grid, so that a base-map change (Map.add_obstacles /
// add_boundary on the shared map, a new Map) never silently drops the per-scenario obstacles.
static int apply_obstacles(mpc_engine* h) {
    const int B = (int)h->h_obs_off.size() - 1;
    if (B <= 0) { h->grids_B = 0; return 0; }
    const int n_obs = h->h_obs_off[B];
    std::vector<int> px(3 * (size_t)std::max(n_obs, 1));
    for (int o = 0; o < n_obs; ++o) {
        const double cx = h->h_obs[3 * o], cy = h->h_obs[3 * o + 1], r = h->h_obs[3 * o + 2];
        px[3 * o] = (int)std::floor((cx - h->g.ox) / h->g.res);      // map.py:131 via w2m
        px[3 * o + 1] = (int)std::floor((cy - h->g.oy) / h->g.res);
        px[3 * o + 2] = (int)std::ceil(r / h->g.res);                // map.py:129
    }
    CUDA_OK(h->d_obs_px.alloc(px.size()));
    CUDA_OK(h->d_obs_off.alloc(B + 1));
    CUDA_OK(cudaMemcpyAsync(h->d_obs_px.p, px.data(), px.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->d_obs_off.p, h->h_obs_off.data(), (B + 1) * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(h->d_grids.alloc((size_t)B * h->words));
    launch_rasterize(h->d_base.p, h->d_grids.p, h->words, h->g, h->d_obs_px.p, h->d_obs_off.p, B, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(h->stream));
    h->grids_B = B;
    return 0;
}

int mpc_set_obstacles(mpc_engine* h, const double* h_obs, const int32_t* h_off, int32_t B) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    if (!h->have_grid) return fail(MPC_E_STATE, "mpc_set_base_grid first");
    drop_graph(h);
    if (B <= 0 || !h_obs || !h_off) { h->h_obs.clear(); h->h_obs_off.clear(); h->grids_B = 0; return 0; }
    if (h_off[0] != 0) return fail(MPC_E_INVALID, "obstacle offsets must start at 0");
    for (int b = 0; b < B; ++b)
        if (h_off[b + 1] < h_off[b]) return fail(MPC_E_INVALID, "obstacle offsets must be non-decreasing");
    h->h_obs.assign(h_obs, h_obs + 3 * (size_t)h_off[B]);
    h->h_obs_off.assign(h_off, h_off + B + 1);
    return apply_obstacles(h);
}

int mpc_get_grid(mpc_engine* h, int32_t b, int8_t* out) {
    if (!h || !out) return fail(MPC_E_INVALID, "null argument");
    if (!h->have_grid) return fail(MPC_E_STATE, "no grid");
    if (h->grids_B && (b < 0 || b >= h->grids_B)) return fail(MPC_E_INVALID, "scenario index out of range");
    const uint32_t* src = h->grids_B ? h->d_grids.p + (size_t)b * h->words : h->d_base.p;
    std::vector<uint32_t> bits((size_t)h->words);
    CUDA_OK(cudaMemcpy(bits.data(), src, (size_t)h->words * 4, cudaMemcpyDeviceToHost));
    for (int y = 0; y < h->g.H; ++y)
        for (int x = 0; x < h->g.W; ++x)
            out[(size_t)y * h->g.W + x] = (bits[(size_t)y * h->g.pitch_words + (x >> 5)] >> (x & 31)) & 1u;
    return 0;
}

int mpc_compute_width(mpc_engine* h, double max_width, double* h_ub, double* h_lb, double* h_border) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    if (!h->have_path || !h->have_grid) return fail(MPC_E_STATE, "mpc_set_path and mpc_set_base_grid first");
    const int n = h->n_wp;
    DevBuf<double> ub, lb;
    CUDA_OK(ub.alloc(n));
    CUDA_OK(lb.alloc(n));
    CUDA_OK(cudaMemsetAsync(h->d_err.p, 0, sizeof(int), h->stream));
    launch_compute_width(h->d_base.p, h->g, h->pv, max_width, ub.p, lb.p, h->d_border.p, h->d_err.p, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    h->h_border.resize(4 * (size_t)n);
    int err = 0;
    std::vector<double> tu(n), tl(n);
    CUDA_OK(cudaMemcpyAsync(tu.data(), ub.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaMemcpyAsync(tl.data(), lb.p, n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaMemcpyAsync(h->h_border.data(), h->d_border.p, 4 * (size_t)n * sizeof(double), cudaMemcpyDeviceToHost,
                            h->stream));
    CUDA_OK(cudaMemcpyAsync(&err, h->d_err.p, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    ub.release();
    lb.release();
    if (err) return fail(MPC_E_INVALID, "a width ray left the map (the reference raises IndexError, rp.py:279)");
    if (h_ub) memcpy(h_ub, tu.data(), n * sizeof(double));
    if (h_lb) memcpy(h_lb, tl.data(), n * sizeof(double));
    if (h_border) memcpy(h_border, h->h_border.data(), 4 * (size_t)n * sizeof(double));
    h->have_border = true;
    drop_graph(h);
    return compute_rowspan(h);
}

// ReferencePath._compute_width for T tracks in one launch (rp.py:206-287; SURVEY 8f-2: per-scenario BASE maps).  Stateless
// with respect to the engine: nothing of it changes the path / grid the engine steps on.
int mpc_compute_width_batch(mpc_engine* h, int32_t T, const int8_t* h_maps, int32_t H, int32_t W, double ox, double oy,
                            double res, const double* h_tables, const int32_t* h_n_wp, int32_t n_wp_max, double max_width,
                            double* h_ub, double* h_lb, double* h_border, int32_t* h_err) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    if (T <= 0) return 0;
    if (!h_maps || !h_tables || !h_n_wp || H <= 0 || W <= 0 || n_wp_max <= 0 || !(res > 0))
        return fail(MPC_E_INVALID, "bad arguments to mpc_compute_width_batch");
    for (int t = 0; t < T; ++t)
        if (h_n_wp[t] < 0 || h_n_wp[t] > n_wp_max) return fail(MPC_E_INVALID, "n_wp[t] outside [0, n_wp_max]");
    GridView g;
    g.H = H; g.W = W; g.ox = ox; g.oy = oy; g.res = res;
    g.pitch_words = ((W + 511) / 512) * 16;  // rows padded to 64 bytes, as mpc_set_base_grid does
    const size_t words = (size_t)H * g.pitch_words;
    std::vector<uint32_t> bits(words * (size_t)T, 0u);
    for (int t = 0; t < T; ++t) {
        const int8_t* data = h_maps + (size_t)t * H * W;
        uint32_t* b = bits.data() + (size_t)t * words;
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x)
                if (data[(size_t)y * W + x] == 1) b[(size_t)y * g.pitch_words + (x >> 5)] |= 1u << (x & 31);
    }
    DevBuf<uint32_t> d_bits;
    DevBuf<double> d_tab, d_ub, d_lb, d_border;
    DevBuf<int> d_n, d_err;
    struct Release {  // DevBuf has no destructor (engine members are released explicitly): free the temporaries on every path
        DevBuf<uint32_t>& a; DevBuf<double>&b, &c, &d, &e; DevBuf<int>&f, &g;
        ~Release() { a.release(); b.release(); c.release(); d.release(); e.release(); f.release(); g.release(); }
    } release_{d_bits, d_tab, d_ub, d_lb, d_border, d_n, d_err};
    cudaStream_t s = h->stream;
    const size_t nt = (size_t)T * n_wp_max;
    CUDA_OK(d_bits.alloc(bits.size()));
    CUDA_OK(d_tab.alloc(12 * nt));
    CUDA_OK(d_ub.alloc(nt));
    CUDA_OK(d_lb.alloc(nt));
    CUDA_OK(d_border.alloc(4 * nt));
    CUDA_OK(d_n.alloc(T));
    CUDA_OK(d_err.alloc(T));
    CUDA_OK(cudaMemcpyAsync(d_bits.p, bits.data(), bits.size() * 4, cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(d_tab.p, h_tables, 12 * nt * sizeof(double), cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemcpyAsync(d_n.p, h_n_wp, T * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(d_err.p, 0, T * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(d_ub.p, 0, nt * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(d_lb.p, 0, nt * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(d_border.p, 0, 4 * nt * sizeof(double), s));
    launch_compute_width_batch(d_bits.p, words, g, d_tab.p, n_wp_max, d_n.p, T, max_width, d_ub.p, d_lb.p, d_border.p, d_err.p, s);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    std::vector<int> err(T, 0);
    if (h_ub) CUDA_OK(cudaMemcpyAsync(h_ub, d_ub.p, nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_lb) CUDA_OK(cudaMemcpyAsync(h_lb, d_lb.p, nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    if (h_border) CUDA_OK(cudaMemcpyAsync(h_border, d_border.p, 4 * nt * sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaMemcpyAsync(err.data(), d_err.p, T * sizeof(int), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    int any = 0;
    for (int t = 0; t < T; ++t) {
        if (h_err) h_err[t] = err[t];
        any |= err[t];
    }
    // the reference raises IndexError for a ray that leaves the map (rp.py:279); per-track flags say which
    if (any && !h_err) return fail(MPC_E_INVALID, "a width ray left the map on at least one track (pass h_err to learn which)");
    return 0;
}

// ---------------------------------------------------------------------------------------------
static int need(mpc_engine* h, bool grid, bool border) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    if (!h->have_path) return fail(MPC_E_STATE, "mpc_set_path has not been called");
    if (grid && !h->have_grid) return fail(MPC_E_STATE, "mpc_set_base_grid has not been called");
    if (border && !h->have_border) return fail(MPC_E_STATE, "static border cells missing (mpc_compute_width / mpc_set_path)");
    return 0;
}

int mpc_localize_t2s(mpc_engine* h, const double* d_state, int32_t* d_wp_id, double* d_spatial, int32_t* d_flags,
                     int32_t B) {
    if (int r = need(h, false, false)) return r;
    if (B <= 0) return 0;
    launch_localize(d_state, d_wp_id, d_spatial, d_flags, h->pv, h->length, B, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_update_path_constraints(mpc_engine* h, const int32_t* d_wp_id, int32_t first_offset, int32_t N, double min_width,
                                double safety_margin, double* d_ub, double* d_lb, double* d_cells_sm, int32_t* d_flags,
                                int32_t B) {
    if (int r = need(h, true, true)) return r;
    if (B <= 0) return 0;
    if (N < 1 || N > 4096) return fail(MPC_E_INVALID, "bad N");
    if (h->grids_B && B > h->grids_B) return fail(MPC_E_INVALID, "B exceeds the number of scenario grids");
    const uint32_t* grids = h->grids_B ? h->d_grids.p : h->d_base.p;
    const size_t stride = h->grids_B ? (size_t)h->words : 0;
    // the row-span table is built for the engine's own horizon (first waypoint wp_id+1, N = cfg.N)
    const bool rowspan_ok = h->rowspan_valid && N == h->cfg.N && first_offset == 1;
    if (!h->rowspan_valid) return fail(MPC_E_STATE, "ray table not built (path, border cells and grid are all required)");
    launch_raycast(grids, stride, h->g, h->pv, h->d_rowspan.p, h->max_rows, h->d_ray_cells.p, h->d_ray_len.p, d_wp_id, first_offset, N, min_width,
                   safety_margin, d_ub, d_lb, d_cells_sm, d_flags, B, rowspan_ok, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_raycast(mpc_engine* h, const int32_t* d_wp_id, double* d_ub, double* d_lb, double* d_cells_sm, int32_t* d_flags,
                int32_t B) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    const double sm = h->cfg.car_width / std::sqrt(2.0);  // sbm.py:252
    return mpc_update_path_constraints(h, d_wp_id, 1, h->cfg.N, 2 * sm, sm, d_ub, d_lb, d_cells_sm, d_flags, B);
}

int mpc_assemble_solve(mpc_engine* h, const double* d_spatial, const int32_t* d_wp_id, double* d_control,
                       const double* d_ub, const double* d_lb, int32_t* d_infeas, double* d_u_out, double* d_x_out,
                       int32_t* d_iters, int32_t* d_qp_status, int32_t* d_flags, int32_t B) {
    if (int r = need(h, false, false)) return r;
    if (B <= 0) return 0;
    if (!d_spatial || !d_wp_id || !d_control || !d_ub || !d_lb || !d_infeas || !d_u_out)
        return fail(MPC_E_INVALID, "null device pointer");
    int r = launch_assemble_solve(h->cfg.precision, h->mp, h->st, h->pv, d_spatial, d_wp_id, d_control, d_ub, d_lb,
                                  d_infeas, d_u_out, d_x_out, d_iters, d_qp_status, d_flags, B, h->stream);
    if (r) return fail(r, "unsupported horizon");
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_solve_qp(mpc_engine* h, const double* d_Pd, const double* d_q, const double* d_Ax, const double* d_l,
                 const double* d_u, double* d_x_out, int32_t* d_iters, int32_t* d_qp_status, int32_t B) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    if (B <= 0) return 0;
    if (!d_Pd || !d_q || !d_Ax || !d_l || !d_u) return fail(MPC_E_INVALID, "null device pointer");
    int r = launch_solve_qp(h->cfg.precision, h->cfg.N, h->st, d_Pd, d_q, d_Ax, d_l, d_u, d_x_out, d_iters, d_qp_status,
                            B, h->stream);
    if (r) return fail(r, "unsupported horizon");
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_rollout(mpc_engine* h, double* d_state, const double* d_spatial, const int32_t* d_wp_id, const double* d_u,
                const int32_t* d_flags, int32_t B) {
    if (int r = need(h, false, false)) return r;
    if (B <= 0) return 0;
    launch_rollout(d_state, d_spatial, d_wp_id, d_u, d_flags, h->pv, h->cfg.car_length, h->cfg.Ts, B, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_predict_xy(mpc_engine* h, const double* d_x_sol, const int32_t* d_wp_id, double* d_xy_out, int32_t B) {
    if (int r = need(h, false, false)) return r;
    if (!d_x_sol || !d_wp_id || !d_xy_out) return fail(MPC_E_INVALID, "null argument");
    if (B <= 0 || h->cfg.N <= 2) return 0;
    launch_predict_xy(d_x_sol, d_wp_id, h->pv, h->cfg.N, d_xy_out, B, h->stream);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    return 0;
}

// ---------------------------------------------------------------------------------------------
// engine-owned scenarios
// ---------------------------------------------------------------------------------------------
int mpc_scenarios_init(mpc_engine* h, const double* h_state, int32_t B) {
    if (int r = need(h, true, true)) return r;
    if (B <= 0 || !h_state) return fail(MPC_E_INVALID, "bad scenario arguments");
    if (h->grids_B && B > h->grids_B) return fail(MPC_E_INVALID, "B exceeds the number of scenario grids");
    const int N = h->cfg.N;
    CUDA_OK(h->s_io.alloc(6 * (size_t)B + ((size_t)B + 1) / 2));
    h->s_state.view(h->s_io.p, 4 * (size_t)B);
    h->s_u.view(h->s_io.p + 4 * (size_t)B, 2 * (size_t)B);
    h->s_flags.view(reinterpret_cast<int*>(h->s_io.p + 6 * (size_t)B), (size_t)B);
    CUDA_OK(h->s_spatial.alloc(2 * (size_t)B));
    CUDA_OK(h->s_control.alloc(2 * (size_t)N * B));
    CUDA_OK(h->s_ub.alloc((size_t)N * B));
    CUDA_OK(h->s_lb.alloc((size_t)N * B));
    CUDA_OK(h->s_acc.alloc(5 * (size_t)B));
    CUDA_OK(h->s_wp_id.alloc(B));
    CUDA_OK(h->s_iters.alloc(B));
    CUDA_OK(h->s_qp_status.alloc(B));
    CUDA_OK(h->s_infeas.alloc(B));
    CUDA_OK(h->s_order.alloc(B));
    CUDA_OK(h->s_bucket.alloc(B));
    h->B = B;
    drop_graph(h);
    preload_solve_kernels(h->cfg.precision, N, B);
    return mpc_scenarios_set_state(h, h_state, nullptr, nullptr);
}

int mpc_scenarios_set_state(mpc_engine* h, const double* h_state, const double* h_control, const int32_t* h_infeas) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    const int B = h->B, N = h->cfg.N;
    cudaStream_t s = h->stream;
    if (h_state) CUDA_OK(cudaMemcpyAsync(h->s_state.p, h_state, 4 * (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s));
    if (h_control)
        CUDA_OK(cudaMemcpyAsync(h->s_control.p, h_control, 2 * (size_t)N * B * sizeof(double), cudaMemcpyHostToDevice, s));
    else
        CUDA_OK(cudaMemsetAsync(h->s_control.p, 0, 2 * (size_t)N * B * sizeof(double), s));  // MPC.py:56
    if (h_infeas) CUDA_OK(cudaMemcpyAsync(h->s_infeas.p, h_infeas, (size_t)B * sizeof(int), cudaMemcpyHostToDevice, s));
    else CUDA_OK(cudaMemsetAsync(h->s_infeas.p, 0, (size_t)B * sizeof(int), s));             // MPC.py:53
    CUDA_OK(cudaMemsetAsync(h->s_flags.p, 0, (size_t)B * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(h->s_iters.p, 0, (size_t)B * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(h->s_qp_status.p, 0, (size_t)B * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(h->s_acc.p, 0, 5 * (size_t)B * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(h->s_wp_id.p, 0, (size_t)B * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(h->s_spatial.p, 0, 2 * (size_t)B * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(h->s_u.p, 0, 2 * (size_t)B * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(h->s_ub.p, 0, (size_t)N * B * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(h->s_lb.p, 0, (size_t)N * B * sizeof(double), s));
    CUDA_OK(cudaStreamSynchronize(s));
    return 0;
}

int mpc_scenarios_set_flags(mpc_engine* h, const int32_t* h_flags) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (!h_flags) return fail(MPC_E_INVALID, "null flags");
    CUDA_OK(cudaMemcpyAsync(h->s_flags.p, h_flags, (size_t)h->B * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CUDA_OK(cudaStreamSynchronize(h->stream));
    return 0;
}

// One closed-loop step (simulation.py:137-140).  Fused path (default): two kernels -- K4a+K3 (localise inside the
// raycast kernel) and K1+K2+K4b (rollout behind the solve).  Profiling path: the four kernels of the ABI, bracketed by
// events, so that each gets its own duration.
// Which solve kernel the next step should use: warps of the paired kernel hold two scenarios and finish with the slower
// one, which costs when a few percent of the QPs run for hundreds of passes (infeasible ones: obstacle scenarios);
// then the lane-per-stage kernel's shorter per-solve latency wins (measured 2.35 vs 2.9 ms per step at 8192 obstacle
// scenarios).  The planner CTA of an earlier step counted the long solves into host-mapped memory; no sync here.
static bool prefer_stage_kernel(const mpc_engine* h) {
    if (!h->h_long || h->cfg.precision != 0) return false;
    const int nl = *reinterpret_cast<volatile const int*>(h->h_long);
    return (long)nl * 50 > (long)h->B;
}

// Builds the width table if the closed-loop step can use it (one grid shared by all scenarios).  NOT capturable (it
// synchronises): the entry points call it before they enqueue or capture a step.
static int ensure_width_memo(mpc_engine* h) {
    if (h->grids_B || h->no_memo || h->memo_valid) return 0;
    if (!h->rowspan_valid) return 0;  // launch_raycast reports the missing tables
    const int n = h->n_wp, N = h->cfg.N;
    const double sm = h->cfg.car_width / std::sqrt(2.0);
    CUDA_OK(h->memo_ub.alloc((size_t)n * N));
    CUDA_OK(h->memo_lb.alloc((size_t)n * N));
    CUDA_OK(h->memo_flags.alloc(n));
    CUDA_OK(h->memo_wp.alloc(n));
    std::vector<int> iota(n);
    for (int i = 0; i < n; ++i) iota[i] = i;
    cudaStream_t s = h->stream;
    CUDA_OK(cudaMemcpyAsync(h->memo_wp.p, iota.data(), n * sizeof(int), cudaMemcpyHostToDevice, s));
    CUDA_OK(cudaMemsetAsync(h->memo_flags.p, 0, n * sizeof(int), s));
    CUDA_OK(cudaMemsetAsync(h->memo_ub.p, 0, (size_t)n * N * sizeof(double), s));
    CUDA_OK(cudaMemsetAsync(h->memo_lb.p, 0, (size_t)n * N * sizeof(double), s));
    // virtual scenario w = "a car whose current waypoint is w": exactly the call MPC.get_control makes (MPC.py:116-118)
    launch_raycast(h->d_base.p, 0, h->g, h->pv, h->d_rowspan.p, h->max_rows, h->d_ray_cells.p, h->d_ray_len.p, h->memo_wp.p, 1, N,
                   2 * sm, sm, h->memo_ub.p, h->memo_lb.p, nullptr, h->memo_flags.p, n, h->rowspan_valid, s);
    ++h->launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s));
    h->memo_valid = true;
    return 0;
}

// `host` (mpc_step_host on page-locked buffers): the caller's state / u / flags as device pointers.  The kernels then read
// and write the caller's memory themselves and the step needs no copy before or after it.
static int enqueue_step(mpc_engine* h, bool with_stats, bool timed, bool stage_hint = false, const HostIO* host = nullptr) {
    const int B = h->B;
    cudaStream_t s = h->stream;
    const double sm = h->cfg.car_width / std::sqrt(2.0);
    const uint32_t* grids = h->grids_B ? h->d_grids.p : h->d_base.p;
    const size_t stride = h->grids_B ? (size_t)h->words : 0;
    if (!timed) {
        // the solve order is planned by one extra CTA of the raycast launch from the previous step's iteration counts
        int* order = (h->cfg.precision == 0 && B <= (1 << 16) && !h->no_solve_order) ? h->s_order.p : nullptr;
        HostMirror hm{nullptr, nullptr, nullptr, nullptr};
        if (host) hm = HostMirror{nullptr, host->flags, h->s_u.p, host->u};
        if (h->memo_valid && !h->grids_B) {
            if (host) hm.state_dev_out = h->s_state.p;  // this kernel fetches the state from the caller's memory itself
            launch_localize_gather(h->pv, h->memo_ub.p, h->memo_lb.p, h->memo_flags.p, h->s_wp_id.p, h->cfg.N, h->s_ub.p, h->s_lb.p,
                                   h->s_flags.p, B, s, host ? host->state : h->s_state.p, h->s_wp_id.p, h->s_spatial.p, h->length,
                                   h->s_iters.p, order, order ? h->d_long : nullptr, h->s_bucket.p, host ? &hm : nullptr);
        } else {
            launch_raycast(grids, stride, h->g, h->pv, h->d_rowspan.p, h->max_rows, h->d_ray_cells.p, h->d_ray_len.p, h->s_wp_id.p, 1, h->cfg.N, 2 * sm, sm,
                           h->s_ub.p, h->s_lb.p, nullptr, h->s_flags.p, B, h->rowspan_valid, s, h->s_state.p, h->s_wp_id.p,
                           h->s_spatial.p, h->length, h->s_iters.p, order, order ? h->d_long : nullptr, h->s_bucket.p,
                           host ? &hm : nullptr);
        }
        int r = launch_assemble_solve(h->cfg.precision, h->mp, h->st, h->pv, h->s_spatial.p, h->s_wp_id.p, h->s_control.p,
                                      h->s_ub.p, h->s_lb.p, h->s_infeas.p, h->s_u.p, nullptr, h->s_iters.p,
                                      h->s_qp_status.p, h->s_flags.p, B, s, h->s_state.p, h->cfg.Ts, order, stage_hint, host);
        if (r) return fail(r, "unsupported horizon");
        // the rollout only touches `state`; flags / iters / e_y of this step are final, so the statistics can follow it
        if (with_stats) launch_accumulate_stats(h->s_flags.p, h->s_iters.p, h->s_spatial.p, h->s_acc.p, B, s);
        h->launches += with_stats ? 3 : 2;
        CUDA_OK(cudaGetLastError());
        return 0;
    }
    cudaEventRecord(h->ev[0], s);
    launch_localize(h->s_state.p, h->s_wp_id.p, h->s_spatial.p, h->s_flags.p, h->pv, h->length, B, s);
    cudaEventRecord(h->ev[1], s);
    if (h->memo_valid && !h->grids_B)
        launch_localize_gather(h->pv, h->memo_ub.p, h->memo_lb.p, h->memo_flags.p, h->s_wp_id.p, h->cfg.N, h->s_ub.p, h->s_lb.p,
                               h->s_flags.p, B, s);
    else
        launch_raycast(grids, stride, h->g, h->pv, h->d_rowspan.p, h->max_rows, h->d_ray_cells.p, h->d_ray_len.p, h->s_wp_id.p, 1, h->cfg.N, 2 * sm, sm,
                       h->s_ub.p, h->s_lb.p, nullptr, h->s_flags.p, B, h->rowspan_valid, s);
    cudaEventRecord(h->ev[2], s);
    int r = launch_assemble_solve(h->cfg.precision, h->mp, h->st, h->pv, h->s_spatial.p, h->s_wp_id.p, h->s_control.p,
                                  h->s_ub.p, h->s_lb.p, h->s_infeas.p, h->s_u.p, nullptr, h->s_iters.p,
                                  h->s_qp_status.p, h->s_flags.p, B, s);
    if (r) return fail(r, "unsupported horizon");
    cudaEventRecord(h->ev[3], s);
    if (with_stats) launch_accumulate_stats(h->s_flags.p, h->s_iters.p, h->s_spatial.p, h->s_acc.p, B, s);
    launch_rollout(h->s_state.p, h->s_spatial.p, h->s_wp_id.p, h->s_u.p, h->s_flags.p, h->pv, h->cfg.car_length,
                   h->cfg.Ts, B, s);
    cudaEventRecord(h->ev[4], s);
    h->launches += with_stats ? 5 : 4;
    CUDA_OK(cudaGetLastError());
    return 0;
}

int mpc_step(mpc_engine* h) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (int r = need(h, true, true)) return r;
    if (int r = ensure_width_memo(h)) return r;
    return enqueue_step(h, false, false, prefer_stage_kernel(h));
}

static int ensure_graph(mpc_engine* h) {
    if (h->graph_exec[0] && h->graph_exec[1] && h->graph_B == h->B) return 0;
    drop_graph(h);
    if (int r0 = ensure_width_memo(h)) return r0;  // before the capture: it synchronises
    cudaStream_t cap = h->stream;
    cudaStream_t own = nullptr;
    if (cap == nullptr) {  // the legacy default stream cannot be captured
        CUDA_OK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
        cap = own;
    }
    cudaStream_t saved = h->stream;
    h->stream = cap;
    int r = 0;
    cudaError_t e = cudaSuccess;
    for (int v = 0; v < 2 && e == cudaSuccess && !r; ++v) {  // one graph per solve kernel (see prefer_stage_kernel)
        cudaGraph_t graph = nullptr;
        e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            const int64_t l0 = h->launches;
            r = enqueue_step(h, true, false, v == 1);
            h->launches = l0;
            e = cudaStreamEndCapture(cap, &graph);
        }
        if (e == cudaSuccess && !r) e = cudaGraphInstantiate(&h->graph_exec[v], graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    h->stream = saved;
    if (own) cudaStreamDestroy(own);
    if (e != cudaSuccess || r) {
        drop_graph(h);
        return r ? r : fail(MPC_E_CUDA, std::string("graph capture: ") + cudaGetErrorString(e));
    }
    h->graph_B = h->B;
    return 0;
}

__global__ void reduce_stats_kernel(const double* __restrict__ acc, const int* __restrict__ flags, int B,
                                    double* __restrict__ out8) {
    __shared__ double sh[8][256];
    double v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < B; b += blockDim.x) {
        v[0] += acc[b];
        v[1] += acc[b];
        v[2] += acc[(size_t)B + b];
        v[3] += acc[2 * (size_t)B + b];
        v[4] += (flags[b] & MPC_ST_DEAD) ? 1.0 : 0.0;
        v[5] += (flags[b] & MPC_ST_FINISHED) ? 1.0 : 0.0;
        v[6] += acc[3 * (size_t)B + b];
        v[7] = fmax(v[7], acc[4 * (size_t)B + b]);
    }
    for (int i = 0; i < 8; ++i) sh[i][threadIdx.x] = v[i];
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s)
            for (int i = 0; i < 8; ++i)
                sh[i][threadIdx.x] = i == 7 ? fmax(sh[i][threadIdx.x], sh[i][threadIdx.x + s])
                                            : sh[i][threadIdx.x] + sh[i][threadIdx.x + s];
        __syncthreads();
    }
    if (threadIdx.x < 8) out8[threadIdx.x] = sh[threadIdx.x][0];
}

int mpc_run_closed_loop(mpc_engine* h, int32_t max_steps, double* h_stats) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (int r = need(h, true, true)) return r;
    if (max_steps < 0) return fail(MPC_E_INVALID, "max_steps < 0");
    if (h->profiling) {
        if (int r = ensure_width_memo(h)) return r;
        for (int k = 0; k < max_steps; ++k) {
            if (int r = enqueue_step(h, true, true)) return r;
            CUDA_OK(cudaEventSynchronize(h->ev[4]));
            for (int i = 0; i < 4; ++i) {
                float ms = 0;
                cudaEventElapsedTime(&ms, h->ev[i], h->ev[i + 1]);
                h->prof_ms[i] += ms;
                h->prof_launches[i] += 1;
            }
        }
    } else {
        if (int r = ensure_graph(h)) return r;
        for (int k = 0; k < max_steps; ++k)
            CUDA_OK(cudaGraphLaunch(h->graph_exec[prefer_stage_kernel(h) ? 1 : 0], h->stream));
        h->launches += 3 * (int64_t)max_steps;
    }
    if (h_stats) {
        CUDA_OK(h->s_stats8.alloc(8));
        reduce_stats_kernel<<<1, 256, 0, h->stream>>>(h->s_acc.p, h->s_flags.p, h->B, h->s_stats8.p);
        ++h->launches;
        CUDA_OK(cudaMemcpyAsync(h_stats, h->s_stats8.p, 8 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
        CUDA_OK(cudaStreamSynchronize(h->stream));
    } else {
        CUDA_OK(cudaStreamSynchronize(h->stream));
    }
    return 0;
}

static int ensure_pinned_io(mpc_engine* h) {
    const int B = h->B;
    if (h->pin_B >= B && h->pin_state) return 0;
    const size_t io_bytes = 6 * (size_t)B * sizeof(double) + (size_t)B * sizeof(int);
    if (h->pin_state) cudaFreeHost(h->pin_state);
    h->pin_state = nullptr;
    h->pin_B = 0;
    CUDA_OK(cudaMallocHost(&h->pin_state, io_bytes));
    h->pin_u = h->pin_state + 4 * (size_t)B;
    h->pin_flags = reinterpret_cast<int*>(h->pin_state + 6 * (size_t)B);
    h->pin_B = B;
    return 0;
}

int mpc_host_io(mpc_engine* h, double** h_state, double** h_u, int32_t** h_flags) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (int r = ensure_pinned_io(h)) return r;
    if (h_state) *h_state = h->pin_state;
    if (h_u) *h_u = h->pin_u;
    if (h_flags) *h_flags = h->pin_flags;
    return 0;
}

// is [p, p + bytes) page-locked host memory the DMA engines can address directly?  (remembered per pointer)
static bool host_pinned(mpc_engine* h, int slot, const void* p, size_t bytes) {
    if (!p) return true;
    if (h->pinned_seen[slot] == p) return true;
    cudaPointerAttributes a0, a1;
    if (cudaPointerGetAttributes(&a0, p) != cudaSuccess ||
        cudaPointerGetAttributes(&a1, static_cast<const char*>(p) + bytes - 1) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    if (a0.type != cudaMemoryTypeHost || a1.type != cudaMemoryTypeHost) return false;
    h->pinned_seen[slot] = p;
    return true;
}

// the caller's page-locked buffers as the device sees them; false if any of them is not mapped into the device's space
static bool host_device_pointers(double* h_state, double* h_u_out, int32_t* h_flags, HostIO* out) {
    void *ds = nullptr, *du = nullptr, *df = nullptr;
    if (cudaHostGetDevicePointer(&ds, h_state, 0) != cudaSuccess || cudaHostGetDevicePointer(&du, h_u_out, 0) != cudaSuccess ||
        (h_flags && cudaHostGetDevicePointer(&df, h_flags, 0) != cudaSuccess)) {
        cudaGetLastError();
        return false;
    }
    *out = HostIO{static_cast<double*>(ds), static_cast<double*>(du), static_cast<int*>(df)};
    return true;
}

// One graph, one launch per step on page-locked caller buffers.  Default: NO copy nodes -- the first kernel reads the state
// from the caller's memory (coalesced, localize_gather_kernel) and the solve kernel's epilogue stores new state, control
// and flags there (admm_epilogue.cuh: store_host_results); scenarios the step skips are covered by the first kernel.  With
// per-scenario grids the ray-cast kernel keeps its device-side state, so that path has one H2D node in front.
// MPC_HOST_IO=copy, or a solve kernel without the host epilogue (MPC_ADMM_KERNEL=tm / quad), selects the graph
// H2D state -> kernels -> D2H (state, u, flags).
static int ensure_io_graph(mpc_engine* h, double* h_state, double* h_u_out, int32_t* h_flags) {
    if (h->io_graph[0] && h->io_graph[1] && h->io_B == h->B && h->io_key[0] == h_state && h->io_key[1] == h_u_out &&
        h->io_key[2] == h_flags)
        return 0;
    for (int i = 0; i < 2; ++i) {
        if (h->io_graph[i]) cudaGraphExecDestroy(h->io_graph[i]);
        h->io_graph[i] = nullptr;
    }
    const size_t B = h->B;
    if (int r0 = ensure_width_memo(h)) return r0;  // before the capture: it synchronises
    const char* io_env = getenv("MPC_HOST_IO");
    HostIO mapped{nullptr, nullptr, nullptr};
    const bool zero_copy = !(io_env && io_env[0] == 'c') && solve_writes_host_io() &&
                           host_device_pointers(h_state, h_u_out, h_flags, &mapped);
    const bool kernel_reads_host = zero_copy && h->memo_valid && !h->grids_B;
    cudaStream_t cap = h->stream, own = nullptr;
    if (cap == nullptr) {
        CUDA_OK(cudaStreamCreateWithFlags(&own, cudaStreamNonBlocking));
        cap = own;
    }
    cudaStream_t saved = h->stream;
    h->stream = cap;
    int r = 0;
    cudaError_t e = cudaSuccess;
    const bool one_block = h_u_out == h_state + 4 * B && (!h_flags || h_flags == reinterpret_cast<int32_t*>(h_state + 6 * B));
    for (int v = 0; v < 2 && e == cudaSuccess && !r; ++v) {
        cudaGraph_t graph = nullptr;
        e = cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal);
        if (e != cudaSuccess) break;
        const int64_t l0 = h->launches;
        if (!kernel_reads_host) cudaMemcpyAsync(h->s_state.p, h_state, 4 * B * sizeof(double), cudaMemcpyHostToDevice, cap);
        r = enqueue_step(h, false, false, v == 1, zero_copy ? &mapped : nullptr);
        h->launches = l0;
        if (zero_copy) {
        } else if (one_block) {  // the caller's buffers are laid out like s_io (mpc_host_io): one copy
            cudaMemcpyAsync(h_state, h->s_io.p, 6 * B * sizeof(double) + (h_flags ? B * sizeof(int) : 0), cudaMemcpyDeviceToHost, cap);
        } else {
            cudaMemcpyAsync(h_state, h->s_state.p, 4 * B * sizeof(double), cudaMemcpyDeviceToHost, cap);
            cudaMemcpyAsync(h_u_out, h->s_u.p, 2 * B * sizeof(double), cudaMemcpyDeviceToHost, cap);
            if (h_flags) cudaMemcpyAsync(h_flags, h->s_flags.p, B * sizeof(int), cudaMemcpyDeviceToHost, cap);
        }
        e = cudaStreamEndCapture(cap, &graph);
        if (e == cudaSuccess && !r) e = cudaGraphInstantiate(&h->io_graph[v], graph, 0);
        if (graph) cudaGraphDestroy(graph);
    }
    h->stream = saved;
    if (own) cudaStreamDestroy(own);
    if (e != cudaSuccess || r) {
        for (int i = 0; i < 2; ++i) {
            if (h->io_graph[i]) cudaGraphExecDestroy(h->io_graph[i]);
            h->io_graph[i] = nullptr;
        }
        return r ? r : fail(MPC_E_CUDA, std::string("io graph capture: ") + cudaGetErrorString(e));
    }
    h->io_B = h->B;
    h->io_key[0] = h_state; h->io_key[1] = h_u_out; h->io_key[2] = h_flags;
    return 0;
}

int mpc_step_host(mpc_engine* h, double* h_state, double* h_u_out, int32_t* h_flags) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (int r = need(h, true, true)) return r;
    if (!h_state || !h_u_out) return fail(MPC_E_INVALID, "null host pointer");
    const size_t B = h->B;
    cudaStream_t s = h->stream;
    // page-locked caller buffers (mpc_host_io, cudaHostAlloc / cudaHostRegister, torch pin_memory): no staging, and the
    // whole step -- both copies included -- is one graph launch
    if (host_pinned(h, 0, h_state, 4 * B * sizeof(double)) && host_pinned(h, 1, h_u_out, 2 * B * sizeof(double)) &&
        host_pinned(h, 2, h_flags, B * sizeof(int))) {
        if (int r = ensure_io_graph(h, h_state, h_u_out, h_flags)) return r;
        CUDA_OK(cudaGraphLaunch(h->io_graph[prefer_stage_kernel(h) ? 1 : 0], s));
        h->launches += 2;
        CUDA_OK(cudaStreamSynchronize(s));
        return 0;
    }
    // pageable caller buffers: staged through the engine's pinned block
    if (int r = ensure_pinned_io(h)) return r;
    if (int r = ensure_width_memo(h)) return r;
    const size_t io_bytes = 6 * B * sizeof(double) + B * sizeof(int);
    memcpy(h->pin_state, h_state, 4 * B * sizeof(double));
    CUDA_OK(cudaMemcpyAsync(h->s_state.p, h->pin_state, 4 * B * sizeof(double), cudaMemcpyHostToDevice, s));
    if (int r = enqueue_step(h, false, false, prefer_stage_kernel(h))) return r;
    // state | u | flags are one device allocation (s_io): one device-to-host copy
    CUDA_OK(cudaMemcpyAsync(h->pin_state, h->s_io.p, io_bytes, cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    memcpy(h_state, h->pin_state, 4 * B * sizeof(double));
    memcpy(h_u_out, h->pin_u, 2 * B * sizeof(double));
    if (h_flags) memcpy(h_flags, h->pin_flags, B * sizeof(int));
    return 0;
}

int mpc_scenarios_ptrs(mpc_engine* h, double** d_state, double** d_spatial, int32_t** d_wp_id, double** d_control,
                       double** d_ub, double** d_lb, double** d_u, int32_t** d_iters, int32_t** d_qp_status,
                       int32_t** d_flags, int32_t** d_infeas) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    if (d_state) *d_state = h->s_state.p;
    if (d_spatial) *d_spatial = h->s_spatial.p;
    if (d_wp_id) *d_wp_id = h->s_wp_id.p;
    if (d_control) *d_control = h->s_control.p;
    if (d_ub) *d_ub = h->s_ub.p;
    if (d_lb) *d_lb = h->s_lb.p;
    if (d_u) *d_u = h->s_u.p;
    if (d_iters) *d_iters = h->s_iters.p;
    if (d_qp_status) *d_qp_status = h->s_qp_status.p;
    if (d_flags) *d_flags = h->s_flags.p;
    if (d_infeas) *d_infeas = h->s_infeas.p;
    return 0;
}

int mpc_scenarios_read(mpc_engine* h, double* h_state, double* h_control, double* h_u, int32_t* h_iters,
                       int32_t* h_qp_status, int32_t* h_flags, int32_t* h_infeas, int32_t* h_wp_id, double* h_ub,
                       double* h_lb) {
    if (!h || h->B <= 0) return fail(MPC_E_STATE, "mpc_scenarios_init first");
    const size_t B = h->B, N = h->cfg.N;
    cudaStream_t s = h->stream;
    CUDA_OK(cudaStreamSynchronize(s));
    if (h_state) CUDA_OK(cudaMemcpy(h_state, h->s_state.p, 4 * B * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_control) CUDA_OK(cudaMemcpy(h_control, h->s_control.p, 2 * N * B * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_u) CUDA_OK(cudaMemcpy(h_u, h->s_u.p, 2 * B * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_iters) CUDA_OK(cudaMemcpy(h_iters, h->s_iters.p, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_qp_status) CUDA_OK(cudaMemcpy(h_qp_status, h->s_qp_status.p, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_flags) CUDA_OK(cudaMemcpy(h_flags, h->s_flags.p, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_infeas) CUDA_OK(cudaMemcpy(h_infeas, h->s_infeas.p, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_wp_id) CUDA_OK(cudaMemcpy(h_wp_id, h->s_wp_id.p, B * sizeof(int), cudaMemcpyDeviceToHost));
    if (h_ub) CUDA_OK(cudaMemcpy(h_ub, h->s_ub.p, N * B * sizeof(double), cudaMemcpyDeviceToHost));
    if (h_lb) CUDA_OK(cudaMemcpy(h_lb, h->s_lb.p, N * B * sizeof(double), cudaMemcpyDeviceToHost));
    return 0;
}

int64_t mpc_launch_count(mpc_engine* h) { return h ? h->launches : 0; }

int mpc_set_profiling(mpc_engine* h, int32_t on) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    h->profiling = on ? 1 : 0;
    for (int i = 0; i < 4; ++i) { h->prof_ms[i] = 0; h->prof_launches[i] = 0; }
    return 0;
}

int mpc_get_profile(mpc_engine* h, double* h_ms, int64_t* h_launches) {
    if (!h) return fail(MPC_E_INVALID, "null engine");
    for (int i = 0; i < 4; ++i) {
        if (h_ms) h_ms[i] = h->prof_ms[i];
        if (h_launches) h_launches[i] = h->prof_launches[i];
    }
    return 0;
}

}  // extern "C"
