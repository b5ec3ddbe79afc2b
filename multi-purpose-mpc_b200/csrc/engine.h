// engine.h -- internal declarations shared by the translation units of libmpc_b200.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/mpc_b200.h"
#include "admm.cuh"

namespace mpcb {

struct HostIO {  // mpc_step_host: the caller's page-locked buffers as device pointers (all null = off)
    double* state;
    double* u;
    int* flags;
};

// mpc_step_host on page-locked buffers without copy nodes: what the first kernel of the step does on the caller's
// (device-mapped) host memory.  The solve kernel's epilogue writes the results of every scenario it solves (HostIO); the
// first kernel covers the ones it skips or retires.
struct HostMirror {
    double* state_dev_out;  // device copy of the state read from the host (localize_gather_kernel only), or null
    int* flags_host;        // scenarios skipped / retired by this kernel: flags ...
    const double* u_dev;    // ... and the control the device holds for them
    double* u_host;
};

struct GridView {
    int H, W, pitch_words;  // row pitch in 32-bit words (multiple of 16 -> 64 B)
    double ox, oy, res;
};

// geometry.cu
void launch_rasterize(const uint32_t* base, uint32_t* grids, int words, const GridView& g, const int* obs_px,
                      const int* offsets, int B, cudaStream_t st);
void launch_compute_width(const uint32_t* grid, const GridView& g, const PathView& pv, double max_width, double* ub,
                          double* lb, double* border, int* err, cudaStream_t st);
void launch_compute_width_batch(const uint32_t* grids, size_t grid_stride_words, const GridView& g, const double* tables,
                                int n_max, const int* n_wp, int T, double max_width, double* ub, double* lb, double* border,
                                int* err, cudaStream_t st);
int raycast_plan(const GridView& g, int N, bool shared_grid, int max_rows, bool rowspan_ok, int* warps, int* stage_rows,
                 size_t* smem, int B = 0);
void launch_build_ray_table(const GridView& g, const PathView& pv, uint32_t* cells, int* len, int max_len,
                            cudaStream_t st);
void launch_raycast(const uint32_t* grids, size_t grid_stride_words, const GridView& g, const PathView& pv,
                    const int2* rowspan, int max_rows, const uint32_t* ray_cells, const int* ray_len, const int* wp_id, int first_offset, int N, double min_width,
                    double sm, double* ub, double* lb, double* cells_sm, int* flags, int B, bool rowspan_ok,
                    cudaStream_t st, const double* state = nullptr, int* wp_id_out = nullptr,
                    double* spatial_out = nullptr, double length = 0.0, const int* prev_iters = nullptr,
                    int* order_out = nullptr, int* long_out = nullptr, unsigned char* bucket_of = nullptr,
                    const HostMirror* mirror = nullptr);
void launch_localize_gather(const PathView& pv, const double* memo_ub, const double* memo_lb, const int* memo_flags,
                            const int* wp_id, int N, double* ub, double* lb, int* flags, int B, cudaStream_t st,
                            const double* state = nullptr, int* wp_id_out = nullptr, double* spatial_out = nullptr,
                            double length = 0.0, const int* prev_iters = nullptr, int* order_out = nullptr,
                            int* long_out = nullptr, unsigned char* bucket_of = nullptr,
                            const HostMirror* mirror = nullptr);
void launch_localize(const double* state, int* wp_id, double* spatial, int* flags, const PathView& pv, double length,
                     int B, cudaStream_t st);
void launch_rollout(double* state, const double* spatial, const int* wp_id, const double* u, const int* flags,
                    const PathView& pv, double L, double Ts, int B, cudaStream_t st);
void launch_predict_xy(const double* x_sol, const int* wp_id, const PathView& pv, int N, double* xy, int B, cudaStream_t st);
void launch_accumulate_stats(const int* flags, const int* iters, const double* spatial, double* acc, int B,
                             cudaStream_t st);

// admm.cu
int launch_solve_qp(int precision, int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax,
                    const double* l, const double* u, double* x_out, int* iters, int* status, int B, cudaStream_t s);
int launch_assemble_solve(int precision, const MpcParams& mp, const AdmmSettings& st, const PathView& pv,
                          const double* spatial, const int* wp_id, double* control, const double* ub, const double* lb,
                          int* infeas, double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B,
                          cudaStream_t s, double* rollout_state = nullptr, double Ts = 0.0, const int* order = nullptr,
                          bool prefer_stage = false, const HostIO* host_io = nullptr);

void launch_build_stage_table(const PathView& pv, const MpcParams& mp, double* tab /*[n_wp][kStageTab]*/, cudaStream_t s);
bool solve_writes_host_io();  // false for the opt-in variants (MPC_ADMM_KERNEL=tm / quad), which only fill device buffers
void preload_solve_kernels(int precision, int N, int B = 0);
void preload_pair_kernels(int N);
void preload_quad_kernels(int N);
void preload_tm_kernels(int N);
int reserve_tm_scratch(int B);   // global-memory rows of the tensor-memory variant (admm_tm.cu), outside capture
int launch_assemble_solve_tm(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                             const int* wp_id, double* control, const double* ub, const double* lb, int* infeas, double* u_out,
                             double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s, double* rollout_state,
                             double Ts, const int* order);
int reserve_quad_scratch(int B);  // global-memory columns of the four-stages-per-lane kernel (admm_quad.cu), outside capture
int launch_assemble_solve_quad(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                               const int* wp_id, double* control, const double* ub, const double* lb, int* infeas,
                               double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s,
                               double* rollout_state, double Ts, const int* order);
// admm_pair.cu (fp32, N + 1 <= 64)
int launch_solve_qp_pair(int N, const AdmmSettings& st, const double* Pd, const double* q, const double* Ax, const double* l,
                         const double* u, double* x_out, int* iters, int* status, int B, cudaStream_t s);
int launch_assemble_solve_pair(const MpcParams& mp, const AdmmSettings& st, const PathView& pv, const double* spatial,
                               const int* wp_id, double* control, const double* ub, const double* lb, int* infeas,
                               double* u_out, double* x_out, int* iters, int* qp_status, int* flags, int B, cudaStream_t s,
                               double* rollout_state, double Ts, const int* order, const HostIO* host_io = nullptr);

}  // namespace mpcb
