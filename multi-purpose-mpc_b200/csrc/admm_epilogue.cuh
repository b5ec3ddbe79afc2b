// admm_epilogue.cuh -- pieces shared by the ADMM kernel translation units (admm.cu, admm_pair.cu)
#pragma once
#include "engine.h"

namespace mpcb {

// BicycleModel.drive (sbm.py:221-244) fused behind the solve on the closed-loop path.  Explicit round-to-nearest
// intrinsics: this file is compiled with FMA contraction on, the standalone rollout_kernel without, and both
// must produce the same bits.
__device__ __forceinline__ void drive_one(double* __restrict__ state, int b, int B, double e_y, double e_psi,
                                          double kappa_wp, double v, double delta, double L, double Ts) {
    const double psi = state[2 * (size_t)B + b];
    const double x_dot = __dmul_rn(v, cos(psi));                       // sbm.py:231
    const double y_dot = __dmul_rn(v, sin(psi));                       // sbm.py:232
    const double psi_dot = __dmul_rn(__ddiv_rn(v, L), tan(delta));     // sbm.py:233
    state[b] = __dadd_rn(state[b], __dmul_rn(x_dot, Ts));              // sbm.py:237
    state[(size_t)B + b] = __dadd_rn(state[(size_t)B + b], __dmul_rn(y_dot, Ts));
    state[2 * (size_t)B + b] = __dadd_rn(psi, __dmul_rn(psi_dot, Ts));
    const double s_dot = __dmul_rn(__dmul_rn(__ddiv_rn(1.0, __dsub_rn(1.0, __dmul_rn(e_y, kappa_wp))), v), cos(e_psi));  // sbm.py:240
    state[3 * (size_t)B + b] = __dadd_rn(state[3 * (size_t)B + b], __dmul_rn(s_dot, Ts));  // sbm.py:244
}

struct RolloutArgs {  // non-null state: fuse the rollout of this scenario behind its solve
    double* state;
    const double* spatial;
    const double* kappa;
    int wp;
    double Ts;
    int B;
    // mpc_step_host on page-locked buffers: the epilogue also stores the step's results straight into the caller's host memory
    // (device-mapped pointers; null = off), so that the step needs no device-to-host copy node
    double* host_state = nullptr;
    double* host_u = nullptr;
    int* host_flags = nullptr;
};

// lane 0 of a scenario, after the epilogue wrote the device copies
__device__ __forceinline__ void store_host_results(const RolloutArgs& ro, const double* u_out, int b, int fl) {
    if (ro.host_flags) ro.host_flags[b] = fl;
    if (ro.host_u) { ro.host_u[2 * (size_t)b] = u_out[2 * (size_t)b]; ro.host_u[2 * (size_t)b + 1] = u_out[2 * (size_t)b + 1]; }
    if (ro.host_state && ro.state) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ro.host_state[(size_t)k * ro.B + b] = ro.state[(size_t)k * ro.B + b];
    }
}

}  // namespace mpcb
