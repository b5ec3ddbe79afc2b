// launch_util.h -- host-side helpers shared by the launch functions.
// NVTX ranges around the engine's kernel launches (per-kernel ranges for nsys / ncu
// --nvtx filtering).  Header-only NVTX3: without a profiler attached a push / pop is one indirect call on a null table.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

namespace mpcb {
struct NvtxRange {
    explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
    ~NvtxRange() { nvtxRangePop(); }
    NvtxRange(const NvtxRange&) = delete;
    NvtxRange& operator=(const NvtxRange&) = delete;
};

// Raise a kernel's dynamic shared-memory limit when a launch needs more than any launch before it -- not on every launch.
// `have` is a function-local static of the caller (one per kernel instantiation; one process drives one GPU).
template <typename K> inline void ensure_dynamic_smem(K kernel, int& have, size_t need) {
    if ((int)need > have) {
        cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need);
        have = (int)need;
    }
}

// SM count of the current device (148 on a B200), queried once
inline int sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    }
    return n;
}
}  // namespace mpcb
