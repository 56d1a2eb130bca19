// kernel_tiled.cuh — shared-memory tiled backward-induction kernels (placeholder plan: none yet).
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "dev_model.cuh"

namespace sdpb {

struct TiledPlan {
    bool available = false;
    const char* why_not = "not implemented";
};

inline void plan_tiled(TiledPlan& P, const sdpb_model&, const DevModel&, const std::vector<int>&,
                       const std::vector<int>&, const std::vector<int>&, bool, const cudaDeviceProp&) {
    P.available = false;
}

inline int launch_tiled(const TiledPlan&, const DevModel&, int, int, int, const double*, double*, int*,
                        long long, long long, cudaStream_t, double*) {
    return SDPB_ERR_ARG;
}

}  // namespace sdpb
