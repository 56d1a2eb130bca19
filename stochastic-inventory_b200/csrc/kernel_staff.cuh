// kernel_staff.cuh — workforce planning with an ACTION-DEPENDENT demand pmf:
// state = staff on hand, action = hires, turnover j ~ pmf chosen by the hire-up-to level y = x + a
// (src/workforce/WorkforcePlanning.java:71-104 lambdas, loop src/workforce/StaffRecursion.java:81-121).
// A warp per state, lanes stride over the actions (each lane runs its own serial j loop over its own pmf
// row), warp-shuffle (value, action) argopt with the first-wins rule.
#pragma once
#include <cuda_runtime.h>

#include "dev_model.cuh"

namespace sdpb {

template <bool IS_MIN, bool LAST>
__global__ void __launch_bounds__(256)
bi_staff(const __grid_constant__ DevModel M, const int t, const double* __restrict__ Vn,
         double* __restrict__ Vt, int* __restrict__ Qt, const long long lo, const long long hi) {
    const long long idx = lo + (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const bool live = idx < hi;
    const int ix = (int)(live ? idx : lo);
    const int xmin = (int)M.inv_min;
    const int xv = xmin + ix;  // iniStaffNum
    const int minStaff = (int)M.min_level_t[t - 1];
    const double v = M.v_t[t - 1];
    const int nA = M.max_order_idx + 1;
    const int* __restrict__ alen = M.apmf_len + (size_t)(t - 1) * M.nI;
    const int* __restrict__ aoff = M.apmf_off + (size_t)(t - 1) * M.nI;

    double best = IS_MIN ? DBL_MAX : -DBL_MAX;
    int besti = kNoAction;
    for (int i = lane; i < nA; i += 32) {
        const int yc = min(ix + i, M.nI - 1);          // StaffRecursion.java:93-96
        const int D = alen[yc];
        const double* __restrict__ row = M.apmf_p + aoff[yc];
        const double fv = (i > 0 ? M.K : 0.0) + v * (double)i;  // fixHireCost + variHireCost
        const int yv = xv + i;
        double acc = 0.0;
        for (int j = 0; j < D; j++) {
            const int next = yv - j;                   // nextStaffNum
            const double salaryCost = M.h * (double)next;
            const double penaltyCost = next > minStaff ? 0.0 : M.pen * (double)(minStaff - next);
            const double c = (fv + salaryCost) + penaltyCost;
            const double p = __ldg(row + j);
            acc += p * c;                              // StaffRecursion.java:102
            if (!LAST) {
                int ni = next - xmin;
                ni = min(ni, M.nI - 1);                // upper clamp first (WorkforcePlanning.java:86-87)
                ni = max(ni, 0);
                acc += p * __ldg(Vn + ni);             // StaffRecursion.java:106
            }
        }
        if (IS_MIN ? (acc < best) : (acc > best)) { best = acc; besti = i; }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, best, s);
        const int oi = __shfl_xor_sync(0xffffffffu, besti, s);
        if (better<IS_MIN>(ov, oi, best, besti)) { best = ov; besti = oi; }
    }
    if (live && lane == 0) {
        Vt[idx] = best;
        Qt[idx] = besti == kNoAction ? -1 : besti;
    }
}

// Forward reachability for the staff kind (the states StaffRecursion's memoisation visits).
__global__ void __launch_bounds__(128)
reach_staff(const __grid_constant__ DevModel M, const int t, const unsigned char* __restrict__ mask_t,
            unsigned char* __restrict__ mask_n) {
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;
    if (ix >= M.nI || !mask_t[ix]) return;
    const int* __restrict__ alen = M.apmf_len + (size_t)(t - 1) * M.nI;
    for (int i = 0; i <= M.max_order_idx; i++) {
        const int D = alen[min(ix + i, M.nI - 1)];
        for (int j = 0; j < D; j++) mask_n[max(min(ix + i - j, M.nI - 1), 0)] = 1;
    }
}

}  // namespace sdpb
