// kernel_two_product.cuh — two products sharing one cash account: state (inv1, inv2, cash), action
// pairs (Q1, Q2) limited by cash (src/cash/multiItem/MultiItemCash.java:69-121 lambdas, loop
// src/sdp/cash/multiItem/CashRecursionMulti.java:81-116).  SURVEY.md §8(f) rank 3.
//
// The reference accepts an action only if it beats the incumbent by MORE than 0.1
// (`actionValues[i] > val + 0.1`, CashRecursionMulti.java:108), so the result depends on the scan
// order and is not a plain maximum: every state scans its feasible pairs serially, i-major, exactly as
// the reference's action list is built.  One thread per state; the cash axis is innermost so the lanes of
// a warp hold consecutive cash levels and gather consecutive V_{t+1} addresses.
#pragma once
#include <cuda_runtime.h>

#include "dev_model.cuh"

namespace sdpb {

template <bool LAST>
__global__ void __launch_bounds__(128)
bi_two_product(const __grid_constant__ DevModel M, const int t, const int D, const int pmf_off,
               const double* __restrict__ Vn, double* __restrict__ Vt, int* __restrict__ Qt,
               const long long lo, const long long hi) {
    const long long idx = lo + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= hi) return;
    long long r = idx;
    const int iw = (int)(r % M.nW); r /= M.nW;
    const int i2 = (int)(r % M.nI);
    const int i1 = (int)(r / M.nI);
    const double x1 = M.inv_min + (double)i1 * M.step;
    const double x2 = M.inv_min + (double)i2 * M.step;
    const double w = (double)(M.kmin + iw);  // (int) nextCash: integer cash grid
    const double v1 = M.v_t[t - 1], price1 = M.price_t[t - 1];
    const double* __restrict__ pd1 = M.pmf_d + pmf_off;
    const double* __restrict__ pd2 = M.pmf_d2 + pmf_off;
    const double* __restrict__ pp = M.pmf_p + pmf_off;
    const double* __restrict__ pg = M.pmf_pg + pmf_off;
    const int* __restrict__ di1 = M.pmf_di + pmf_off;
    const int* __restrict__ di2 = M.pmf_di2 + pmf_off;
    const int Q = M.max_order_idx + 1;

    double val = -DBL_MAX;
    int besti = -1;  // bestActions stays (0, 0)
    for (int a1i = 0; a1i < Q; a1i++) {
        const double action1 = (double)a1i;
        const double orderingCost1 = v1 * action1;
        for (int a2i = 0; a2i < Q; a2i++) {
            const double action2 = (double)a2i;
            const double orderingCost2 = M.v2 * action2;
            // MultiItemCash.java:73: affordable pairs only
            if (!(orderingCost1 + orderingCost2 < w + 0.1)) continue;
            const double orderingCosts = orderingCost1 + orderingCost2;
            const double s1 = x1 + action1, s2 = x2 + action2;
            double acc = 0.0;
            for (int j = 0; j < D; j++) {
                const double demand1 = (double)(int)__ldg(pd1 + j);  // new Demands((int) d1, (int) d2)
                const double demand2 = (double)(int)__ldg(pd2 + j);
                const double endInventory1 = fmax(0.0, s1 - demand1);
                const double endInventory2 = fmax(0.0, s2 - demand2);
                const double revenue1 = price1 * (s1 - endInventory1);
                const double revenue2 = M.price2 * (s2 - endInventory2);
                const double revenue = revenue1 + revenue2;
                double salValue = 0.0;
                if (LAST) salValue = M.salvage * endInventory1 + M.salvage2 * endInventory2;
                const double c = (revenue - orderingCosts) + salValue;
                acc += __ldg(pp + j) * c;  // CashRecursionMulti.java:101
                if (!LAST) {
                    // MultiItemCash.java:107-121: upper clamp on item 1, lower clamp on item 2 (the upper
                    // clip on item 2 is only the memory-safety bound of the dense grid)
                    int il1 = max(i1 + a1i - __ldg(di1 + j), M.i_zero);
                    int il2 = max(i2 + a2i - __ldg(di2 + j), M.i_zero);
                    il1 = max(min(il1, M.nI - 1), 0);
                    il2 = min(max(il2, 0), M.nI - 1);
                    double nw = w + c;
                    nw = nw > M.cash_max ? M.cash_max : nw;
                    nw = nw < M.cash_min ? M.cash_min : nw;
                    long long kw = (long long)nw - M.kmin;  // (int) nextCash
                    kw = kw < 0 ? 0 : (kw >= M.nW ? M.nW - 1 : kw);
                    acc += __ldg(pg + j) * __ldg(Vn + ((long long)il1 * M.nI + il2) * M.nW + kw);  // :104
                }
            }
            if (acc > val + M.tie_tol) { val = acc; besti = a1i * Q + a2i; }  // CashRecursionMulti.java:108
        }
    }
    Vt[idx] = val;
    Qt[idx] = besti;
}

// Forward reachability for the two-product kind (what CashRecursionMulti's memoisation visits).
__global__ void __launch_bounds__(128)
reach_two_product(const __grid_constant__ DevModel M, const int t, const int D, const int pmf_off,
                  const unsigned char* __restrict__ mask_t, unsigned char* __restrict__ mask_n) {
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M.S || !mask_t[idx]) return;
    long long r = idx;
    const int iw = (int)(r % M.nW); r /= M.nW;
    const int i2 = (int)(r % M.nI);
    const int i1 = (int)(r / M.nI);
    const double x1 = M.inv_min + (double)i1 * M.step, x2 = M.inv_min + (double)i2 * M.step;
    const double w = (double)(M.kmin + iw);
    const double v1 = M.v_t[t - 1], price1 = M.price_t[t - 1];
    const int Q = M.max_order_idx + 1;
    for (int a1i = 0; a1i < Q; a1i++)
        for (int a2i = 0; a2i < Q; a2i++) {
            const double oc1 = v1 * (double)a1i, oc2 = M.v2 * (double)a2i;
            if (!(oc1 + oc2 < w + 0.1)) continue;
            const double orderingCosts = oc1 + oc2;
            const double s1 = x1 + (double)a1i, s2 = x2 + (double)a2i;
            for (int j = 0; j < D; j++) {
                const double demand1 = (double)(int)M.pmf_d[pmf_off + j], demand2 = (double)(int)M.pmf_d2[pmf_off + j];
                const double e1 = fmax(0.0, s1 - demand1), e2 = fmax(0.0, s2 - demand2);
                const double revenue = price1 * (s1 - e1) + M.price2 * (s2 - e2);
                const double c = revenue - orderingCosts;  // t < T here: no salvage
                int il1 = max(i1 + a1i - M.pmf_di[pmf_off + j], M.i_zero);
                int il2 = max(i2 + a2i - M.pmf_di2[pmf_off + j], M.i_zero);
                il1 = max(min(il1, M.nI - 1), 0);
                il2 = min(max(il2, 0), M.nI - 1);
                double nw = w + c;
                nw = nw > M.cash_max ? M.cash_max : nw;
                nw = nw < M.cash_min ? M.cash_min : nw;
                long long kw = (long long)nw - M.kmin;
                kw = kw < 0 ? 0 : (kw >= M.nW ? M.nW - 1 : kw);
                mask_n[((long long)il1 * M.nI + il2) * M.nW + kw] = 1;
            }
        }
}

}  // namespace sdpb
