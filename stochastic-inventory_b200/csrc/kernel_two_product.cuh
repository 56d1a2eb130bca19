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

#include <cmath>

#include "dev_model.cuh"

namespace sdpb {

// LAST: period T (salvage, MultiItemCash.java:93-95).  NEXT: a continuation exists -- every period but T,
// and period T too when the model carries a terminal boundary table (CashRecursionV.java:125-128).
template <bool LAST, bool NEXT>
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
                if (NEXT) {
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

// ---------------------------------------------------------------------------------------------
// bi_two_product_row — the same scan with everything that does not depend on the cash level hoisted out
// of the thread.  A CTA is one (inv1, inv2) pair and 128 consecutive cash levels; for a fixed action pair
// the immediate value c(a, d_j), the product p_j*c and the successor's inventory row are the same for
// every cash level, so the CTA computes them ONCE per (action, demand) -- with the reference's own
// double arithmetic, one demand point per thread -- into shared memory, and a thread then spends
//     acc += (p*c)[j];  acc += (p*gamma)[j] * V_{t+1}[row[j] + clamp(iw + c[j])]
// per evaluation: 3 fp64 instructions (bi_two_product: ~20) and an integer cash index.  That last step
// needs c to be an integer (prices and unit costs integer-valued: then `(int)(w + c)` is iw + c exactly);
// plan_two_product_row checks it, otherwise bi_two_product runs.  The serial i-major scan with the
// `> val + 0.1` acceptance rule is unchanged: lanes that cannot afford a pair skip the update.
struct TwoProductRowTables {
    double2 pcg;  // (p_j * c, p_j * gamma)
    int ci, row;  // c as an integer, successor row offset (il1 * nI + il2) * nW
    int pad0, pad1;
};

template <bool LAST>
__global__ void __launch_bounds__(128)
bi_two_product_row(const __grid_constant__ DevModel M, const int t, const int D, const int pmf_off,
                   const double* __restrict__ Vn, double* __restrict__ Vt, int* __restrict__ Qt,
                   const long long lo, const long long hi, const long long pair0, const int segs) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TwoProductRowTables* TB = reinterpret_cast<TwoProductRowTables*>(smem_raw);
    const int seg = (int)(blockIdx.x % segs);
    const long long pair = pair0 + blockIdx.x / segs;  // i1 * nI + i2
    const int i2 = (int)(pair % M.nI), i1 = (int)(pair / M.nI);
    const int iw = seg * 128 + threadIdx.x;
    const long long idx = pair * M.nW + iw;
    const bool valid = iw < M.nW && idx >= lo && idx < hi;
    const double x1 = M.inv_min + (double)i1 * M.step;
    const double x2 = M.inv_min + (double)i2 * M.step;
    const double w = (double)(M.kmin + iw);
    const double w_top = (double)(M.kmin + min(seg * 128 + 127, M.nW - 1));  // richest lane of the CTA
    const double v1 = M.v_t[t - 1], price1 = M.price_t[t - 1];
    const int Q = M.max_order_idx + 1, nW1 = M.nW - 1;

    double val = -DBL_MAX;
    int besti = -1;  // bestActions stays (0, 0)
    for (int a1i = 0; a1i < Q; a1i++) {
        const double action1 = (double)a1i;
        const double orderingCost1 = v1 * action1;
        for (int a2i = 0; a2i < Q; a2i++) {
            const double action2 = (double)a2i;
            const double orderingCost2 = M.v2 * action2;
            if (!(orderingCost1 + orderingCost2 < w_top + 0.1)) continue;  // nobody here can afford it (CTA-uniform)
            const double orderingCosts = orderingCost1 + orderingCost2;
            const double s1 = x1 + action1, s2 = x2 + action2;
            __syncthreads();  // the previous pair's table is no longer read
            for (int j = threadIdx.x; j < D; j += 128) {
                const double demand1 = (double)(int)M.pmf_d[pmf_off + j];  // new Demands((int) d1, (int) d2)
                const double demand2 = (double)(int)M.pmf_d2[pmf_off + j];
                const double endInventory1 = fmax(0.0, s1 - demand1);
                const double endInventory2 = fmax(0.0, s2 - demand2);
                const double revenue1 = price1 * (s1 - endInventory1);
                const double revenue2 = M.price2 * (s2 - endInventory2);
                const double revenue = revenue1 + revenue2;
                double salValue = 0.0;
                if (LAST) salValue = M.salvage * endInventory1 + M.salvage2 * endInventory2;
                const double c = (revenue - orderingCosts) + salValue;
                TwoProductRowTables e;
                e.pcg = make_double2(M.pmf_p[pmf_off + j] * c, M.pmf_pg[pmf_off + j]);  // CashRecursionMulti.java:101
                e.ci = LAST ? 0 : (int)c;
                int il1 = max(i1 + a1i - M.pmf_di[pmf_off + j], M.i_zero);
                int il2 = max(i2 + a2i - M.pmf_di2[pmf_off + j], M.i_zero);
                il1 = max(min(il1, M.nI - 1), 0);
                il2 = min(max(il2, 0), M.nI - 1);
                e.row = (il1 * M.nI + il2) * M.nW;
                e.pad0 = e.pad1 = 0;
                TB[j] = e;
            }
            __syncthreads();
            if (!valid || !(orderingCost1 + orderingCost2 < w + 0.1)) continue;  // MultiItemCash.java:73
            double acc = 0.0;
#pragma unroll 4
            for (int j = 0; j < D; j++) {
                const double2 pcg = TB[j].pcg;
                acc += pcg.x;
                if (!LAST) {
                    const int2 cr = *reinterpret_cast<const int2*>(&TB[j].ci);
                    const int kw = max(min(iw + cr.x, nW1), 0);  // (int) nextCash on the clamped cash, integer-exact
                    acc += pcg.y * __ldg(Vn + (unsigned)(cr.y + kw));  // CashRecursionMulti.java:104
                }
            }
            if (acc > val + M.tie_tol) { val = acc; besti = a1i * Q + a2i; }  // CashRecursionMulti.java:108
        }
    }
    if (valid) {
        Vt[idx] = val;
        Qt[idx] = besti;
    }
}

// Integer-valued prices and unit costs on an integer cash grid: then c and w + c are exact integers.
inline bool plan_two_product_row(const sdpb_model& m, const DevModel& d, int t, int D) {
    auto is_whole = [](double v) { return std::isfinite(v) && v == std::floor(v) && std::fabs(v) < 1e6; };
    const double p1 = m.price_t ? m.price_t[t - 1] : m.price, v1 = m.vari_cost_t ? m.vari_cost_t[t - 1] : m.vari_cost;
    if (!is_whole(p1) || !is_whole(v1) || !is_whole(m.price2) || !is_whole(m.vari_cost2)) return false;
    if (!is_whole(m.inv_min) || m.step != 1.0) return false;
    if (m.cash_min != (double)d.kmin || m.cash_max != (double)(d.kmin + d.nW - 1)) return false;
    if ((double)d.nI * d.nI * d.nW >= 2147483647.0) return false;  // 32-bit row offsets
    const double cmax = (std::fabs(p1) + std::fabs(m.price2) + std::fabs(v1) + std::fabs(m.vari_cost2)) *
                        (std::fabs(m.inv_min) + d.nI + m.max_order_idx + 1);
    if (cmax + d.nW + std::fabs((double)d.kmin) >= 1e9) return false;
    return (size_t)D * sizeof(TwoProductRowTables) <= 96 * 1024;
}

// Forward reachability for the two-product kind (what CashRecursionMulti's memoisation visits).
__global__ void __launch_bounds__(128)
reach_two_product(const __grid_constant__ DevModel M, const int t, const int D, const int pmf_off,
                  const unsigned char* __restrict__ mask_t, unsigned char* __restrict__ mask_n,
                  unsigned long long* __restrict__ counters) {
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
    unsigned n_clip = 0, n_cash = 0;
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
                n_clip += il2 > M.nI - 1;  // item 2 has no upper clamp in MultiItemCash.java:107-121: the grid's own clip
                il1 = max(min(il1, M.nI - 1), 0);
                il2 = min(max(il2, 0), M.nI - 1);
                double nw = w + c;
                n_cash += (nw > M.cash_max || nw < M.cash_min);
                nw = nw > M.cash_max ? M.cash_max : nw;
                nw = nw < M.cash_min ? M.cash_min : nw;
                long long kw = (long long)nw - M.kmin;
                kw = kw < 0 ? 0 : (kw >= M.nW ? M.nW - 1 : kw);
                mask_n[((long long)il1 * M.nI + il2) * M.nW + kw] = 1;
            }
        }
    if (n_clip) atomicAdd(counters + 0, (unsigned long long)n_clip);
    if (n_cash) atomicAdd(counters + 2, (unsigned long long)n_cash);
}

}  // namespace sdpb
