// kernel_fused.cuh — the whole horizon of a SMALL 1-D inventory model in ONE cooperative launch.
//
// Configs C1 / C2 (1,001 and 2,001 states) are the sizes the reference's own drivers run, and hundreds of
// them (CLSPTesting.java:57-61).  With one launch per period such a solve is a chain of T short, under-filled
// kernels: ~10 us each, of which ~3 us is arithmetic (profiles/r01_launches_c2.csv) -- the rest is launch
// ramp, the level-window fill and the merge of the action split through global memory.  Here the grid is
// one CTA per SM for the whole solve: CTA b owns the states [b*BX, (b+1)*BX) for every period and ALL their
// actions, so the argopt never leaves the CTA; periods are separated by a grid-wide barrier (cooperative
// groups) instead of a kernel boundary, and V_t goes through L2 only.
//
// Per period a CTA stages W[il] = (h*l+ + pi*l-, V_{t+1}[succ(il)]) for the levels its states can reach
// (as bi_inv_tiled does) and the demand table; a thread then owns (state, action) pairs -- consecutive
// threads = consecutive states, so the 16-byte window reads of a warp are consecutive -- and spends
//     c = fv_a + W.x;  acc += p*c;  acc += (p*gamma)*W.y          (5 fp64 instructions, 1 LDS.128)
// per evaluation, 8 pairs at a time for instruction-level parallelism.  Q-values go to shared memory and one
// warp per state takes the lexicographic (value, action) optimum == the reference's ascending first-wins
// scan (Recursion.java:146-157).  Same arithmetic, same order as bi_inv_tiled / bi_generic: bit-identical.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <vector>

#include "dev_model.cuh"

namespace sdpb {

namespace cg = cooperative_groups;

constexpr int kFusedThreads = 512;
constexpr int kFusedBatch = 8;   // pairs per thread in flight
constexpr int kFusedMaxBX = 64;  // states per CTA

struct FusedArgs {
    int T, A, BX, WN, Dmax;
    const int* len;       // [T] demand points per period
    const int* off;       // [T] offset into the flattened pmf
    const int* di_max;    // [T] largest demand index of the period
    double* const* V;     // [T] value tables
    int* const* Q;        // [T] policy tables
};

// NK (state, action) pairs of one thread, p0 + k*NT: all demand points, Q-values to shared memory.
template <int NK>
__device__ __forceinline__ void fused_pairs(const DevModel& M, const double2* __restrict__ W, const double2* __restrict__ PP,
                                            const int* __restrict__ E, double* __restrict__ QV, int D, int P, int nx,
                                            int p0, int NT, double v, bool last) {
    double fv[NK], acc[NK];
    int iy[NK];
#pragma unroll
    for (int k = 0; k < NK; k++) {
        const int p = min(p0 + k * NT, P - 1);  // (clamped: surplus lanes of the last round recompute the last pair)
        const int ai = p / nx, s = p - ai * nx;
        const double av = (double)ai * M.step;
        fv[k] = (av > 0.0 ? M.K : 0.0) + v * av;  // fixedCost + variableCost (CLSPTesting.java:98-99,104)
        iy[k] = s + ai;
        acc[k] = 0.0;
    }
    if (last) {
        for (int j = 0; j < D; j++) {
            const double2 pp = PP[j];
            const int e = E[j];
#pragma unroll
            for (int k = 0; k < NK; k++) acc[k] += pp.x * (fv[k] + W[iy[k] + e].x);  // Recursion.java:139
        }
    } else {
        for (int j = 0; j < D; j++) {
            const double2 pp = PP[j];
            const int e = E[j];
#pragma unroll
            for (int k = 0; k < NK; k++) {
                const double2 w = W[iy[k] + e];
                acc[k] += pp.x * (fv[k] + w.x);  // Recursion.java:139
                acc[k] += pp.y * w.y;            // Recursion.java:142
            }
        }
    }
#pragma unroll
    for (int k = 0; k < NK; k++)
        if (p0 + k * NT < P) QV[p0 + k * NT] = acc[k];  // pair index = a * nx + s
}

template <bool IS_MIN>
__global__ void __launch_bounds__(kFusedThreads, 1)
bi_inv_fused(const __grid_constant__ DevModel M, const __grid_constant__ FusedArgs a) {
    constexpr int NT = kFusedThreads, B = kFusedBatch;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double2* W = reinterpret_cast<double2*>(smem_raw);
    double2* PP = W + a.WN;
    double* QV = reinterpret_cast<double*>(PP + a.Dmax);            // [A][nx] Q-values of the CTA's pairs
    int* E = reinterpret_cast<int*>(QV + (size_t)a.A * a.BX);       // di_max - di_j

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int A = a.A;
    const long long x0 = (long long)blockIdx.x * a.BX;
    const int nx = (int)max(0LL, min((long long)a.BX, M.S - x0));
    const int P = nx * A;
    const bool lost = (M.flags & SDPB_F_LOST_SALES) != 0;

    for (int t = a.T; t >= 1; t--) {
        const bool last = (t == a.T);
        if (nx > 0) {
            const int D = a.len[t - 1], poff = a.off[t - 1], dmax = a.di_max[t - 1];
            const double* Vn = last ? nullptr : a.V[t];
            for (int j = tid; j < D; j += NT) {
                PP[j] = make_double2(M.pmf_p[poff + j], M.pmf_pg[poff + j]);
                E[j] = dmax - M.pmf_di[poff + j];
            }
            // window index wi <-> level index il = x0 - dmax + wi; the pairs reach wi < nx + A - 1 + span + 1 <= WN
            for (int wi = tid; wi < a.WN; wi += NT) {
                const long long il = x0 - dmax + wi;
                const double lvl = M.inv_min + (double)il * M.step;  // exact on the validated grid
                const double hold = M.h * fmax(lvl, 0.0);
                const double pen = M.pen * fmax(-lvl, 0.0);
                double vn = 0.0;
                if (!last) {
                    long long is = il;
                    if (lost) is = is > M.i_zero ? is : M.i_zero;
                    is = is < M.nI - 1 ? is : M.nI - 1;  // upper clamp first (CLSPTesting.java:91-92)
                    is = is > 0 ? is : 0;
                    vn = __ldcg(Vn + is);                // written earlier in THIS launch by another CTA: L2, not the nc path
                }
                W[wi] = make_double2(hold + pen, vn);
            }
            __syncthreads();

            const double v = M.v_t[t - 1];
            for (int base = 0; base < P; base += NT * B) {
                // pairs base + tid + k*NT, k < nk: nk is CTA-uniform, so only whole rounds of pairs are issued
                const int nk = min(B, (P - base + NT - 1) / NT);
                switch (nk) {
                case 1: fused_pairs<1>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 2: fused_pairs<2>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 3: fused_pairs<3>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 4: fused_pairs<4>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 5: fused_pairs<5>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 6: fused_pairs<6>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                case 7: fused_pairs<7>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                default: fused_pairs<8>(M, W, PP, E, QV, D, P, nx, base + tid, NT, v, last); break;
                }
            }
            __syncthreads();
            // ---- argopt: one warp per state, lanes over ascending actions ----
            for (int s = warp; s < nx; s += NT / 32) {
                double best = IS_MIN ? DBL_MAX : -DBL_MAX;
                int besti = kNoAction;
                for (int i = lane; i < A; i += 32) {
                    const double q = QV[i * nx + s];
                    if (IS_MIN ? (q < best) : (q > best)) { best = q; besti = i; }
                }
#pragma unroll
                for (int sh = 16; sh > 0; sh >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, best, sh);
                    const int oi = __shfl_xor_sync(0xffffffffu, besti, sh);
                    if (better<IS_MIN>(ov, oi, best, besti)) { best = ov; besti = oi; }
                }
                if (lane == 0) {
                    a.V[t - 1][x0 + s] = best;
                    a.Q[t - 1][x0 + s] = besti == kNoAction ? -1 : besti;
                }
            }
        }
        if (t > 1) {
            __threadfence();
            grid.sync();  // V_t is complete before anyone stages it for period t-1
        }
    }
}

struct FusedPlan {
    bool ok = false;
    int BX = 0, grid = 0, WN = 0, Dmax = 0;
    size_t smem = 0;
    int *d_len = nullptr, *d_off = nullptr, *d_dimax = nullptr;
    double** d_V = nullptr;
    int** d_Q = nullptr;
    std::vector<int> dimax;
};

// Small unsharded lead-0 backorder grids only: every CTA must be co-resident (cooperative launch).
inline void plan_fused(FusedPlan& P, const sdpb_model& m, const DevModel& d, const std::vector<int>& pmf_len,
                       const std::vector<int>& pmf_off, const std::vector<int>& pdi, int sm_count, int shard_count) {
    P.ok = false;
    if (m.cost_kind != SDPB_COST_BACKORDER || m.lead_time != 0 || shard_count != 1) return;
    if (!(m.flags & SDPB_F_CLAMP_INV) || (m.flags & (SDPB_F_GY_MODE | SDPB_F_NO_ORDER_LAST))) return;
    if (m.terminal_value) return;  // period T would need the boundary table: per-period kernels handle it
    const long long S = d.S;
    const int A = d.max_order_idx + 1;
    const int G = (int)std::min<long long>(sm_count, S);
    P.BX = (int)((S + G - 1) / G);
    if (P.BX > kFusedMaxBX) return;
    P.grid = (int)((S + P.BX - 1) / P.BX);
    int span = 0;
    P.Dmax = 0;
    P.dimax.assign(m.T, 0);
    for (int t = 0; t < m.T; t++) {
        const int* di = pdi.data() + pmf_off[t];
        int lo = di[0], hi = di[0];
        for (int j = 0; j < pmf_len[t]; j++) { lo = std::min(lo, di[j]); hi = std::max(hi, di[j]); }
        span = std::max(span, hi - lo);
        P.Dmax = std::max(P.Dmax, pmf_len[t]);
        P.dimax[t] = hi;
    }
    P.WN = P.BX + A + span;
    P.smem = (size_t)(P.WN + P.Dmax) * 16 + (size_t)A * P.BX * 8 + (size_t)P.Dmax * 4 + 16;
    if (P.smem > 200 * 1024) return;
    P.ok = true;
}

}  // namespace sdpb
