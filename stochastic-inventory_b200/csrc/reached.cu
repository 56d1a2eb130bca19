// reached.cu — the reference's two-product recursions on the GPU, over the states they REACH.
//
//   kind 0  src/sdp/cash/multiItem/CashRecursionMultiLead.java:54-95 (loop, `> val + 0.1` acceptance rule)
//           src/cash/overdraft/MultiProductLeadtime.java:150-224 (action list, immediate value, transition)
//   kind 1  src/sdp/cash/multiItem/CashRecursionMultiXR.java:61-95, lambdas src/cash/multiItem/MultiItemCashXR.java:73-126
//           (state (x1, x2, R), actions = order-up-to pairs, `(int)` casts)
//   kind 2  src/sdp/cash/multiItem/CashRecursionV.java:83-131 (V / Pi form: no immediate value, `> val + 0.01`,
//           V_{T+1} = boundFinalCash), lambdas src/cash/multiItem/MultiItemYR.java:89-146
//   kind 3  src/sdp/cash/CashRecursion.java:79-140 with the lambdas of src/cash/singleItem/CashConstraintTest.java:76-116
//           (one product; both state components rounded as Math.round(v * 0.1) / 0.1, so neither axis has an exact step)
//
// (what follows describes kind 0, for which the engine was written; the other two plug their lambdas into it)
//
// State (period, x1, x2, preQ1, preQ2, cash): the cash balance is NOT quantised (MultiProductLeadtime.java:219 is
// commented out), so the state space is not a grid and the dense kernels of sdpb200.cu do not apply -- this is the
// model the reference's author gave up on ("3 hour no solution", MultiProductLeadtime.java:28).  The reference
// memoises top-down; here the same set of states is built bottom-up:
//
//   forward   F_1 = {initial state};  F_{t+1} = unique{ f(s, a, d) : s in F_t, a in A, d in D_t }   (expand, sort, unique)
//   backward  V_T on F_T, then V_t on F_t with V_{t+1}(f(s,a,d)) found by binary search in the sorted F_{t+1}
//
// The arithmetic of every (s, a, d) is the reference's, operation for operation and in its order (no FMA: the file is
// compiled with -fmad=false); the action scan is serial in the reference's i-major order because the acceptance rule
// `value > best + 0.1` is order dependent.  Two thread mappings: a thread per state (large frontiers: the last
// period of the T = 3 instance has 1.7e7 states x 2500 actions x 4 demand pairs) and a thread per (state, action)
// pair followed by a per-state scan (small frontiers, where a thread per state would leave the GPU empty).
#include <cuda_runtime.h>
#include <thrust/execution_policy.h>
#include <thrust/sort.h>
#include <thrust/unique.h>

#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/sdpb200.h"

namespace {

struct MLState {
    int x1, x2, q1, q2;
    double cash;
};

// sort / search key: the four integers packed, then the cash bits (any total order will do; equality is exact)
struct MLKey {
    unsigned long long a, b;
};
struct MLKeyLess {
    __host__ __device__ bool operator()(const MLKey& l, const MLKey& r) const { return l.a < r.a || (l.a == r.a && l.b < r.b); }
};
struct MLKeyEq {
    __host__ __device__ bool operator()(const MLKey& l, const MLKey& r) const { return l.a == r.a && l.b == r.b; }
};

struct MLParams {
    int kind;  // sdpb_reached_kind
    double dr; // depositeRate of the XR / YR lambdas, interestRate of CashConstraintTest
    double K, h, min_cash_req, sq;  // CASH_ROUNDED: fixed cost, holding cost, minCashRequired, rounding factor q
    int T, Q, Q2, nD;  // the action list has Q * Q2 entries (Q2 = Q for the two-product kinds, 1 for one product)
    double price1, price2, v1, v2, sal1, sal2;
    double r0, r1, r2, limit, interest_free;
    double min_inv, max_inv, min_cash, max_cash, gamma, tie;
    const double* d1;   // [T * nD] demand of product 1 (cast to int as the reference does)
    const double* d2;
    const double* p;    // [T * nD]
    const double* pg;   // p * gamma
    const double* ovh;  // [T]
    const int* nDt;     // [T] demand points of each period (<= nD, the row stride)
};

__host__ __device__ inline MLKey ml_key(const MLState& s) {
    MLKey k;
    k.a = ((unsigned long long)(unsigned)s.x1 << 48) | ((unsigned long long)(unsigned)s.x2 << 32) |
          ((unsigned long long)(unsigned)s.q1 << 16) | (unsigned long long)(unsigned)s.q2;
    const double c = s.cash + 0.0;  // -0.0 and +0.0 are one key in the reference's map (compared with ==)
    unsigned long long bits;
    memcpy(&bits, &c, sizeof bits);
    k.b = bits;
    return k;
}

__host__ __device__ inline MLState ml_state(const MLKey& k) {
    MLState s;
    s.x1 = (int)((k.a >> 48) & 0xffff); s.x2 = (int)((k.a >> 32) & 0xffff);
    s.q1 = (int)((k.a >> 16) & 0xffff); s.q2 = (int)(k.a & 0xffff);
    memcpy(&s.cash, &k.b, sizeof s.cash);
    return s;
}

__device__ __forceinline__ double jmax0(double x) { return x > 0.0 ? x : 0.0; }  // Math.max(0, x) for non-NaN x
__device__ __forceinline__ double jmin(double a, double b) { return a < b ? a : b; }

// MultiProductLeadtime.java:162-198
__device__ __forceinline__ double ml_immediate(const MLParams& P, int t, bool last, const MLState& s, int a1, int a2, int d1,
                                               int d2) {
    const double action1 = (double)a1, action2 = (double)a2, demand1 = (double)d1, demand2 = (double)d2;
    const double stock1 = (double)s.x1 + (double)s.q1, stock2 = (double)s.x2 + (double)s.q2;
    const double endInventory1 = jmax0(stock1 - demand1);
    const double endInventory2 = jmax0(stock2 - demand2);
    const double revenue1 = P.price1 * jmin(demand1, stock1);
    const double revenue2 = P.price2 * jmin(stock2, demand2);
    const double revenue = revenue1 + revenue2;
    const double orderingCosts = P.v1 * action1 + P.v2 * action2;
    double salValue = 0.0;
    if (last) salValue = P.sal1 * endInventory1 + P.sal2 * endInventory2;
    const double before = (s.cash - orderingCosts) - P.ovh[t - 1];
    double interest;
    if (before >= 0.0) interest = -P.r0 * before;
    else if (before >= -P.interest_free) interest = 0.0;
    else if (before >= -P.limit) interest = P.r1 * (-before - P.interest_free);
    else interest = P.r2 * (-before - P.limit) + P.r1 * (P.limit - P.interest_free);
    const double after = ((before - interest) + revenue) + salValue;
    return after - s.cash;
}

// MultiProductLeadtime.java:202-224 (the asymmetric clamps are the reference's)
__device__ __forceinline__ MLState ml_transition(const MLParams& P, const MLState& s, int a1, int a2, int d1, int d2, double c) {
    double e1 = jmax0(((double)s.x1 + (double)s.q1) - (double)d1);
    double e2 = jmax0(((double)s.x2 + (double)s.q2) - (double)d2);
    double nextCash = s.cash + c;
    nextCash = nextCash > P.max_cash ? P.max_cash : nextCash;
    nextCash = nextCash < P.min_cash ? P.min_cash : nextCash;
    e1 = e1 > P.max_inv ? P.max_inv : e1;
    e2 = e2 < P.min_inv ? P.min_inv : e2;
    MLState n;
    n.x1 = (int)e1; n.x2 = (int)e2; n.q1 = a1; n.q2 = a2; n.cash = nextCash;
    return n;
}

__device__ __forceinline__ long long ml_find(const MLKey* __restrict__ F, long long n, const MLKey& k) {
    long long lo = 0, hi = n;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        const MLKey m = F[mid];
        if (m.a < k.a || (m.a == k.a && m.b < k.b)) lo = mid + 1; else hi = mid;
    }
    return lo;  // the key is present by construction
}

// ---- kind 1: MultiItemCashXR.java:87-126 ----
__device__ __forceinline__ double xr_immediate(const MLParams& P, bool last, const MLState& s, double action1, double action2,
                                               double demand1, double demand2) {
    const double endInventory1 = jmax0(action1 - demand1);
    const double endInventory2 = jmax0(action2 - demand2);
    const double revenue1 = P.price1 * (action1 - endInventory1);
    const double revenue2 = P.price2 * (action2 - endInventory2);
    const double revenue = revenue1 + revenue2;
    const double initialCash = (s.cash - P.v1 * (double)s.x1) - P.v2 * (double)s.x2;
    const double orderingCostsY = P.v1 * action1 + P.v2 * action2;
    double salValue = 0.0;
    if (last) salValue = P.sal1 * endInventory1 + P.sal2 * endInventory2;
    return ((revenue + (1.0 - P.dr) * (s.cash - orderingCostsY)) + salValue) - initialCash;
}

__device__ __forceinline__ MLState xr_transition(const MLParams& P, const MLState& s, double a1, double a2, double d1, double d2,
                                                 double c) {
    double e1 = jmax0(a1 - d1), e2 = jmax0(a2 - d2);
    const double initialCash = (s.cash - P.v1 * (double)s.x1) - P.v2 * (double)s.x2;
    double nextCash = initialCash + c;
    nextCash = nextCash > P.max_cash ? P.max_cash : nextCash;
    nextCash = nextCash < P.min_cash ? P.min_cash : nextCash;
    e1 = e1 > P.max_inv ? P.max_inv : e1;
    e2 = e2 < P.min_inv ? P.min_inv : e2;
    nextCash = (double)(int)nextCash;
    MLState n;
    n.x1 = (int)e1; n.x2 = (int)e2; n.q1 = 0; n.q2 = 0;
    n.cash = (nextCash + P.v1 * (double)n.x1) + P.v2 * (double)n.x2;  // nextR
    return n;
}

// ---- kind 2: MultiItemYR.java:122-146 (Math.round(x * 10) / 10 is a LONG division) ----
__device__ __forceinline__ long long ml_jround(double x) {
    return __double2ll_rd(__dadd_rd(x, 0.5));  // floor of the exact x + 1/2: see jround() in dev_model.cuh
}

__device__ __forceinline__ MLState yr_transition(const MLParams& P, double y1, double y2, double iniR, double d1, double d2) {
    double e1 = jmax0(y1 - d1), e2 = jmax0(y2 - d2);
    const double revenue1 = P.price1 * jmin(y1, d1);
    const double revenue2 = P.price2 * jmin(y2, d2);
    double nextW = (revenue1 + revenue2) + (1.0 + P.dr) * ((iniR - P.v1 * y1) - P.v2 * y2);
    e1 = (double)(ml_jround(e1 * 10.0) / 10);
    e2 = (double)(ml_jround(e2 * 10.0) / 10);
    nextW = (double)(ml_jround(nextW * 10.0) / 10);
    nextW = nextW > P.max_cash ? P.max_cash : nextW;
    nextW = nextW < P.min_cash ? P.min_cash : nextW;
    e1 = e1 > P.max_inv ? P.max_inv : e1;
    e2 = e2 < P.min_inv ? P.min_inv : e2;
    MLState n;
    n.x1 = (int)e1; n.x2 = (int)e2; n.q1 = 0; n.q2 = 0; n.cash = nextW;
    return n;
}

// ---- kind 3: CashConstraintTest.java:76-116 (one product; x = k / q with k kept in the key) ----
__device__ __forceinline__ double cct_immediate(const MLParams& P, bool last, double x, double w, double action, double d) {
    const double revenue = P.price1 * jmin(x + action, d);
    const double fixedCost = action > 0.0 ? P.K : 0.0;
    const double variableCost = P.v1 * action;
    const double inventoryLevel = (x + action) - d;
    const double holdCosts = P.h * jmax0(inventoryLevel);
    const double interests = P.dr * (w - action * P.v1);
    double cashIncrement = (((revenue - fixedCost) - variableCost) - holdCosts) + interests;
    const double salValue = last ? P.sal1 * jmax0(inventoryLevel) : 0.0;
    cashIncrement += salValue;
    return cashIncrement;
}

__device__ __forceinline__ MLState cct_transition(const MLParams& P, double x, double w, double action, double d, double c) {
    double nextInventory = jmax0((x + action) - d);
    double nextCash = w + c;
    nextCash = nextCash > P.max_cash ? P.max_cash : nextCash;
    nextCash = nextCash < P.min_cash ? P.min_cash : nextCash;
    nextInventory = nextInventory > P.max_inv ? P.max_inv : nextInventory;
    nextInventory = nextInventory < P.min_inv ? P.min_inv : nextInventory;
    MLState n;
    n.cash = (double)ml_jround(nextCash * P.sq) / P.sq;   // Math.round(nextCash * 0.1) / 0.1
    n.x1 = (int)ml_jround(nextInventory * P.sq);           // the inventory is kept as k; its value is k / q
    n.x2 = 0; n.q1 = 0; n.q2 = 0;
    return n;
}

// the action with flat index ai of state s: its two components as the lambdas see them, and whether the reference's
// action list contains it (kind 2: `v1 * i + v2 * j < iniR + 0.1`, MultiItemYR.java:104-108)
__device__ __forceinline__ bool ml_action(const MLParams& P, const MLState& s, int ai, double& a1, double& a2) {
    const int i = ai / P.Q2, j = ai - i * P.Q2;
    if (P.kind == 0) { a1 = (double)i; a2 = (double)j; return true; }
    if (P.kind == 3) {  // CashConstraintTest.java:76-80
        a1 = (double)i; a2 = 0.0;
        const double maxQ = (double)(int)fmin((double)(P.Q - 1), fmax(0.0, ((s.cash - P.min_cash_req) - P.K) / P.v1));
        return i < (int)maxQ + 1;
    }
    a1 = (double)(s.x1 + i); a2 = (double)(s.x2 + j);  // order-up-to levels from (int) x
    if (P.kind == 1) return true;
    const double iniR = (s.cash + P.v1 * (double)s.x1) + P.v2 * (double)s.x2;
    return P.v1 * a1 + P.v2 * a2 < iniR + 0.1;
}

__device__ __forceinline__ MLState ml_successor(const MLParams& P, int t, const MLState& s, double a1, double a2, int j,
                                                double* c_out) {
    const double d1 = P.d1[(t - 1) * P.nD + j], d2 = P.d2[(t - 1) * P.nD + j];
    if (P.kind == 0) {
        const double c = ml_immediate(P, t, false, s, (int)a1, (int)a2, (int)d1, (int)d2);
        if (c_out) *c_out = c;
        return ml_transition(P, s, (int)a1, (int)a2, (int)d1, (int)d2, c);
    }
    if (P.kind == 1) {
        const double c = xr_immediate(P, false, s, a1, a2, d1, d2);
        if (c_out) *c_out = c;
        return xr_transition(P, s, a1, a2, d1, d2, c);
    }
    if (P.kind == 3) {
        const double x = (double)s.x1 / P.sq;
        const double c = cct_immediate(P, false, x, s.cash, a1, d1);
        if (c_out) *c_out = c;
        return cct_transition(P, x, s.cash, a1, d1, c);
    }
    const double iniR = (s.cash + P.v1 * (double)s.x1) + P.v2 * (double)s.x2;
    return yr_transition(P, a1, a2, iniR, d1, d2);
}

// F_{t+1} candidates: one entry per (state, action, demand)
__global__ void ml_expand(MLParams P, int t, const MLKey* __restrict__ F, long long nF, MLKey* __restrict__ out) {
    const long long A = (long long)P.Q * P.Q2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nF * A) return;
    const long long si = gid / A;
    const int ai = (int)(gid - si * A);
    const MLState s = ml_state(F[si]);
    double a1, a2;
    const bool feasible = ml_action(P, s, ai, a1, a2);
    // an action the reference's list does not contain reaches nothing: its slots repeat the state's first successor
    // under a listed action... there may be none, so they are filled with an all-ones key that unique() folds into
    // one entry and the caller drops
    const int nDt = P.nDt[t - 1];
    for (int j = 0; j < P.nD; j++) {
        MLKey k;
        if (feasible && j < nDt) k = ml_key(ml_successor(P, t, s, a1, a2, j, nullptr));
        else { k.a = ~0ull; k.b = ~0ull; }
        out[gid * P.nD + j] = k;
    }
}

// Q(s, a): kinds 0, 1: sum_j [ p_j c + (p_j gamma) V_{t+1}(f) ] (CashRecursionMultiLead.java:74-80,
// CashRecursionMultiXR.java:78-85); kind 2: sum_j p_j V_{t+1}(f) with V_{T+1} = boundFinalCash (CashRecursionV.java:88-97,125-128)
__device__ __forceinline__ double ml_action_value(const MLParams& P, int t, bool last, const MLState& s, double a1, double a2,
                                                  const MLKey* __restrict__ Fn, const double* __restrict__ Vn, long long nFn) {
    double q = 0.0;
    const int nDt = P.nDt[t - 1];
    for (int j = 0; j < nDt; j++) {
        const double pj = P.p[(t - 1) * P.nD + j];
        if (P.kind == 2) {
            const MLState n = ml_successor(P, t, s, a1, a2, j, nullptr);
            const double vn = last ? (n.cash + P.sal1 * (double)n.x1) + P.sal2 * (double)n.x2   // MultiItemYR.java:116-119
                                   : Vn[ml_find(Fn, nFn, ml_key(n))];
            q += pj * vn;
            continue;
        }
        const double d1 = P.d1[(t - 1) * P.nD + j], d2 = P.d2[(t - 1) * P.nD + j];
        const double x3 = (double)s.x1 / P.sq;  // kind 3: the inventory value
        const double c = P.kind == 0 ? ml_immediate(P, t, last, s, (int)a1, (int)a2, (int)d1, (int)d2)
                       : P.kind == 1 ? xr_immediate(P, last, s, a1, a2, d1, d2)
                                     : cct_immediate(P, last, x3, s.cash, a1, d1);
        q += pj * c;
        if (!last) {
            const MLState n = P.kind == 0 ? ml_transition(P, s, (int)a1, (int)a2, (int)d1, (int)d2, c)
                            : P.kind == 1 ? xr_transition(P, s, a1, a2, d1, d2, c)
                                          : cct_transition(P, x3, s.cash, a1, d1, c);
            q += P.pg[(t - 1) * P.nD + j] * Vn[ml_find(Fn, nFn, ml_key(n))];
        }
    }
    return q;
}

// a thread per state, serial i-major action scan with the `> val + tie` rule (CashRecursionMultiLead.java:81-85)
__global__ void ml_backward_state(MLParams P, int t, const MLKey* __restrict__ F, long long nF, const MLKey* __restrict__ Fn,
                                  const double* __restrict__ Vn, long long nFn, double* __restrict__ V, int* __restrict__ Qa) {
    const long long si = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nF) return;
    const MLState s = ml_state(F[si]);
    const bool last = t == P.T;
    double val = -DBL_MAX;
    int best = 0;  // bestActions = new Actions(0, 0); kind 2: bestYs = (x1, x2) = offsets (0, 0) as well
    const int A = P.Q * P.Q2;
    for (int ai = 0; ai < A; ai++) {
        double a1, a2;
        if (!ml_action(P, s, ai, a1, a2)) continue;
        const double q = ml_action_value(P, t, last, s, a1, a2, Fn, Vn, nFn);
        if (q > val + P.tie) { val = q; best = ai; }
    }
    V[si] = val;
    Qa[si] = best;
}

// small frontiers: a thread per (state, action) pair ...
__global__ void ml_backward_pairs(MLParams P, int t, const MLKey* __restrict__ F, long long nF, const MLKey* __restrict__ Fn,
                                  const double* __restrict__ Vn, long long nFn, double* __restrict__ QV) {
    const long long A = (long long)P.Q * P.Q2;
    const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= nF * A) return;
    const long long si = gid / A;
    const int ai = (int)(gid - si * A);
    const MLState s = ml_state(F[si]);
    double a1, a2;
    // an action that is not in the reference's list can never be accepted: NaN fails `q > val + tie`
    QV[gid] = ml_action(P, s, ai, a1, a2) ? ml_action_value(P, t, t == P.T, s, a1, a2, Fn, Vn, nFn) : __longlong_as_double(0x7ff8000000000000ll);
}

// ... then the serial scan per state
__global__ void ml_scan_pairs(MLParams P, const double* __restrict__ QV, long long nF, double* __restrict__ V, int* __restrict__ Qa) {
    const long long si = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (si >= nF) return;
    const long long A = (long long)P.Q * P.Q2;
    double val = -DBL_MAX;
    int best = 0;
    for (long long ai = 0; ai < A; ai++) {
        const double q = QV[si * A + ai];
        if (q > val + P.tie) { val = q; best = (int)ai; }
    }
    V[si] = val;
    Qa[si] = best;
}

thread_local std::string g_ml_error;

}  // namespace

extern "C" {

const char* sdpb_reached_last_error(void) { return g_ml_error.c_str(); }

int sdpb_reached_solve(const sdpb_reached_model* m, int device, const double* init_state, double* value, double* action1,
                       double* action2, int64_t* n_states, double* solve_ms) {
    g_ml_error.clear();
    auto fail = [&](int rc, const std::string& msg) { g_ml_error = msg; return rc; };
    if (!m || !init_state || m->struct_size != sizeof(sdpb_reached_model)) return fail(SDPB_ERR_ARG, "bad argument / struct_size");
    if (m->kind < SDPB_REACHED_MULTILEAD || m->kind > SDPB_REACHED_CASH_ROUNDED) return fail(SDPB_ERR_ARG, "bad kind");
    if (m->kind == SDPB_REACHED_CASH_ROUNDED && !(m->state_q > 0.0 && m->vari_cost[0] > 0.0)) return fail(SDPB_ERR_ARG, "CASH_ROUNDED needs state_q > 0 and a positive unit cost");
    if (m->T < 1 || m->q_bound < 1 || m->q_bound > 65535 || m->n_demands < 1 || !m->d1 || !m->d2 || !m->p || !m->overhead_t)
        return fail(SDPB_ERR_ARG, "T, q_bound, n_demands must be positive and the tables non-null");
    // init_state: kind 0 (x1, x2, preQ1, preQ2, cash); kind 1 (x1, x2, R); kind 2 (x1, x2, cash)
    const int n_int = m->kind == SDPB_REACHED_MULTILEAD ? 4 : (m->kind == SDPB_REACHED_CASH_ROUNDED ? 1 : 2);
    double init_k[4] = {0, 0, 0, 0};
    for (int k = 0; k < n_int; k++) {
        init_k[k] = init_state[k];
        if (m->kind == SDPB_REACHED_CASH_ROUNDED) {  // the inventory is kept as k = round(x * q); x must be k / q exactly
            init_k[k] = std::floor(init_state[k] * m->state_q + 0.5);
            if (init_k[k] / m->state_q != init_state[k]) return fail(SDPB_ERR_OFFGRID, "the initial inventory is not of the form k / state_q");
        }
        if (init_k[k] != (double)(int)init_k[k] || init_k[k] < 0 || init_k[k] > 65535)
            return fail(SDPB_ERR_OFFGRID, "initial inventories and pipeline quantities must be integers in [0, 65535]");
    }
    if (!(m->max_inv <= 65535.0)) return fail(SDPB_ERR_ARG, "max_inv must fit 16 bits");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(SDPB_ERR_NO_DEVICE, "no CUDA device (libsdpb200 has no CPU path)"); }
    if (device < 0) cudaGetDevice(&device);
    if (device >= ndev || cudaSetDevice(device) != cudaSuccess) return fail(SDPB_ERR_NO_DEVICE, "cannot select the CUDA device");

    const int T = m->T, nD = m->n_demands;
    const int q2 = m->kind == SDPB_REACHED_CASH_ROUNDED ? 1 : m->q_bound;
    const long long A = (long long)m->q_bound * q2;
    std::vector<double> pg((size_t)T * nD);
    for (size_t i = 0; i < pg.size(); i++) pg[i] = m->p[i] * m->gamma;  // first product of p * gamma * V
    std::vector<void*> owned;
    cudaStream_t stream = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto cleanup = [&]() {
        for (void* p : owned) cudaFree(p);
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        if (stream) cudaStreamDestroy(stream);
        cudaGetLastError();
    };
    auto dalloc = [&](size_t bytes) -> void* {
        void* p = nullptr;
        if (cudaMalloc(&p, bytes > 0 ? bytes : 16) != cudaSuccess) { cudaGetLastError(); return nullptr; }
        owned.push_back(p);
        return p;
    };
#define ML_CU(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { const std::string msg_ = std::string(#call) + ": " + cudaGetErrorString(e_); cleanup(); return fail(SDPB_ERR_CUDA, msg_); } } while (0)
    ML_CU(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
    ML_CU(cudaEventCreate(&e0));
    ML_CU(cudaEventCreate(&e1));
    auto upload = [&](const double* src, size_t n) -> double* {
        double* p = (double*)dalloc(n * sizeof(double));
        if (p && cudaMemcpyAsync(p, src, n * sizeof(double), cudaMemcpyHostToDevice, stream) != cudaSuccess) return nullptr;
        return p;
    };
    MLParams P;
    P.kind = m->kind; P.dr = m->deposit_rate;
    P.K = m->fixed_cost; P.h = m->hold_cost; P.min_cash_req = m->min_cash_required; P.sq = m->state_q > 0.0 ? m->state_q : 1.0;
    P.T = T; P.Q = m->q_bound; P.Q2 = q2; P.nD = nD;
    P.price1 = m->price[0]; P.price2 = m->price[1]; P.v1 = m->vari_cost[0]; P.v2 = m->vari_cost[1];
    P.sal1 = m->salvage[0]; P.sal2 = m->salvage[1];
    P.r0 = m->r0; P.r1 = m->r1; P.r2 = m->r2; P.limit = m->limit; P.interest_free = m->interest_free;
    P.min_inv = m->min_inv; P.max_inv = m->max_inv; P.min_cash = m->min_cash; P.max_cash = m->max_cash;
    P.gamma = m->gamma; P.tie = m->tie_tolerance;
    P.d1 = upload(m->d1, (size_t)T * nD); P.d2 = upload(m->d2, (size_t)T * nD); P.p = upload(m->p, (size_t)T * nD);
    P.pg = upload(pg.data(), pg.size()); P.ovh = upload(m->overhead_t, (size_t)T);
    {
        std::vector<int> lens(T, nD);
        if (m->n_demands_t)
            for (int t = 0; t < T; t++) {
                if (m->n_demands_t[t] < 1 || m->n_demands_t[t] > nD) { cleanup(); return fail(SDPB_ERR_ARG, "n_demands_t out of range"); }
                lens[t] = m->n_demands_t[t];
            }
        int* dl = (int*)dalloc((size_t)T * sizeof(int));
        if (dl && cudaMemcpyAsync(dl, lens.data(), (size_t)T * sizeof(int), cudaMemcpyHostToDevice, stream) != cudaSuccess) dl = nullptr;
        if (dl) cudaStreamSynchronize(stream);  // (lens is a local)
        P.nDt = dl;
    }
    if (!P.d1 || !P.d2 || !P.p || !P.pg || !P.ovh || !P.nDt) { cleanup(); return fail(SDPB_ERR_NOMEM, "allocation of the model tables failed"); }
    ML_CU(cudaEventRecord(e0, stream));

    // ---- forward: the states the reference's recursion visits, period by period ----
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    std::vector<MLKey*> F(T, nullptr);
    std::vector<long long> nF(T, 0);
    MLState s0;
    s0.x1 = (int)init_k[0]; s0.x2 = n_int >= 2 ? (int)init_k[1] : 0;
    s0.q1 = n_int == 4 ? (int)init_k[2] : 0; s0.q2 = n_int == 4 ? (int)init_k[3] : 0;
    s0.cash = init_state[n_int];
    const MLKey k0 = ml_key(s0);
    F[0] = (MLKey*)dalloc(sizeof(MLKey));
    if (!F[0]) { cleanup(); return fail(SDPB_ERR_NOMEM, "allocation failed"); }
    ML_CU(cudaMemcpyAsync(F[0], &k0, sizeof k0, cudaMemcpyHostToDevice, stream));
    nF[0] = 1;
    for (int t = 1; t < T; t++) {
        const long long cand = nF[t - 1] * A * nD;
        if ((double)cand * sizeof(MLKey) * 2.2 > (double)free_b) {
            cleanup();
            return fail(SDPB_ERR_NOMEM, "period " + std::to_string(t + 1) + " has " + std::to_string(cand) +
                                            " candidate states: more than this GPU's memory holds");
        }
        MLKey* buf = (MLKey*)dalloc((size_t)cand * sizeof(MLKey));
        if (!buf) { cleanup(); return fail(SDPB_ERR_NOMEM, "allocation of the candidate states failed"); }
        const long long pairs = nF[t - 1] * A;
        ml_expand<<<(unsigned)((pairs + 127) / 128), 128, 0, stream>>>(P, t, F[t - 1], nF[t - 1], buf);
        ML_CU(cudaGetLastError());
        try {
            thrust::sort(thrust::cuda::par.on(stream), buf, buf + cand, MLKeyLess());
            MLKey* end = thrust::unique(thrust::cuda::par.on(stream), buf, buf + cand, MLKeyEq());
            nF[t] = (long long)(end - buf);
            if (nF[t] > 0) {  // the slots of actions the reference's list does not contain sort last, as one key
                MLKey tail;
                if (cudaMemcpyAsync(&tail, buf + nF[t] - 1, sizeof tail, cudaMemcpyDeviceToHost, stream) == cudaSuccess &&
                    cudaStreamSynchronize(stream) == cudaSuccess && tail.a == ~0ull && tail.b == ~0ull)
                    nF[t]--;
            }
        } catch (const std::exception& ex) {
            const std::string msg = std::string("sort / unique: ") + ex.what();
            cleanup();
            return fail(SDPB_ERR_CUDA, msg);
        }
        F[t] = buf;
        cudaMemGetInfo(&free_b, &total_b);
    }

    // ---- backward ----
    std::vector<double*> V(T, nullptr);
    std::vector<int*> Qa(T, nullptr);
    for (int t = T; t >= 1; t--) {
        const long long n = nF[t - 1];
        V[t - 1] = (double*)dalloc((size_t)n * sizeof(double));
        Qa[t - 1] = (int*)dalloc((size_t)n * sizeof(int));
        if (!V[t - 1] || !Qa[t - 1]) { cleanup(); return fail(SDPB_ERR_NOMEM, "allocation of the value table failed"); }
        const MLKey* Fn = t < T ? F[t] : nullptr;
        const double* Vn = t < T ? V[t] : nullptr;
        const long long nFn = t < T ? nF[t] : 0;
        if (n * A <= 200000000LL && n < 50000) {  // few states: spread their actions over the GPU
            double* QV = (double*)dalloc((size_t)(n * A) * sizeof(double));
            if (!QV) { cleanup(); return fail(SDPB_ERR_NOMEM, "allocation of the action-value scratch failed"); }
            ml_backward_pairs<<<(unsigned)((n * A + 127) / 128), 128, 0, stream>>>(P, t, F[t - 1], n, Fn, Vn, nFn, QV);
            ml_scan_pairs<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(P, QV, n, V[t - 1], Qa[t - 1]);
        } else {
            ml_backward_state<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(P, t, F[t - 1], n, Fn, Vn, nFn, V[t - 1], Qa[t - 1]);
        }
        ML_CU(cudaGetLastError());
    }
    ML_CU(cudaEventRecord(e1, stream));
    double v0 = 0;
    int a0 = 0;
    ML_CU(cudaMemcpyAsync(&v0, V[0], sizeof v0, cudaMemcpyDeviceToHost, stream));
    ML_CU(cudaMemcpyAsync(&a0, Qa[0], sizeof a0, cudaMemcpyDeviceToHost, stream));
    ML_CU(cudaStreamSynchronize(stream));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
#undef ML_CU
    if (value) *value = v0;
    // kind 0: order quantities (i, j); kinds 1, 2: order-up-to levels (x1 + i, x2 + j) (bestYs; the reference's default
    // for a state without any accepted action is {0, 0} for kind 1 and (x1, x2) for kind 2)
    const int ai = a0 / q2, aj = a0 % q2;
    const bool none = v0 == -DBL_MAX;
    const bool qty = m->kind == SDPB_REACHED_MULTILEAD || m->kind == SDPB_REACHED_CASH_ROUNDED;  // order quantities
    if (action1) *action1 = qty ? ai : (m->kind == SDPB_REACHED_MULTI_XR && none ? 0.0 : s0.x1 + ai);
    if (action2) *action2 = qty ? aj : (m->kind == SDPB_REACHED_MULTI_XR && none ? 0.0 : s0.x2 + aj);
    if (n_states) for (int t = 0; t < T; t++) n_states[t] = nF[t];
    if (solve_ms) *solve_ms = ms;
    cleanup();
    return SDPB_OK;
}

}  // extern "C"
