// dev_model.cuh — device-side form of sdpb_model and the Java-semantics helpers the kernels share.
//
// Everything here is compiled with -fmad=false: the reference is Java, which never contracts
// a*b+c into a fused multiply-add, and results must match it bit for bit
// (/root/reference/src/sdp/inventory/Recursion.java:139,142).
#pragma once
#include <cfloat>
#include <cstdint>

#include "../../include/sdpb200.h"

namespace sdpb {

// Scalars by value, tables by device pointer.  Passed to kernels as a __grid_constant__ param.
struct DevModel {
    int cost_kind, recursion, is_min, T, lead, max_order_idx;
    unsigned flags;
    int quantiser;
    int nI, nQ, nW;       // axis sizes: inventory, one pipeline slot, cash (1 when absent)
    int i_zero;           // inventory index of the value 0.0 (may lie outside [0,nI))
    long long S;          // states per period = nI * nQ^lead * nW
    long long kmin;       // integer cash index of the lowest cash grid point
    long long q_idiv;     // (long) q_div for SDPB_Q_LONGDIV
    unsigned long long q_magic;  // ceil(2^47 / q_idiv) when q_idiv < 2^15, else 0: division by multiplication
    double inv_min, step;
    double cash_min, cash_max, q_mul, q_div;
    double K, h, pen, salvage;
    double one_plus_dr;   // 1 + depositeRate        (CashConstraint.java:107)
    double one_minus_rho; // 1 - overheadRate        (CashConstraint.java:110)
    double neg_r0, r2, r3, od_limit, interest_free;
    double r2_limit_term; // r2 * (limit - interestFreeAmount)   (CashOverdraft.java:95)
    double reserve2;
    double dr;            // depositeRate itself (CashOverdraftLimit.java:79, TestPaper.java:87)
    int q_from_period;    // SDPB_Q_TRUNC: round with (q_mul, q_div) from this period on; 0 = never
    int small;            // 1: S and every quantised cash magnitude fit in 31 bits, so the successor index can be
                          //    formed in 32-bit integer arithmetic (same values, a third of the instructions)
    double price2, v2, salvage2, tie_tol;  // second product / tie tolerance (SDPB_COST_CASH_TWO_PRODUCT)
    const double* pmf_d2;
    const int* pmf_di2;
    // action-dependent pmf (SDPB_COST_STAFF): rows indexed by (t, hire-up-to level)
    const int* apmf_len;
    const int* apmf_off;
    const double* apmf_p;
    const double* min_level_t;
    // per-period parameter tables, always expanded to T entries on the host
    const double* price_t;
    const double* v_t;
    const double* ovh_t;
    const double* reserve_t;
    // demand pmf, flattened; pmf_pg[j] = p_j * gamma (same IEEE product the Java loop forms first,
    // CashRecursion.java:120 `dAndP[j][1] * discountFactor * V`), pmf_di[j] = d_j / step as an int
    const double* pmf_d;
    const double* pmf_p;
    const double* pmf_pg;
    const int* pmf_di;
    // the same four streams interleaved, two 16-byte loads per demand point from one pointer:
    // pmf_rec[2j] = (d_j, p_j), pmf_rec[2j+1] = (p_j*gamma, d_j/step in the low word)
    const double2* pmf_rec;
};

// Math.round(double) -> long (Java >= 7): floor(x + 1/2) of the EXACT sum.  With round-down addition,
// s = RD(x + 0.5) is the largest double <= x + 0.5; floor(x + 0.5) is an integer, hence representable (|x| < 2^52; above
// that x is an integer and s = x), and it is <= the exact sum, so floor(x + 0.5) <= s <= x + 0.5 and floor(s) is the
// answer: one DADD.RM and one flooring conversion, no compare.  (-2.5 -> -2, 2.5 -> 3, 0.49999999999999994 -> 0.)
__device__ __forceinline__ long long jround(double x) {
    return __double2ll_rd(__dadd_rd(x, 0.5));
}

// The same for |x| < 2^31 (DevModel::small), without an int<->double conversion instruction (F2I.F64 / I2F.F64 issue
// at a fraction of the DADD rate): RD(s + 1.5*2^52) lands in [2^52, 2^53), where doubles are the integers, so it IS
// floor(s) + 1.5*2^52 and floor(s) sits in the low mantissa word as a two's-complement int (2^51 = 0 mod 2^32).
// Two fp64 instructions per Math.round; the first version (x + magic, - magic, x - n, compare with 0.5) took four.
__device__ __forceinline__ int jround32(double x) {
    const double magic = 6755399441055744.0;  // 2^52 + 2^51
    return __double2loint(__dadd_rd(__dadd_rd(x, 0.5), magic));
}

// Java `long / long` (truncation toward zero) by the run-time constant q_idiv.  For q_idiv < 2^15 and
// |kk| < q_idiv * 2^16 (< 2^31), floor(|kk| * ceil(2^47/d) / 2^47) == floor(|kk| / d) exactly (|kk|*d < 2^47)
// and the 64-bit product cannot overflow (|kk| * magic < d 2^16 (2^47/d + 1) < 2^64), which replaces the
// ~40-instruction 64-bit division of the overdraft models' quantiser (Math.round(w*10)/10,
// CashOverdraft.java:116) by two wide multiplies and a shift.  Larger magnitudes take the real division.
__device__ __forceinline__ long long jdiv(long long kk, long long d, unsigned long long magic) {
    const long long a = kk < 0 ? -kk : kk;
    if (magic != 0 && a < (d << 16)) {
        const long long q = (long long)(((unsigned long long)a * magic) >> 47);
        return kk < 0 ? -q : q;
    }
    return kk / d;
}

__device__ __forceinline__ int jdiv32(int kk, int d, unsigned long long magic) {
    const unsigned a = (unsigned)(kk < 0 ? -kk : kk);
    if (magic != 0 && a < ((unsigned)d << 16)) {
        const int q = (int)(((unsigned long long)a * magic) >> 47);
        return kk < 0 ? -q : q;
    }
    return kk / d;
}

// Lexicographic (value, action) update used by every argopt reduction: strictly better value
// wins; on an exact tie the lower action index wins — the same result as the reference's
// ascending scan with a strict compare (Recursion.java:146-157).
template <bool IS_MIN>
__device__ __forceinline__ bool better(double v, int i, double bv, int bi) {
    return IS_MIN ? (v < bv || (v == bv && i < bi)) : (v > bv || (v == bv && i < bi));
}

constexpr int kNoAction = 0x7fffffff;

}  // namespace sdpb
