// sdp_oracle.cpp — CPU restatement of the reference's SDP recursion.  TEST INFRASTRUCTURE ONLY.
//
// This file is the parity oracle for libsdpb200.so.  It is loaded by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs and by nothing
// else; the product library never links or calls it.
//
// HOW THIS ORACLE IS PINNED.  The reference (Java 21 + SSJ 3.3.0) cannot run in the build container (no
// JDK, no jar) and its repository holds no golden vectors, fixtures or known-answer tests for this path
// (SURVEY.md §4, §8c).  The only program outputs the reference records are in the header comment of
// src/cash/overdraft/MultiProductLeadtime.java:30-50 (a two-product model built on a copy of the same
// loop).  `oracle_multi_lead` below drives THIS file's solve_state / TopDown with that model's lambdas
// and reproduces two of those outputs to the last digit:
//     T = 2, demands {20,30,40}x{10,15,20}:  -17.800000000000008, Q1 = 40, Q2 = 20   (tests/test_oracle.py)
//     T = 3, demands {10,30}x{5,15}:         -76.56,              Q1 = 30, Q2 = 15   (34 CPU-min, opt-in)
// which pins the loop order (s += p*c; s += p*gamma*V), the init / strict-compare / first-wins rule,
// the memo-map control flow, the four-branch overdraft interest and the clamp / (int) idioms against
// the real Java program.  PARITY UNPINNED for the rest: the single-product lambdas of each family and
// the SSJ-built pmf tables have no recorded reference output; they are held in place by
//   * being written line-by-line after the Java sources cited at each function;
//   * two independent evaluation orders — `oracle_topdown` (memoised recursion over real-valued states
//     in ordered maps, exactly the reference's control flow) and `oracle_dense` (period-by-period over
//     the full grid, states addressed by index) — agreeing bit-for-bit on every visited state;
//   * a pure-Python transliteration of Recursion.java + the family-A and family-C lambdas, closed-form
//     cases, and the survey's independent scratch value for C1 (tests/test_oracle.py).
//
// Arithmetic: IEEE-754 double, operations in the order the Java source writes them, compiled
// with -ffp-contract=off (Java never fuses a multiply-add).
//
// The reference loop (src/sdp/inventory/Recursion.java:129-161, repeated in
// src/sdp/cash/CashRecursion.java:98-138, src/sdp/inventory/LeadtimeRecursion.java:49-73,
// src/sdp/cash/CashLeadtimeRecursion.java:50-77, src/sdp/cash/CashRecursionXR.java:82-124,
// src/capacitated/CLSP.java:111-136):
//
//     val = MIN ? Double.MAX_VALUE : -Double.MAX_VALUE;  bestOrderQty = 0;
//     for i over feasibleActions:  thisQ = 0
//         for j over pmf[t-1]:     thisQ += p_j * c(s, a_i, d_j)
//                                  if (t < T) thisQ += p_j * gamma * V(f(s, a_i, d_j))
//         strict compare -> (val, bestOrderQty)
//
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include <atomic>
#include <thread>

#include "../include/sdpb200.h"

namespace {

// ---- Java semantics -------------------------------------------------------------------------
// Math.round(double): round half up to long (exact; Java >= 7 does not compute x + 0.5).
inline int64_t jround(double x) {
    double r = std::floor(x);
    double diff = x - r;  // exact for |x| < 2^52
    return (int64_t)r + (diff >= 0.5 ? 1 : 0);
}
// Math.max / Math.min on doubles (no NaNs on this path); +0.0 > -0.0 as in Java.
inline double jmax(double a, double b) {
    if (a == 0.0 && b == 0.0) return (std::signbit(a) && std::signbit(b)) ? -0.0 : 0.0;
    return a >= b ? a : b;
}
inline double jmin(double a, double b) {
    if (a == 0.0 && b == 0.0) return (std::signbit(a) || std::signbit(b)) ? -0.0 : 0.0;
    return a <= b ? a : b;
}

struct St {  // State / LeadtimeState / CashState / CashLeadtimeState / RiskState / CashStateXR
    int t;   // 1-based period
    double x = 0, w = 0, q1 = 0, q2 = 0;  // w holds R for the XR kind
    double x2 = 0;                         // second product's inventory (two-product known-answer model only)
};

// Parameters of the reference's two-product cash + lead-time model (src/cash/overdraft/
// MultiProductLeadtime.java:86-118), used only to pin this file's loop against the program output the
// reference author recorded in that file's header comment.
struct MultiLead {
    int Qbound = 50;
    double price[2] = {5, 10}, variCost[2] = {1, 2}, salValueUnit[2] = {0.5, 1.0};
    double r0 = 0, r1 = 0.1, r2 = 2, limit = 500, interestFreeAmount = 0;
    double minInventoryState = 0, maxInventoryState = 200, minCashState = -500, maxCashState = 5000;
    std::vector<double> overheadCost;
    std::vector<int> d1, d2;  // demand pairs, same table every period (GetPmfMulti.java:158-171)
};

struct Model {
    sdpb_model m;
    const MultiLead* multi = nullptr;
    double tie_tol = 0.0;  // CashRecursionMultiLead.java:82 accepts an action only if it is better by > 0.1
    std::vector<int> off;  // pmf offsets
    int ndemand(int t) const { return m.pmf_len[t - 1]; }
    double d(int t, int j) const { return two() ? (double)j : m.pmf_d[off[t - 1] + j]; }  // two-product: row index
    double d1(int t, int j) const { return m.pmf_d[off[t - 1] + j]; }
    double d2(int t, int j) const { return m.pmf_d2[off[t - 1] + j]; }
    bool two() const { return m.cost_kind == SDPB_COST_CASH_TWO_PRODUCT; }
    bool staff() const { return m.cost_kind == SDPB_COST_STAFF; }
    std::vector<int64_t> aoff;  // staff: offsets of the (t, y) rows of apmf_p
    double p(int t, int j) const { return m.pmf_p[off[t - 1] + j]; }
    bool cash() const { return m.cost_kind != SDPB_COST_BACKORDER && m.cost_kind != SDPB_COST_STAFF; }
    bool flag(uint32_t f) const { return (m.flags & f) != 0; }
    double price(int t) const { return m.price_t ? m.price_t[t - 1] : m.price; }
    double vcost(int t) const { return m.vari_cost_t ? m.vari_cost_t[t - 1] : m.vari_cost; }
    double ovh(int t) const { return m.overhead_t ? m.overhead_t[t - 1] : m.overhead; }
    double reserve(int t) const { return m.reserve_t ? m.reserve_t[t - 1] : 0.0; }
    int n_inv() const { return (int)jround((m.inv_max - m.inv_min) / m.step) + 1; }
    int n_q() const { return m.lead_time > 0 ? m.max_order_idx + 1 : 1; }
    // integer cash index of a cash value that is already on the quantised grid
    int64_t cash_k(double w) const {
        if (m.cost_kind == SDPB_COST_CASH_XR) return jround(w);
        return m.quantiser == SDPB_Q_DIV ? jround(w * m.q_div) : (int64_t)w;  // LONGDIV, TRUNC: integers
    }
    double cash_of_k(int64_t k) const {
        if (m.cost_kind == SDPB_COST_CASH_XR) return (double)k;
        return m.quantiser == SDPB_Q_DIV ? (double)k / m.q_div : (double)k;
    }
    // `t` is the period of the state being left (TestPaper.java:107 rounds only when t > 2)
    double quantise(double w, int t = 1 << 30) const {
        if (m.quantiser == SDPB_Q_TRUNC) {  // TestPaper.java:107-109
            if (m.q_from_period > 0 && t >= m.q_from_period) w = (double)jround(w * m.q_mul) / m.q_div;
            return (double)(int)w;
        }
        int64_t kk = jround(w * m.q_mul);
        if (m.quantiser == SDPB_Q_DIV) return (double)kk / m.q_div;
        return (double)(kk / (int64_t)m.q_div);  // Java long division truncates toward zero
    }
    double trunc_only(double w) const { return m.quantiser == SDPB_Q_TRUNC ? (double)(int)w : quantise(w); }
    int64_t k_min() const {
        if (m.cost_kind == SDPB_COST_CASH_XR)
            return jround(quantise(m.cash_min) + m.vari_cost * m.inv_min);
        return cash_k(trunc_only(m.cash_min));
    }
    int64_t k_max() const {
        if (m.cost_kind == SDPB_COST_CASH_XR)
            return jround(quantise(m.cash_max) + m.vari_cost * m.inv_max);
        return cash_k(trunc_only(m.cash_max));
    }
    int n_cash() const { return cash() ? (int)(k_max() - k_min() + 1) : 1; }
    int64_t n_states() const {
        int64_t n = n_inv();
        if (two()) n *= n_inv();
        for (int l = 0; l < m.lead_time; l++) n *= n_q();
        return n * n_cash();
    }
};

// ---- feasible actions -----------------------------------------------------------------------
// CLSPTesting.java:78-86 (0..maxQ), CashConstraint.java:95-100 and cashSurvival.java:102-110
// (cash-limited), SingleProductLeadtime.java:72-77 ({0} in the last period),
// CashConstraintXR.java:71-75 (order-up-to levels from x).
int n_actions(const Model& M, const St& s) {
    const sdpb_model& m = M.m;
    if (M.multi) return M.multi->Qbound * M.multi->Qbound;  // MultiProductLeadtime.java:150-158
    if (M.two()) return (m.max_order_idx + 1) * (m.max_order_idx + 1);  // scanned i-major; see action_feasible
    if (m.cost_kind == SDPB_COST_CASH_XR) {
        double v = M.vcost(s.t);
        double maxY = s.w / v < s.x ? s.x : s.w / v;
        int length = (int)(maxY - s.x) + 1;
        return std::min(length, m.max_order_idx + 1);  // dense-grid cap, see DESIGN.md
    }
    double maxQ = (double)m.max_order_idx;
    if (M.flag(SDPB_F_CASH_LIMITED_ACTIONS))
        maxQ = (int)std::min((double)m.max_order_idx,
                             std::max(0.0, (s.w - M.reserve(s.t) - m.reserve2) / M.vcost(s.t)));
    if (M.flag(SDPB_F_NO_ORDER_LAST) && s.t == m.T) maxQ = 0;
    return (int)maxQ + 1;
}
// MultiItemCash.java:72-76: the pair (i, j) is in the action list only if it is affordable
inline bool action_feasible(const Model& M, const St& s, int i) {
    if (!M.two()) return true;
    const int Q = M.m.max_order_idx + 1;
    return M.m.vari_cost * (i / Q) + M.m.vari_cost2 * (i % Q) < s.w + 0.1;
}
inline double action_value(const Model& M, const St& s, int i) {
    if (M.multi || M.two()) return (double)i;  // flat index of the pair (i / Qbound, i % Qbound), i-major as in the driver
    if (M.m.cost_kind == SDPB_COST_CASH_XR) return s.x + i * M.m.step;
    return i * M.m.step;
}

// ---- the two-product model (known-answer pin only) -----------------------------------------------
// MultiProductLeadtime.java:162-198
double multi_immediate(const Model& M, const St& s, int a1, int a2, int d1, int d2) {
    const MultiLead& p = *M.multi;
    double action1 = a1, action2 = a2, demand1 = d1, demand2 = d2;
    double preQ1 = s.q1, preQ2 = s.q2;
    double endInventory1 = jmax(0, s.x + preQ1 - demand1);
    double endInventory2 = jmax(0, s.x2 + preQ2 - demand2);
    double revenue1 = p.price[0] * jmin(demand1, s.x + preQ1);
    double revenue2 = p.price[1] * jmin(s.x2 + preQ2, demand2);
    double revenue = revenue1 + revenue2;
    double orderingCost1 = p.variCost[0] * action1;
    double orderingCost2 = p.variCost[1] * action2;
    double orderingCosts = orderingCost1 + orderingCost2;
    double salValue = 0;
    if (s.t == M.m.T) salValue = p.salValueUnit[0] * endInventory1 + p.salValueUnit[1] * endInventory2;
    int t = s.t - 1;
    double cashBalanceBefore = s.w - orderingCosts - p.overheadCost[t];
    double interest = 0;
    if (cashBalanceBefore >= 0)
        interest = -p.r0 * cashBalanceBefore;
    else if (cashBalanceBefore >= -p.interestFreeAmount)
        interest = 0;
    else if (cashBalanceBefore >= -p.limit)
        interest = p.r1 * (-cashBalanceBefore - p.interestFreeAmount);
    else
        interest = p.r2 * (-cashBalanceBefore - p.limit) + p.r1 * (p.limit - p.interestFreeAmount);
    double cashBalanceAfter = cashBalanceBefore - interest + revenue + salValue;
    double cashIncrement = cashBalanceAfter - s.w;
    return cashIncrement;
}

// MultiProductLeadtime.java:202-224 (asymmetric clamps as written there)
St multi_transition(const Model& M, const St& s, int a1, int a2, int d1, int d2) {
    const MultiLead& p = *M.multi;
    double endInventory1 = s.x + s.q1 - d1;
    endInventory1 = jmax(0, endInventory1);
    double endInventory2 = s.x2 + s.q2 - d2;
    endInventory2 = jmax(0, endInventory2);
    double nextCash = s.w + multi_immediate(M, s, a1, a2, d1, d2);
    nextCash = nextCash > p.maxCashState ? p.maxCashState : nextCash;
    nextCash = nextCash < p.minCashState ? p.minCashState : nextCash;
    endInventory1 = endInventory1 > p.maxInventoryState ? p.maxInventoryState : endInventory1;
    endInventory2 = endInventory2 < p.minInventoryState ? p.minInventoryState : endInventory2;
    endInventory1 = (int)endInventory1;
    endInventory2 = (int)endInventory2;
    St n;
    n.t = s.t + 1; n.x = endInventory1; n.x2 = endInventory2; n.q1 = a1; n.q2 = a2; n.w = nextCash;
    return n;
}

// ---- immediate value ------------------------------------------------------------------------
double immediate(const Model& M, const St& s, double action, double demand) {
    const sdpb_model& m = M.m;
    const int T = m.T;
    if (M.multi) {  // action = flat action index, demand = flat demand index
        const int i = (int)action, j = (int)demand, Q = M.multi->Qbound;
        return multi_immediate(M, s, i / Q, i % Q, M.multi->d1[j], M.multi->d2[j]);
    }
    if (M.two()) {
        // MultiItemCash.java:82-103 (action = flat pair index, demand = pmf row index)
        const int Q = m.max_order_idx + 1, fi = (int)action, j = (int)demand;
        double action1 = fi / Q, action2 = fi % Q;
        double demand1 = (int)M.d1(s.t, j), demand2 = (int)M.d2(s.t, j);
        double endInventory1 = jmax(0, s.x + action1 - demand1);
        double endInventory2 = jmax(0, s.x2 + action2 - demand2);
        double revenue1 = m.price * (s.x + action1 - endInventory1);
        double revenue2 = m.price2 * (s.x2 + action2 - endInventory2);
        double revenue = revenue1 + revenue2;
        double orderingCost1 = m.vari_cost * action1;
        double orderingCost2 = m.vari_cost2 * action2;
        double orderingCosts = orderingCost1 + orderingCost2;
        double salValue = 0;
        if (s.t == T) salValue = m.salvage * endInventory1 + m.salvage2 * endInventory2;
        return revenue - orderingCosts + salValue;
    }
    switch (m.cost_kind) {
    case SDPB_COST_BACKORDER: {
        // CLSPTesting.java:96-106; lead time: Leadtime.java:71-81; G(y) pass: CLSPforDraw.java:156-170
        double fixedCost, variableCost, inventoryLevel;
        if (M.flag(SDPB_F_GY_MODE) && s.t == 1) {
            fixedCost = 0;
            variableCost = M.vcost(s.t) * s.x;
            inventoryLevel = s.x - demand;
        } else {
            fixedCost = action > 0 ? m.fixed_cost : 0;
            variableCost = M.vcost(s.t) * action;
            inventoryLevel = (m.lead_time > 0 ? s.x + s.q1 : s.x + action) - demand;
        }
        double holdingCosts = m.hold_cost * jmax(inventoryLevel, 0);
        double penaltyCosts = m.penalty_cost * jmax(-inventoryLevel, 0);
        return fixedCost + variableCost + holdingCosts + penaltyCosts;
    }
    case SDPB_COST_CASH_DEPOSIT: {
        // CashConstraint.java:103-121, cashSurvival.java:116-129
        double stock = m.lead_time > 0 ? s.x + s.q1 : s.x + action;
        double revenue = M.price(s.t) * jmin(stock, demand);
        double fixedCost = action > 0 ? m.fixed_cost : 0;
        double variableCost = M.vcost(s.t) * action;
        double deposite = (s.w - fixedCost - variableCost) * (1 + m.deposit_rate);
        double inventoryLevel = stock - demand;
        double holdCosts = m.hold_cost * jmax(inventoryLevel, 0);
        double cashIncrement =
            (1 - m.overhead_rate) * revenue + deposite - holdCosts - M.ovh(s.t) - s.w;
        double salValue = s.t == T ? m.salvage * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        double endCash = s.w + cashIncrement;
        if (endCash < 0) cashIncrement += m.penalty_cost * endCash;
        return cashIncrement;
    }
    case SDPB_COST_CASH_OVERDRAFT: {
        // CashOverdraft.java:80-104, SingleProductLeadtime.java:82-103
        double stock = m.lead_time > 0 ? s.x + s.q1 : s.x + action;
        double revenue = M.price(s.t) * jmin(stock, demand);
        double fixedCost = action > 0 ? m.fixed_cost : 0;
        double variableCost = M.vcost(s.t) * action;
        double inventoryLevel = stock - demand;
        double cashBalanceBefore = s.w - fixedCost - variableCost - M.ovh(s.t);
        double interest = 0;
        if (cashBalanceBefore >= 0)
            interest = -m.r0 * cashBalanceBefore;
        else if (cashBalanceBefore >= -m.interest_free)
            interest = 0;
        else if (cashBalanceBefore >= -m.od_limit)
            interest = m.r2 * (-cashBalanceBefore - m.interest_free);
        else
            interest = m.r3 * (-cashBalanceBefore - m.od_limit) + m.r2 * (m.od_limit - m.interest_free);
        double cashBalanceAfter = cashBalanceBefore - interest + revenue;
        double cashIncrement = cashBalanceAfter - s.w;
        double salValue = s.t == T ? m.salvage * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        return cashIncrement;
    }
    case SDPB_COST_CASH_OD_LIMIT: {
        // CashOverdraftLimit.java:70-86
        double revenue = M.price(s.t) * jmin(s.x + action, demand);
        double fixedCost = action > 0 ? m.fixed_cost : 0;
        double variableCost = M.vcost(s.t) * action;
        double inventoryLevel = s.x + action - demand;
        double holdCosts = m.hold_cost * jmax(inventoryLevel, 0);
        double cashBalanceBeforeRevenue = s.w - fixedCost - variableCost - holdCosts - M.ovh(s.t);
        double interest = m.r2 * jmax(-cashBalanceBeforeRevenue, 0);
        double deposite = m.deposit_rate * jmax(cashBalanceBeforeRevenue, 0);
        double cashBalanceAfter = cashBalanceBeforeRevenue - interest + deposite + revenue;
        double cashIncrement = cashBalanceAfter - s.w;
        double salValue = s.t == T ? m.salvage * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        return cashIncrement;
    }
    case SDPB_COST_CASH_OD_TESTING: {
        // CashOverdraftTesting.java:85-99
        double revenue = M.price(s.t) * jmin(s.x + action, demand);
        double fixedCost = action > 0 ? m.fixed_cost : 0;
        double variableCost = M.vcost(s.t) * action;
        double inventoryLevel = s.x + action - demand;
        double holdCosts = m.hold_cost * jmax(inventoryLevel, 0);
        double cashBalanceBefore = s.w + revenue - fixedCost - variableCost - holdCosts;
        double interest = m.r2 * jmax(-cashBalanceBefore, 0);
        double cashBalanceAfter = cashBalanceBefore - interest;
        double cashIncrement = cashBalanceAfter - s.w;
        return cashIncrement;
    }
    case SDPB_COST_CASH_LOAN: {
        // TestPaper.java:82-93
        double revenue = M.price(s.t) * jmin(s.x + action, demand);
        double variableCost = M.vcost(s.t) * action;
        double inventoryLevel = s.x + action - demand;
        double holdCosts = s.t == T ? 0 : m.hold_cost * jmax(inventoryLevel, 0);
        double deposites = m.deposit_rate * jmax(s.w - variableCost, 0);
        double loanPayed = m.r2 * jmax(variableCost - s.w, 0);
        double cashIncrement = revenue - variableCost - holdCosts + deposites - loanPayed;
        double salValue = s.t == T ? m.salvage * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        return cashIncrement;
    }
    case SDPB_COST_STAFF: {
        // WorkforcePlanning.java:92-101 (all-integer state, action, turnover)
        int act = (int)action, randomDemand = (int)demand, iniStaffNum = (int)s.x;
        double fixHireCost = act > 0 ? m.fixed_cost : 0;
        double variHireCost = M.vcost(s.t) * act;
        int nextStaffNum = iniStaffNum + act - randomDemand;
        double salaryCost = m.hold_cost * nextStaffNum;
        int minStaff = (int)m.min_level_t[s.t - 1];
        double penaltyCost = nextStaffNum > minStaff ? 0 : m.penalty_cost * (minStaff - nextStaffNum);
        return fixHireCost + variHireCost + salaryCost + penaltyCost;
    }
    case SDPB_COST_CASH_XR: {
        // CashConstraintXR.java:78-92 (action is the order-up-to level y; s.w holds R)
        double v = M.vcost(s.t);
        double actionY = action;
        double revenue = M.price(s.t) * jmin(actionY, demand);
        double act = actionY - s.x;
        double fixedCost = actionY > s.x ? m.fixed_cost : 0;
        double variableCost = v * act;
        double initCash = s.w - v * s.x;
        double deposite = (initCash - fixedCost - variableCost) * (1 + m.deposit_rate);
        double inventoryLevel = actionY - demand;
        double holdCosts = m.hold_cost * jmax(inventoryLevel, 0);
        double cashIncrement =
            (1 - m.overhead_rate) * revenue + deposite - holdCosts - M.ovh(s.t) - initCash;
        double salValue = s.t == T ? m.salvage * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        return cashIncrement;
    }
    }
    return 0;
}

// ---- state transition -----------------------------------------------------------------------
St transition(const Model& M, const St& s, double action, double demand) {
    const sdpb_model& m = M.m;
    if (M.multi) {
        const int i = (int)action, j = (int)demand, Q = M.multi->Qbound;
        return multi_transition(M, s, i / Q, i % Q, M.multi->d1[j], M.multi->d2[j]);
    }
    if (M.staff()) {
        // WorkforcePlanning.java:84-89
        int nextStaffNum = (int)s.x + (int)action - (int)demand;
        nextStaffNum = nextStaffNum > (int)m.inv_max ? (int)m.inv_max : nextStaffNum;
        nextStaffNum = nextStaffNum < (int)m.inv_min ? (int)m.inv_min : nextStaffNum;
        St n3; n3.t = s.t + 1; n3.x = nextStaffNum;
        return n3;
    }
    if (M.two()) {
        // MultiItemCash.java:107-121 (upper clamp on item 1, lower clamp on item 2, as written there)
        const int Q = m.max_order_idx + 1, fi = (int)action, j = (int)demand;
        double endInventory1 = s.x + (double)(fi / Q) - (double)(int)M.d1(s.t, j);
        endInventory1 = jmax(0, endInventory1);
        double endInventory2 = s.x2 + (double)(fi % Q) - (double)(int)M.d2(s.t, j);
        endInventory2 = jmax(0, endInventory2);
        double nextCash = s.w + immediate(M, s, action, demand);
        nextCash = nextCash > m.cash_max ? m.cash_max : nextCash;
        nextCash = nextCash < m.cash_min ? m.cash_min : nextCash;
        endInventory1 = endInventory1 > m.inv_max ? m.inv_max : endInventory1;
        endInventory2 = endInventory2 < m.inv_min ? m.inv_min : endInventory2;
        nextCash = (int)nextCash;
        endInventory1 = (int)endInventory1;
        endInventory2 = (int)endInventory2;
        St n2;
        n2.t = s.t + 1; n2.x = endInventory1; n2.x2 = endInventory2; n2.w = nextCash;
        return n2;
    }
    St n;
    n.t = s.t + 1;
    if (m.cost_kind == SDPB_COST_BACKORDER) {
        // CLSPTesting.java:89-94, Leadtime.java:61-68, CLSPforDraw.java:147-153
        double nextInventory;
        if (M.flag(SDPB_F_GY_MODE) && s.t == 1)
            nextInventory = s.x - demand;
        else
            nextInventory = (m.lead_time > 0 ? s.x + s.q1 : s.x + action) - demand;
        if (M.flag(SDPB_F_LOST_SALES)) nextInventory = jmax(0, nextInventory);
        if (M.flag(SDPB_F_CLAMP_INV)) {
            nextInventory = nextInventory > m.inv_max ? m.inv_max : nextInventory;
            nextInventory = nextInventory < m.inv_min ? m.inv_min : nextInventory;
        }
        n.x = nextInventory;
    } else if (m.cost_kind == SDPB_COST_CASH_XR) {
        // CashConstraintXR.java:95-110
        double v = M.vcost(s.t);
        double nextInventory = jmax(0, action - demand);
        double initCash = s.w - v * s.x;
        double nextCash = initCash + immediate(M, s, action, demand);
        nextCash = nextCash > m.cash_max ? m.cash_max : nextCash;
        nextCash = nextCash < m.cash_min ? m.cash_min : nextCash;
        nextInventory = nextInventory > m.inv_max ? m.inv_max : nextInventory;
        nextInventory = nextInventory < m.inv_min ? m.inv_min : nextInventory;
        nextCash = M.quantise(nextCash);
        n.x = nextInventory;
        n.w = nextCash + v * nextInventory;  // nextR
        return n;
    } else {
        // CashConstraint.java:123-133, CashOverdraft.java:107-118, SingleProductLeadtime.java:106-119,
        // cashSurvival.java:132-147
        double stock = m.lead_time > 0 ? s.x + s.q1 : s.x + action;
        double nextInventory = stock - demand;
        if (M.flag(SDPB_F_LOST_SALES)) nextInventory = jmax(0, nextInventory);
        double nextCash = s.w + immediate(M, s, action, demand);
        if (m.cost_kind == SDPB_COST_CASH_OD_TESTING) {
            // CashOverdraftTesting.java:103-111: the transition recomputes the balance itself
            double revenue = M.price(s.t) * jmin(s.x + action, demand);
            double fixedCost = action > 0 ? m.fixed_cost : 0;
            double variableCost = M.vcost(s.t) * action;
            double holdCosts = m.hold_cost * jmax(nextInventory, 0);
            nextCash = s.w + revenue - fixedCost - variableCost - holdCosts;
            nextCash = nextCash - jmax(-nextCash, 0) * m.r2;
        }
        nextCash = nextCash > m.cash_max ? m.cash_max : nextCash;
        nextCash = nextCash < m.cash_min ? m.cash_min : nextCash;
        if (M.flag(SDPB_F_CLAMP_INV)) {
            nextInventory = nextInventory > m.inv_max ? m.inv_max : nextInventory;
            nextInventory = nextInventory < m.inv_min ? m.inv_min : nextInventory;
        }
        nextCash = M.quantise(nextCash, s.t);
        n.x = nextInventory;
        n.w = nextCash;
    }
    if (m.lead_time == 1) {
        n.q1 = action;
    } else if (m.lead_time == 2) {
        n.q1 = s.q2;
        n.q2 = action;
    }
    return n;
}

// ---- one state: the loop of Recursion.java:129-161 / RiskRecursion.java:66-104 ----------------
template <class NextValue>
void solve_state(const Model& M, const St& s, NextValue&& next_value, double* val_out,
                 double* best_out, double* evals) {
    const sdpb_model& m = M.m;
    const int T = m.T;
    const bool survival = m.recursion == SDPB_REC_SURVIVAL;
    const bool is_min = !survival && m.direction == SDPB_MIN;
    int nA = n_actions(M, s);
    int D = M.ndemand(s.t);
    double val = is_min ? DBL_MAX : -DBL_MAX;
    double bestOrderQty = 0;
    int nFeasible = 0;
    double nEvalsStaff = 0;
    for (int i = 0; i < nA; i++) {
        if (!action_feasible(M, s, i)) continue;
        nFeasible++;
        double orderQty = action_value(M, s, i);
        double thisQValue = 0;
        // StaffRecursion.java:93-97: the pmf row is chosen by the hire-up-to level, capped at the last row
        const double* arow = nullptr;
        if (M.staff()) {
            int hireUpTo = (int)jround((s.x - m.inv_min) / m.step) + i;
            if (hireUpTo >= M.n_inv() - 1) hireUpTo = M.n_inv() - 1;
            D = m.apmf_len[(s.t - 1) * M.n_inv() + hireUpTo];
            arow = m.apmf_p + M.aoff[(size_t)(s.t - 1) * M.n_inv() + hireUpTo];
            nEvalsStaff += D;
        }
        for (int j = 0; j < D; j++) {
            double randomDemand = M.staff() ? (double)j : M.d(s.t, j);
            double dProb = M.staff() ? arow[j] : M.p(s.t, j);
            if (!survival) {
                thisQValue += dProb * immediate(M, s, orderQty, randomDemand);
                // Recursion.java:140 `if (n < T)`; with a boundary function the recursion goes one period further and
                // the state of period T+1 is worth boundFinalCash.apply(s) (CashRecursionV.java:125-128)
                if (s.t < T || m.terminal_value) {
                    St ns = transition(M, s, orderQty, randomDemand);
                    thisQValue += dProb * m.gamma * next_value(ns);
                }
            } else {
                if (s.t == T) {
                    double thisDFinalCash = s.w + immediate(M, s, orderQty, randomDemand);
                    double thisDProb = thisDFinalCash >= 0 ? 1 : 0;
                    thisQValue += dProb * thisDProb;
                }
                if (s.t < T) {
                    St ns = transition(M, s, orderQty, randomDemand);
                    double thisDProb = 0;
                    if (ns.w < 0)
                        thisDProb = 0;
                    else
                        thisDProb = next_value(ns);
                    thisQValue += dProb * m.gamma * thisDProb;
                }
            }
        }
        // strict compare (Recursion.java:147,153); M.tie_tol is 0 except for the two-product pin, where
        // the reference writes `> val + 0.1` (CashRecursionMultiLead.java:82).  val +/- 0.0 == val exactly.
        if (is_min) {
            if (thisQValue < val - M.tie_tol) { val = thisQValue; bestOrderQty = orderQty; }
        } else {
            if (thisQValue > val + M.tie_tol) { val = thisQValue; bestOrderQty = orderQty; }
        }
    }
    if (evals) *evals += M.staff() ? nEvalsStaff : (double)nFeasible * D;
    *val_out = val;
    *best_out = bestOrderQty;
}

// ---- key order of the memo maps ---------------------------------------------------------------
// Recursion.java:58-60 (period, inv); LeadtimeRecursion.java:37-40 (period, inv, preQ);
// CashRecursion.java:51-54 (period, inv, cash).  CashLeadtimeRecursion.java:37-41 is an
// inconsistent comparator in the reference; (period, inv, preQ, cash) is used here, which
// changes memo hits only, not values.
struct KeyLess {
    bool operator()(const St& a, const St& b) const {
        if (a.t != b.t) return a.t < b.t;
        if (a.x != b.x) return a.x < b.x;
        if (a.x2 != b.x2) return a.x2 < b.x2;  // CashRecursionMultiLead.java:45-51 (0 otherwise)
        if (a.q1 != b.q1) return a.q1 < b.q1;
        if (a.q2 != b.q2) return a.q2 < b.q2;
        return a.w < b.w;
    }
};

struct TopDown {
    const Model& M;
    std::map<St, double, KeyLess> cacheValues, cacheActions;
    double evals = 0;
    std::function<double(const St&)> boundary;  // FinalCash.BoundaryFuncton: value of a state of period T+1
    explicit TopDown(const Model& m) : M(m) {}
    double getExpectedValue(const St& s) {
        if (s.t > M.m.T) return boundary(s);  // CashRecursionV.java:125-128
        auto it = cacheValues.find(s);
        if (it != cacheValues.end()) return it->second;
        double val, best;
        solve_state(M, s, [this](const St& ns) { return getExpectedValue(ns); }, &val, &best, &evals);
        cacheValues.emplace(s, val);
        cacheActions.emplace(s, best);
        return val;
    }
};

// ---- dense grid -------------------------------------------------------------------------------
struct Grid {
    const Model& M;
    int nI, nQ, nW, L;
    int64_t kmin, S;
    explicit Grid(const Model& m) : M(m) {
        nI = M.n_inv(); nQ = M.n_q(); nW = M.n_cash(); L = M.m.lead_time;
        kmin = M.cash() ? M.k_min() : 0;
        S = M.n_states();
    }
    // library state order: inventory outermost, then the order pipeline, cash innermost
    int64_t index(const St& s, bool* off) const {
        int64_t ix = jround((s.x - M.m.inv_min) / M.m.step);
        if (ix < 0) { ix = 0; *off = true; }
        if (ix >= nI) { ix = nI - 1; *off = true; }
        int64_t idx = ix;
        if (M.two()) {
            int64_t ix2 = jround((s.x2 - M.m.inv_min) / M.m.step);
            if (ix2 < 0) { ix2 = 0; *off = true; }
            if (ix2 >= nI) { ix2 = nI - 1; *off = true; }
            idx = idx * nI + ix2;
        }
        if (L >= 1) { int64_t iq = jround(s.q1 / M.m.step); idx = idx * nQ + iq; }
        if (L >= 2) { int64_t iq = jround(s.q2 / M.m.step); idx = idx * nQ + iq; }
        if (M.cash()) {
            int64_t k = M.cash_k(s.w) - kmin;
            if (k < 0) { k = 0; *off = true; }
            if (k >= nW) { k = nW - 1; *off = true; }
            idx = idx * nW + k;
        }
        return idx;
    }
    St state(int t, int64_t idx) const {
        St s; s.t = t;
        if (M.cash()) { s.w = M.cash_of_k(kmin + idx % nW); idx /= nW; }
        if (L >= 2) { s.q2 = (double)(idx % nQ) * M.m.step; idx /= nQ; }
        if (L >= 1) { s.q1 = (double)(idx % nQ) * M.m.step; idx /= nQ; }
        if (M.two()) { s.x2 = M.m.inv_min + (double)(idx % nI) * M.m.step; idx /= nI; }
        s.x = M.m.inv_min + (double)idx * M.m.step;
        return s;
    }
    int ndim() const { return 1 + (M.two() ? 1 : 0) + (M.cash() ? 1 : 0) + L; }
    // API order: (inv) | (inv, preQ[, preQ2]) | (inv, cash) | (inv, cash, preQ)
    void to_api(const St& s, double* out) const {
        int k = 0; out[k++] = s.x;
        if (M.two()) out[k++] = s.x2;
        if (M.cash()) out[k++] = s.w;
        if (L >= 1) out[k++] = s.q1;
        if (L >= 2) out[k++] = s.q2;
    }
    St from_api(int t, const double* in) const {
        St s; s.t = t; int k = 0; s.x = in[k++];
        if (M.two()) s.x2 = in[k++];
        if (M.cash()) s.w = in[k++];
        if (L >= 1) s.q1 = in[k++];
        if (L >= 2) s.q2 = in[k++];
        return s;
    }
};

Model make_model(const sdpb_model* m) {
    Model M; M.m = *m;
    if (m->cost_kind == SDPB_COST_CASH_TWO_PRODUCT) M.tie_tol = m->tie_tolerance;
    if (m->cost_kind == SDPB_COST_STAFF) {
        const size_t rows = (size_t)m->T * M.n_inv();
        M.aoff.assign(rows + 1, 0);
        for (size_t r = 0; r < rows; r++) M.aoff[r + 1] = M.aoff[r] + m->apmf_len[r];
    }
    M.off.resize(m->T + 1, 0);
    for (int t = 0; t < m->T; t++) M.off[t + 1] = M.off[t] + m->pmf_len[t];
    return M;
}

}  // namespace

// ---- two more two-product recursions of the reference, restated for the reached-state engine's parity tests ----
// kind 1: CashRecursionMultiXR.getExpectedValue (src/sdp/cash/multiItem/CashRecursionMultiXR.java:61-95) with the
//         lambdas of src/cash/multiItem/MultiItemCashXR.java:73-126 -- state (x1, x2, R), action = order-up-to pair;
// kind 2: CashRecursionV.getExpectedValueV / getExpectedValuePai (src/sdp/cash/multiItem/CashRecursionV.java:83-131)
//         with the lambdas of src/cash/multiItem/MultiItemYR.java:89-146 -- state (x1, x2, w), V_{T+1} = boundFinalCash.
// The reference records no output for either (parity unpinned); the restatement follows the cited lines.
struct RKey {
    int t; double x1, x2, c;
    bool operator<(const RKey& o) const {
        if (t != o.t) return t < o.t;
        if (x1 != o.x1) return x1 < o.x1;
        if (x2 != o.x2) return x2 < o.x2;
        return c < o.c;
    }
};
struct RParamsO {
    int kind, T, Q, nD;
    const int* nDt;  // demand points per period (rows are nD apart)
    const double *d1, *d2, *p;
    double price[2], v[2], sal[2], deposit_rate, min_inv, max_inv, min_cash, max_cash, gamma;
    double K, h, min_cash_required, q;  // kind 3
};
struct ReachedTopDown {
    const RParamsO& P;
    std::map<RKey, double> values;
    std::map<RKey, std::pair<double, double>> actions;
    explicit ReachedTopDown(const RParamsO& p) : P(p) {}

    // ---- kind 1 ----
    double xr_immediate(const RKey& s, double action1, double action2, double demand1, double demand2) const {
        double endInventory1 = jmax(0, action1 - demand1);
        double endInventory2 = jmax(0, action2 - demand2);
        double revenue1 = P.price[0] * (action1 - endInventory1);
        double revenue2 = P.price[1] * (action2 - endInventory2);
        double revenue = revenue1 + revenue2;
        double initialCash = s.c - P.v[0] * s.x1 - P.v[1] * s.x2;
        double orderingCostY1 = P.v[0] * action1;
        double orderingCostY2 = P.v[1] * action2;
        double orderingCostsY = orderingCostY1 + orderingCostY2;
        double salValue = 0;
        if (s.t == P.T) salValue = P.sal[0] * endInventory1 + P.sal[1] * endInventory2;
        return revenue + (1 - P.deposit_rate) * (s.c - orderingCostsY) + salValue - initialCash;
    }
    RKey xr_transition(const RKey& s, double a1, double a2, double d1, double d2) const {
        double endInventory1 = a1 - d1;
        endInventory1 = jmax(0, endInventory1);
        double endInventory2 = a2 - d2;
        endInventory2 = jmax(0, endInventory2);
        double initialCash = s.c - P.v[0] * s.x1 - P.v[1] * s.x2;
        double nextCash = initialCash + xr_immediate(s, a1, a2, d1, d2);
        nextCash = nextCash > P.max_cash ? P.max_cash : nextCash;
        nextCash = nextCash < P.min_cash ? P.min_cash : nextCash;
        endInventory1 = endInventory1 > P.max_inv ? P.max_inv : endInventory1;
        endInventory2 = endInventory2 < P.min_inv ? P.min_inv : endInventory2;
        nextCash = (int)nextCash;
        endInventory1 = (int)endInventory1;
        endInventory2 = (int)endInventory2;
        double nextR = (int)nextCash + P.v[0] * endInventory1 + P.v[1] * endInventory2;
        return RKey{s.t + 1, endInventory1, endInventory2, nextR};
    }
    double xr_value(const RKey& s) {
        auto it = values.find(s);
        if (it != values.end()) return it->second;
        double val = -DBL_MAX;
        std::pair<double, double> best{0, 0};
        const int miny1 = (int)s.x1, miny2 = (int)s.x2;
        for (int i = miny1; i < miny1 + P.Q; i++)
            for (int j = miny2; j < miny2 + P.Q; j++) {
                double thisActionsValue = 0;
                for (int k = 0; k < P.nDt[s.t - 1]; k++) {
                    const double dd1 = P.d1[(s.t - 1) * P.nD + k], dd2 = P.d2[(s.t - 1) * P.nD + k], pr = P.p[(s.t - 1) * P.nD + k];
                    thisActionsValue += pr * xr_immediate(s, i, j, dd1, dd2);
                    if (s.t < P.T) thisActionsValue += pr * P.gamma * xr_value(xr_transition(s, i, j, dd1, dd2));
                }
                if (thisActionsValue > val + 0.1) { val = thisActionsValue; best = {(double)i, (double)j}; }
            }
        values.emplace(s, val);
        actions.emplace(s, best);
        return val;
    }

    // ---- kind 2 ----
    RKey yr_transition(int period, double y1, double y2, double iniR, double d1, double d2) const {
        double endInventory1 = y1 - d1;
        endInventory1 = jmax(0, endInventory1);
        double endInventory2 = y2 - d2;
        endInventory2 = jmax(0, endInventory2);
        double revenue1 = P.price[0] * jmin(y1, d1);
        double revenue2 = P.price[1] * jmin(y2, d2);
        double nextW = revenue1 + revenue2 + (1 + P.deposit_rate) * (iniR - P.v[0] * y1 - P.v[1] * y2);
        endInventory1 = (double)(jround(endInventory1 * 10) / 10);   // long division
        endInventory2 = (double)(jround(endInventory2 * 10) / 10);
        nextW = (double)(jround(nextW * 10) / 10);
        nextW = nextW > P.max_cash ? P.max_cash : nextW;
        nextW = nextW < P.min_cash ? P.min_cash : nextW;
        endInventory1 = endInventory1 > P.max_inv ? P.max_inv : endInventory1;
        endInventory2 = endInventory2 < P.min_inv ? P.min_inv : endInventory2;
        return RKey{period + 1, endInventory1, endInventory2, nextW};
    }
    double yr_value(const RKey& s) {
        if (s.t > P.T) return s.c + P.sal[0] * s.x1 + P.sal[1] * s.x2;  // boundFinalCash, MultiItemYR.java:116-119
        auto it = values.find(s);
        if (it != values.end()) return it->second;
        double val = -DBL_MAX;
        std::pair<double, double> best{s.x1, s.x2};
        const int miny1 = (int)s.x1, miny2 = (int)s.x2;
        const double iniRf = s.c + P.v[0] * s.x1 + P.v[1] * s.x2;
        for (double i = miny1; i < miny1 + P.Q; i = i + 1)
            for (double j = miny2; j < miny2 + P.Q; j = j + 1) {
                if (!(P.v[0] * i + P.v[1] * j < iniRf + 0.1)) continue;
                const double iniR = s.c + P.v[0] * s.x1 + P.v[1] * s.x2;
                double expectValue = 0;
                for (int k = 0; k < P.nDt[s.t - 1]; k++) {
                    const double dd1 = P.d1[(s.t - 1) * P.nD + k], dd2 = P.d2[(s.t - 1) * P.nD + k], pr = P.p[(s.t - 1) * P.nD + k];
                    expectValue += pr * yr_value(yr_transition(s.t, i, j, iniR, dd1, dd2));
                }
                if (expectValue > val + 0.01) { val = expectValue; best = {i, j}; }
            }
        values.emplace(s, val);
        actions.emplace(s, best);
        return val;
    }

    // ---- kind 3: CashRecursion.getExpectedValue (CashRecursion.java:98-138, MAX) with CashConstraintTest.java:76-116 ----
    double cct_immediate(const RKey& s, double action, double randomDemand) const {
        double revenue = P.price[0] * jmin(s.x1 + action, randomDemand);
        double fixedCost = action > 0 ? P.K : 0;
        double variableCost = P.v[0] * action;
        double inventoryLevel = s.x1 + action - randomDemand;
        double holdCosts = P.h * jmax(inventoryLevel, 0);
        double interests = P.deposit_rate * (s.c - action * P.v[0]);
        double cashIncrement = revenue - fixedCost - variableCost - holdCosts + interests;
        double salValue = s.t == P.T ? P.sal[0] * jmax(inventoryLevel, 0) : 0;
        cashIncrement += salValue;
        return cashIncrement;
    }
    RKey cct_transition(const RKey& s, double action, double randomDemand) const {
        double nextInventory = jmax(0, s.x1 + action - randomDemand);
        double nextCash = s.c + cct_immediate(s, action, randomDemand);
        nextCash = nextCash > P.max_cash ? P.max_cash : nextCash;
        nextCash = nextCash < P.min_cash ? P.min_cash : nextCash;
        nextInventory = nextInventory > P.max_inv ? P.max_inv : nextInventory;
        nextInventory = nextInventory < P.min_inv ? P.min_inv : nextInventory;
        nextCash = jround(nextCash * P.q) / P.q;
        nextInventory = jround(nextInventory * P.q) / P.q;
        return RKey{s.t + 1, nextInventory, 0.0, nextCash};
    }
    double cct_value(const RKey& s) {
        auto it = values.find(s);
        if (it != values.end()) return it->second;
        double maxQ = (int)jmin((double)(P.Q - 1), jmax(0, (s.c - P.min_cash_required - P.K) / P.v[0]));
        const int n = (int)maxQ + 1;
        double val = -DBL_MAX, bestOrderQty = 0;
        for (int i = 0; i < n; i++) {
            double orderQty = i;
            double thisQValue = 0;
            for (int j = 0; j < P.nDt[s.t - 1]; j++) {
                const double randomDemand = P.d1[(s.t - 1) * P.nD + j], dProb = P.p[(s.t - 1) * P.nD + j];
                thisQValue += dProb * cct_immediate(s, orderQty, randomDemand);
                if (s.t < P.T) thisQValue += dProb * P.gamma * cct_value(cct_transition(s, orderQty, randomDemand));
            }
            if (thisQValue > val) { val = thisQValue; bestOrderQty = orderQty; }
        }
        values.emplace(s, val);
        actions.emplace(s, std::make_pair(bestOrderQty, 0.0));
        return val;
    }
};

extern "C" int oracle_reached(int kind, int T, int Qbound, int nD, const double* d1, const double* d2, const double* p,
                              const double* price, const double* v, const double* sal, double deposit_rate, double min_inv,
                              double max_inv, double min_cash, double max_cash, double gamma, const double* init /* x1, x2, c */,
                              double* value, double* a1, double* a2, int64_t* n_states, const int* nDt, double fixed_cost,
                              double hold_cost, double min_cash_required, double state_q) {
    RParamsO P;
    P.kind = kind; P.T = T; P.Q = Qbound; P.nD = nD; P.d1 = d1; P.d2 = d2; P.p = p; P.nDt = nDt;
    P.K = fixed_cost; P.h = hold_cost; P.min_cash_required = min_cash_required; P.q = state_q;
    for (int k = 0; k < 2; k++) { P.price[k] = price[k]; P.v[k] = v[k]; P.sal[k] = sal[k]; }
    P.deposit_rate = deposit_rate; P.min_inv = min_inv; P.max_inv = max_inv; P.min_cash = min_cash; P.max_cash = max_cash;
    P.gamma = gamma;
    ReachedTopDown td(P);
    const RKey s0{1, init[0], init[1], init[2]};
    *value = kind == 1 ? td.xr_value(s0) : kind == 2 ? td.yr_value(s0) : td.cct_value(s0);
    *a1 = td.actions[s0].first;
    *a2 = td.actions[s0].second;
    if (n_states) *n_states = (int64_t)td.values.size();
    return 0;
}

extern "C" {

// Grid sizes the dense oracle uses for `m` (states per period, API state length).
int oracle_grid(const sdpb_model* m, int64_t* n_states, int* ndim) {
    Model M = make_model(m);
    Grid G(M);
    *n_states = G.S;
    *ndim = G.ndim();
    return 0;
}

// Period-by-period solve of the whole grid.  V and Q are [T][n_states] (period 1 first); Q is the
// order quantity as a double.  `offgrid` (may be NULL) counts (s,a,d) triples whose successor fell
// outside the grid and was clipped (only possible without SDPB_F_CLAMP_INV).  `threads` <= 0 means
// all cores.  Returns total evaluations in *evals.
int oracle_dense(const sdpb_model* m, double* V, double* Q, double* evals, int64_t* offgrid,
                 int threads) {
    Model M = make_model(m);
    Grid G(M);
    const int T = m->T;
    const int64_t S = G.S;
    double total = 0;
    int64_t off_total = 0;
    int nthreads = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (nthreads < 1) nthreads = 1;
    for (int t = T; t >= 1; t--) {
        const double* Vn = t < T ? V + (size_t)t * S : m->terminal_value;
        double* Vt = V + (size_t)(t - 1) * S;
        double* Qt = Q + (size_t)(t - 1) * S;
        std::atomic<int64_t> next{0};
        const int64_t chunk = 64;
        std::vector<double> ev_part(nthreads, 0.0);
        std::vector<int64_t> off_part(nthreads, 0);
        auto worker = [&](int tid) {
            double ev = 0;
            int64_t offc = 0;
            for (;;) {
                int64_t lo = next.fetch_add(chunk);
                if (lo >= S) break;
                int64_t hi = std::min(S, lo + chunk);
                for (int64_t idx = lo; idx < hi; idx++) {
                    St s = G.state(t, idx);
                    double val, best;
                    solve_state(M, s,
                                [&](const St& ns) {
                                    bool o = false;
                                    int64_t ni = G.index(ns, &o);
                                    if (o) offc++;
                                    return Vn[ni];
                                },
                                &val, &best, &ev);
                    Vt[idx] = val;
                    Qt[idx] = best;
                }
            }
            ev_part[tid] = ev;
            off_part[tid] = offc;
        };
        if (nthreads == 1) {
            worker(0);
        } else {
            std::vector<std::thread> pool;
            for (int i = 0; i < nthreads; i++) pool.emplace_back(worker, i);
            for (auto& th : pool) th.join();
        }
        for (int i = 0; i < nthreads; i++) { total += ev_part[i]; off_total += off_part[i]; }
    }
    if (evals) *evals = total;
    if (offgrid) *offgrid = off_total;
    return 0;
}

// Literal top-down memoised recursion from `n_init` period-1 states (API order, ndim doubles
// each).  Returns the number of visited states; when `rows` is non-NULL it receives, sorted by
// the memo-map order, rows of (ndim + 3) doubles: [t, state dims (API order)..., Q*, V].
int64_t oracle_topdown(const sdpb_model* m, const double* init_states, int n_init, double* rows,
                       int64_t max_rows, double* init_values, double* evals) {
    Model M = make_model(m);
    Grid G(M);
    TopDown td(M);
    if (m->terminal_value)  // the boundary function, tabulated on the grid by the caller
        td.boundary = [&](const St& s) { bool o = false; return m->terminal_value[G.index(s, &o)]; };
    for (int i = 0; i < n_init; i++) {
        St s = G.from_api(1, init_states + (size_t)i * G.ndim());
        double v = td.getExpectedValue(s);
        if (init_values) init_values[i] = v;
    }
    if (evals) *evals = td.evals;
    int64_t n = (int64_t)td.cacheActions.size();
    if (rows) {
        int w = G.ndim() + 3;
        int64_t r = 0;
        for (auto& kv : td.cacheActions) {
            if (r >= max_rows) break;
            double* row = rows + (size_t)r * w;
            row[0] = kv.first.t;
            G.to_api(kv.first, row + 1);
            row[1 + G.ndim()] = kv.second;
            row[2 + G.ndim()] = td.cacheValues[kv.first];
            r++;
        }
    }
    return n;
}

// The reference's two-product cash + lead-time model, solved top-down from the all-zero state with THIS
// file's solve_state / TopDown (the loop every parity test relies on).  `vals1/probs1`, `vals2/probs2`:
// the discrete demand distributions of the two products (n points each).  Returns the optimal value
// and first-period order quantities; used to reproduce the outputs recorded in
// src/cash/overdraft/MultiProductLeadtime.java:30-50.
int oracle_multi_lead(int T, int Qbound, int n, const double* vals1, const double* probs1, const double* vals2,
                      const double* probs2, double overhead, double* value, int* q1, int* q2, int64_t* n_states) {
    MultiLead ml;
    ml.Qbound = Qbound;
    ml.overheadCost.assign(T, overhead);
    std::vector<double> pd, pp;
    for (int i = 0; i < n; i++)  // GetPmfMulti.java:158-171
        for (int j = 0; j < n; j++) {
            ml.d1.push_back((int)vals1[i]);
            ml.d2.push_back((int)vals2[j]);
            pd.push_back((double)(i * n + j));  // "demand" handed to the loop = flat pair index
            pp.push_back(probs1[i] * probs2[j]);
        }
    std::vector<int32_t> len(T, n * n);
    std::vector<double> d_all, p_all;
    for (int t = 0; t < T; t++) { d_all.insert(d_all.end(), pd.begin(), pd.end()); p_all.insert(p_all.end(), pp.begin(), pp.end()); }
    sdpb_model m;
    std::memset(&m, 0, sizeof(m));
    m.T = T; m.direction = SDPB_MAX; m.recursion = SDPB_REC_EXPECT; m.gamma = 1.0;
    m.pmf_len = len.data(); m.pmf_d = d_all.data(); m.pmf_p = p_all.data();
    Model M = make_model(&m);
    M.multi = &ml;
    M.tie_tol = 0.1;
    TopDown td(M);
    St ini; ini.t = 1;
    double v = td.getExpectedValue(ini);
    double best = td.cacheActions[ini];
    *value = 0.0 + v;  // iniCash + recursion.getExpectedValue(iniState), MultiProductLeadtime.java:236
    *q1 = (int)best / Qbound;
    *q2 = (int)best % Qbound;
    if (n_states) *n_states = (int64_t)td.cacheValues.size();
    return 0;
}

// One backward-induction step for selected states: given the full period-(t+1) value table
// `Vnext` (NULL when t == T), recompute V_t and Q_t at the `n` flattened indices `idx`.
// Lets a test check a full-size GPU solve period by period on a sample of states.
int oracle_step_states(const sdpb_model* m, int period, const double* Vnext, const int64_t* idx, int n,
                       double* v_out, double* q_out) {
    Model M = make_model(m);
    Grid G(M);
    if (!Vnext) Vnext = m->terminal_value;  // period T of a model with a boundary function
    for (int i = 0; i < n; i++) {
        St s = G.state(period, idx[i]);
        double val, best;
        solve_state(M, s,
                    [&](const St& ns) {
                        bool o = false;
                        return Vnext[G.index(ns, &o)];
                    },
                    &val, &best, nullptr);
        v_out[i] = val;
        q_out[i] = best;
    }
    return 0;
}

// Policy roll-out (Simulation.java:59-70, CashSimulation.java:100-111) over caller-supplied sample
// paths, using the dense oracle's policy table Q [T][n_states] (order quantities as doubles).
int oracle_simulate(const sdpb_model* m, const double* Q, const double* init_state, const double* samples, int n,
                    double discount, double* values) {
    Model M = make_model(m);
    Grid G(M);
    const int T = m->T;
    for (int i = 0; i < n; i++) {
        double sum = 0;
        St state = G.from_api(1, init_state);
        for (int t = 0; t < T; t++) {
            state.t = t + 1;
            bool off = false;
            const int64_t idx = G.index(state, &off);
            double optQ = Q[(size_t)t * G.S + idx];
            double randomDemand = (double)jround(samples[(size_t)i * T + t]);
            double thisValue = immediate(M, state, optQ, randomDemand);
            sum += std::pow(discount, (double)t) * thisValue;
            state = transition(M, state, optQ, randomDemand);
        }
        values[i] = sum;
    }
    return 0;
}

// Single (s, a, d) evaluation for descriptor spot checks: c, and the successor in API order.
int oracle_eval(const sdpb_model* m, int period, const double* state, double action, double demand,
                double* c, double* next_state) {
    Model M = make_model(m);
    Grid G(M);
    St s = G.from_api(period, state);
    *c = immediate(M, s, action, demand);
    St n = transition(M, s, action, demand);
    G.to_api(n, next_state);
    return 0;
}

// Flattened grid index of an API-order state (for tests), -1 when outside the grid.
int64_t oracle_index(const sdpb_model* m, const double* state) {
    Model M = make_model(m);
    Grid G(M);
    St s = G.from_api(1, state);
    bool off = false;
    int64_t i = G.index(s, &off);
    return off ? -1 : i;
}

int oracle_n_actions(const sdpb_model* m, int period, const double* state) {
    Model M = make_model(m);
    Grid G(M);
    St s = G.from_api(period, state);
    return n_actions(M, s);
}

}  // extern "C"
