// multi_item_oracle.cpp — literal restatement of the reference's two-product cash + lead-time SDP,
// the ONLY model in the reference whose source records program output (known answers in the header
// comment of src/cash/overdraft/MultiProductLeadtime.java:30-50).  TEST INFRASTRUCTURE ONLY.
//
// Why it is here although the two-product solvers are outside the five measured configurations: it
// runs the same loop (src/sdp/cash/multiItem/CashRecursionMultiLead.java:54-90 is a copy of
// Recursion.java:129-161 with a 2-D action/demand and a `> val + 0.1` tie tolerance), the same
// `p * discount * V` association, the same four-branch overdraft interest
// (MultiProductLeadtime.java:186-194 == CashOverdraft.java:88-95) and the same lost-sales / clamp /
// (int) idioms as the single-product cash models, and needs no SSJ (discrete demand: the pmf is a
// product of the given probabilities, GetPmfMulti.java:158-171).  Reproducing the author's recorded
// outputs to the last digit therefore pins those shared semantics against the real Java program.
//
// Build: make -C oracle multi_item_oracle ; usage: multi_item_oracle T Qbound nvals v1.. p1.. v2.. p2..
#include <cfloat>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <vector>

struct St { int t; double I1, I2, q1, q2, w; };
struct Less {  // CashRecursionMultiLead.java:45-51
    bool operator()(const St& a, const St& b) const {
        if (a.t != b.t) return a.t < b.t;
        if (a.I1 != b.I1) return a.I1 < b.I1;
        if (a.I2 != b.I2) return a.I2 < b.I2;
        if (a.q1 != b.q1) return a.q1 < b.q1;
        if (a.q2 != b.q2) return a.q2 < b.q2;
        return a.w < b.w;
    }
};

struct P {
    int T, Qbound;
    double price[2] = {5, 10}, variCost[2] = {1, 2}, salValueUnit[2] = {0.5, 1.0};
    double r0 = 0, r1 = 0.1, r2 = 2, limit = 500, interestFreeAmount = 0;
    double minInventoryState = 0, maxInventoryState = 200, minCashState = -500, maxCashState = 5000;
    double discountFactor = 1;
    std::vector<double> overheadCost;
    std::vector<std::vector<double>> pmf;  // rows (d1, d2, p), same for every period here
};

static double jmax(double a, double b) { return a >= b ? a : b; }
static double jmin(double a, double b) { return a <= b ? a : b; }

// MultiProductLeadtime.java:162-198
static double immediateValue(const P& p, const St& s, int a1, int a2, int d1, int d2) {
    double action1 = a1, action2 = a2, demand1 = d1, demand2 = d2;
    double preQ1 = s.q1, preQ2 = s.q2;
    double endInventory1 = jmax(0, s.I1 + preQ1 - demand1);
    double endInventory2 = jmax(0, s.I2 + preQ2 - demand2);
    double revenue1 = p.price[0] * jmin(demand1, s.I1 + preQ1);
    double revenue2 = p.price[1] * jmin(s.I2 + preQ2, demand2);
    double revenue = revenue1 + revenue2;
    double orderingCost1 = p.variCost[0] * action1;
    double orderingCost2 = p.variCost[1] * action2;
    double orderingCosts = orderingCost1 + orderingCost2;
    double salValue = 0;
    if (s.t == p.T) salValue = p.salValueUnit[0] * endInventory1 + p.salValueUnit[1] * endInventory2;
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

// MultiProductLeadtime.java:202-224 (note the asymmetric clamps: upper on item 1, lower on item 2)
static St stateTransition(const P& p, const St& s, int a1, int a2, int d1, int d2) {
    double endInventory1 = s.I1 + s.q1 - d1;
    endInventory1 = jmax(0, endInventory1);
    double endInventory2 = s.I2 + s.q2 - d2;
    endInventory2 = jmax(0, endInventory2);
    double nextCash = s.w + immediateValue(p, s, a1, a2, d1, d2);
    nextCash = nextCash > p.maxCashState ? p.maxCashState : nextCash;
    nextCash = nextCash < p.minCashState ? p.minCashState : nextCash;
    endInventory1 = endInventory1 > p.maxInventoryState ? p.maxInventoryState : endInventory1;
    endInventory2 = endInventory2 < p.minInventoryState ? p.minInventoryState : endInventory2;
    endInventory1 = (int)endInventory1;
    endInventory2 = (int)endInventory2;
    return St{s.t + 1, endInventory1, endInventory2, (double)a1, (double)a2, nextCash};
}

struct Rec {
    const P& p;
    std::map<St, double, Less> cacheValues;
    std::map<St, std::pair<int, int>, Less> cacheActions;
    double evals = 0;
    explicit Rec(const P& pp) : p(pp) {}
    // CashRecursionMultiLead.java:54-90
    double getExpectedValue(const St& s) {
        auto it = cacheValues.find(s);
        if (it != cacheValues.end()) return it->second;
        double val = -DBL_MAX;
        std::pair<int, int> bestActions(0, 0);
        for (int i = 0; i < p.Qbound; i++)          // buildActionList, MultiProductLeadtime.java:150-158
            for (int j2 = 0; j2 < p.Qbound; j2++) {
                double thisActionsValue = 0;
                for (const auto& row : p.pmf) {
                    int d1 = (int)row[0], d2 = (int)row[1];
                    thisActionsValue += row[2] * immediateValue(p, s, i, j2, d1, d2);
                    if (s.t < p.T) {
                        St ns = stateTransition(p, s, i, j2, d1, d2);
                        thisActionsValue += row[2] * p.discountFactor * getExpectedValue(ns);
                    }
                }
                evals += (double)p.pmf.size();
                if (thisActionsValue > val + 0.1) {
                    val = thisActionsValue;
                    bestActions = {i, j2};
                }
            }
        cacheValues.emplace(s, val);
        cacheActions.emplace(s, bestActions);
        return val;
    }
};

int main(int argc, char** argv) {
    if (argc < 4) { std::fprintf(stderr, "usage: %s T Qbound nvals v1.. p1.. v2.. p2.. [overhead]\n", argv[0]); return 2; }
    P p;
    p.T = std::atoi(argv[1]);
    p.Qbound = std::atoi(argv[2]);
    int n = std::atoi(argv[3]);
    if (argc < 4 + 4 * n) return 2;
    std::vector<double> v1(n), p1(n), v2(n), p2(n);
    int k = 4;
    for (int i = 0; i < n; i++) v1[i] = std::atof(argv[k++]);
    for (int i = 0; i < n; i++) p1[i] = std::atof(argv[k++]);
    for (int i = 0; i < n; i++) v2[i] = std::atof(argv[k++]);
    for (int i = 0; i < n; i++) p2[i] = std::atof(argv[k++]);
    double overhead = argc > k ? std::atof(argv[k]) : 100.0;
    p.overheadCost.assign(p.T, overhead);
    for (int i = 0; i < n; i++)  // GetPmfMulti.java:158-171
        for (int j = 0; j < n; j++) p.pmf.push_back({v1[i], v2[j], p1[i] * p2[j]});
    Rec rec(p);
    St ini{1, 0, 0, 0, 0, 0};
    double iniCash = 0;
    double finalValue = iniCash + rec.getExpectedValue(ini);
    auto a = rec.cacheActions[ini];
    std::printf("%.17g %d %d %zu %.0f\n", finalValue, a.first, a.second, rec.cacheValues.size(), rec.evals);
    return 0;
}
