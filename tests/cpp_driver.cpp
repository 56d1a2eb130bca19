// A C++ client of include/sdpb200.hpp written the way the reference's drivers are:
//   section A  src/capacitated/CLSPTesting.java:29-143   (Recursion, getOptTable)
//   section B  src/leadtime/Leadtime.java:25-103          (LeadtimeRecursion)
//   section C  src/cash/singleItem/CashConstraint.java:44-163 (CashRecursion)
//   section F  src/cash/risk/cashSurvival.java             (RiskRecursion.getSurvProb)
// It writes every pmf table it built to <out>.pmf (so the Python test can hand the SAME table to the oracle)
// and prints its results as "key value..." lines.  Exit code 77: no CUDA device (there is no CPU path).
#include <algorithm>
#include <cstdio>
#include <fstream>
#include <iostream>

#include "sdpb200.hpp"

using namespace sdpb200;

static void dump(std::ofstream& f, const char* name, const Pmf& pmf) {
    f << name << " " << pmf.size() << "\n";
    for (const auto& row : pmf) {
        f << row.size();
        for (const auto& dp : row) { char b[80]; std::snprintf(b, sizeof b, " %.17g %.17g", dp[0], dp[1]); f << b; }
        f << "\n";
    }
}

int main(int argc, char** argv) {
    std::ofstream pf(argc > 1 ? argv[1] : "cpp_driver.pmf");
    try {
        {   // ---- A: CLSPTesting.main ----
            double meanDemand[] = {9, 23, 53, 29};
            double fixedOrderingCost = 500, variOrderingCost = 0, penaltyCost = 10, holdingCost = 2;
            int maxOrderQuantity = 60, minInventory = -120, maxInventory = 200;
            double truncationQuantile = 0.9999, stepSize = 1, iniInventory = 0;
            std::vector<PoissonDist> distributions;
            for (double m : meanDemand) distributions.emplace_back(m);
            Pmf pmf = GetPmf(distributions, truncationQuantile, stepSize).getpmf();
            dump(pf, "A", pmf);
            Recursion recursion(Model::inventory(OptDirection::MIN, pmf, fixedOrderingCost, variOrderingCost, holdingCost,
                                                 penaltyCost, maxOrderQuantity, minInventory, maxInventory, stepSize));
            int threw = 0;
            try { recursion.getAction(State{1, iniInventory}); } catch (const SdpbError&) { threw = 1; }
            State initialState{1, iniInventory};
            double finalValue = recursion.getExpectedValue(initialState);
            double optQ = recursion.getAction(initialState);
            auto optTable = recursion.getOptTable();
            std::printf("A %.17g %.17g %zu %d\n", finalValue, optQ, optTable.size(), threw);
            std::printf("A_row0 %.17g %.17g %.17g\n", optTable[0][0], optTable[0][1], optTable[0][2]);
            const auto& last = optTable.back();
            std::printf("A_rowN %.17g %.17g %.17g\n", last[0], last[1], last[2]);
        }
        {   // ---- B: Leadtime.main (lead time 1, unclamped transition) ----
            double meanDemand[] = {4, 6, 5};
            std::vector<PoissonDist> distributions;
            for (double m : meanDemand) distributions.emplace_back(m);
            Pmf pmf = GetPmf(distributions, 0.999, 1).getpmf();
            dump(pf, "B", pmf);
            // the grid is the reachable hull: x in [-sum dmax, T * maxQ]
            LeadtimeRecursion recursion(Model::leadtime(pmf, 0, 1, 2, 10, 12, -45, 36, 1, 1, false));
            LeadtimeState initialState{1, 0, 0};
            double finalValue = recursion.getExpectedValue(initialState);  // (the solve happens here, as in the reference)
            std::printf("B %.17g %.17g\n", finalValue, recursion.getAction(initialState));
        }
        {   // ---- C: CashConstraint.main (0.1 cash grid) ----
            double meanDemand[] = {5, 7, 4};
            std::vector<PoissonDist> distributions;
            for (double m : meanDemand) distributions.emplace_back(m);
            Pmf pmf = GetPmf(distributions, 0.999, 1).getpmf();
            dump(pf, "C", pmf);
            double price = 8, variCost = 1, fixOrderCost = 10, holdingCost = 0.5, salvageValue = 0.5, overheadCost = 4;
            double depositeRate = 0.05, overheadRate = 0.02, penaltyCost = 0.3;
            CashRecursion recursion(Model::cashConstraint(pmf, price, variCost, fixOrderCost, holdingCost, salvageValue,
                                                          overheadCost, overheadRate, depositeRate, penaltyCost, 15, 0, 30,
                                                          0, 150, 0.95));
            CashState initialState{1, 0, 30};
            double finalValue = recursion.getExpectedValue(initialState);
            double optQ = recursion.getAction(initialState);
            std::printf("C %.17g %.17g %zu\n", finalValue, optQ, recursion.getOptTable().size());
        }
        {   // ---- F: cashSurvival (RiskRecursion.getSurvProb) ----
            double meanDemand[] = {6, 5, 7};
            std::vector<PoissonDist> distributions;
            for (double m : meanDemand) distributions.emplace_back(m);
            Pmf pmf = GetPmf(distributions, 0.999, 1).getpmf();
            dump(pf, "F", pmf);
            RiskRecursion recursion(Model::cashSurvival(pmf, {4, 5, 4}, {1, 2, 1}, {12, 10, 14}, 0.5, 0, 0, 0, 20, 0, 40,
                                                        -30, 120));
            CashState initialState{1, 0, 15};
            double survProb = recursion.getSurvProb(initialState);
            std::printf("F %.17g %.17g\n", survProb, recursion.getAction(initialState));
        }
        {   // ---- G: the capacitated model of section A on three shards (sdpb_group_*), and with a boundary function ----
            double meanDemand[] = {9, 23, 53, 29};
            std::vector<PoissonDist> distributions;
            for (double m : meanDemand) distributions.emplace_back(m);
            Pmf pmf = GetPmf(distributions, 0.9999, 1).getpmf();
            Model model = Model::inventory(OptDirection::MIN, pmf, 500, 0, 2, 10, 60, -120, 200, 1);
            Recursion sharded(model, std::vector<int>{0, 0, 0});
            State initialState{1, 0};
            const double shardedValue = sharded.getExpectedValue(initialState);   // (argument evaluation order is unspecified)
            std::printf("G %.17g %.17g\n", shardedValue, sharded.getAction(initialState));
            // FinalCash.BoundaryFuncton: leftover stock is worth 0.75 a unit, a backorder costs 2.5 (CashRecursionV.java:125-128 form)
            std::vector<double> boundFinalCash;
            for (int x = -120; x <= 200; x++) boundFinalCash.push_back(-0.75 * std::max(x, 0) + 2.5 * std::max(-x, 0));
            Model withBoundary(model);
            withBoundary.boundFinalCash(boundFinalCash);
            Recursion bounded(withBoundary);
            const double boundedValue = bounded.getExpectedValue(initialState);
            std::printf("H %.17g %.17g\n", boundedValue, bounded.getAction(initialState));
            // Leadtime.java:63-67 with the grid sized by sdpb_reachable_hull
            double meanB[] = {4, 6, 5};
            std::vector<PoissonDist> distB;
            for (double m : meanB) distB.emplace_back(m);
            Model lead = Model::leadtime(GetPmf(distB, 0.999, 1).getpmf(), 0, 1, 2, 10, 12, 0, 0, 1, 1, false);
            lead.reachableHull({0, 0});
            LeadtimeRecursion hull(lead);
            LeadtimeState s0{1, 0, 0};
            const double hullValue = hull.getExpectedValue(s0);
            std::printf("I %.17g %.17g %.17g %.17g\n", hullValue, hull.getAction(s0), lead.m.inv_min, lead.m.inv_max);
            // CLSPTesting.java:57-61: a sweep over the fixed ordering cost, solved as one batch (sdpb_solve_batch)
            Recursion sweep0(Model::inventory(OptDirection::MIN, pmf, 300, 0, 2, 10, 60, -120, 200, 1));
            Recursion sweep1(Model::inventory(OptDirection::MIN, pmf, 500, 0, 2, 10, 60, -120, 200, 1));
            Recursion sweep2(Model::inventory(OptDirection::MIN, pmf, 700, 0, 2, 10, 60, -120, 200, 1));
            for (int round = 0; round < 3; round++)   // plain, capture, replay
                Engine::solveBatch({&sweep0.engine(), &sweep1.engine(), &sweep2.engine()});
            const double v0 = sweep0.getExpectedValue(initialState), v1 = sweep1.getExpectedValue(initialState);
            const double v2 = sweep2.getExpectedValue(initialState);
            std::printf("J %.17g %.17g %.17g\n", v0, v1, v2);
            // the opt-in collapsed kernel: close to, not identical with, the exact value of section A
            Recursion collapsed(model, -1, SDPB_KERNEL_COLLAPSED);
            const double vc = collapsed.getExpectedValue(initialState);
            std::printf("K %.17g %.17g\n", vc, collapsed.getAction(initialState));
        }
    } catch (const SdpbError& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return e.code == SDPB_ERR_NO_DEVICE ? 77 : 1;
    }
    return 0;
}
