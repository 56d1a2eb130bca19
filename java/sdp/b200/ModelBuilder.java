package sdp.b200;

import java.lang.foreign.*;

import static java.lang.foreign.ValueLayout.*;

/**
 * Fluent builder of sdpb_model descriptors: one factory per reference driver whose lambdas it stands for
 * (the same mapping as stochastic-inventory_b200/models.py, which is the tested one).
 * NOT COMPILED in the build image (no JDK); see INTEGRATION.md.
 */
public final class ModelBuilder {
    private final Arena arena;
    private final MemorySegment m;

    private ModelBuilder(Arena arena, int costKind, double[][][] pmf) {
        this.arena = arena;
        this.m = arena.allocate(SdpB200.MODEL);
        m.fill((byte) 0);
        setI("struct_size", (int) SdpB200.MODEL.byteSize());
        setI("cost_kind", costKind);
        setI("T", pmf.length);
        setD("gamma", 1.0);
        setD("step", 1.0);
        setD("q_mul", 1.0);
        setD("q_div", 1.0);
        int n = 0;
        for (double[][] row : pmf) n += row.length;
        MemorySegment len = arena.allocate(JAVA_INT, pmf.length), d = arena.allocate(JAVA_DOUBLE, n),
                p = arena.allocate(JAVA_DOUBLE, n);
        for (int t = 0, k = 0; t < pmf.length; t++) {
            len.setAtIndex(JAVA_INT, t, pmf[t].length);
            for (double[] dp : pmf[t]) { d.setAtIndex(JAVA_DOUBLE, k, dp[0]); p.setAtIndex(JAVA_DOUBLE, k++, dp[1]); }
        }
        setA("pmf_len", len); setA("pmf_d", d); setA("pmf_p", p);
    }

    /** src/capacitated/CLSPTesting.java:78-106 (and CLSP, CLSPforDraw, LevelFitsS). */
    public static ModelBuilder inventory(Arena a, double[][][] pmf, boolean min, double K, double v, double h, double pi,
                                         int maxOrderQuantity, double minInventory, double maxInventory) {
        return new ModelBuilder(a, SdpB200.COST_BACKORDER, pmf).direction(min).flags(SdpB200.F_CLAMP_INV)
                .maxOrder(maxOrderQuantity).inventory(minInventory, maxInventory)
                .d("fixed_cost", K).d("vari_cost", v).d("hold_cost", h).d("penalty_cost", pi);
    }

    /** src/leadtime/Leadtime.java:50-81: lead time 1, transition NOT clamped (choose the reachable hull as grid). */
    public static ModelBuilder leadtime(Arena a, double[][][] pmf, double K, double v, double h, double pi,
                                        int maxOrderQuantity, double minInventory, double maxInventory) {
        return inventory(a, pmf, true, K, v, h, pi, maxOrderQuantity, minInventory, maxInventory).flags(0).i("lead_time", 1);
    }

    /** src/cash/singleItem/CashConstraint.java:95-133: quantiser Math.round(w*10)/10.0. */
    public static ModelBuilder cashConstraint(Arena a, double[][][] pmf, double price, double v, double K, double h,
                                              double salvage, double overhead, double maxOrderQuantity,
                                              double minInv, double maxInv, double minCash, double maxCash,
                                              double discountFactor) {
        ModelBuilder b = new ModelBuilder(a, SdpB200.COST_CASH_DEPOSIT, pmf).direction(false)
                .flags(SdpB200.F_CLAMP_INV | SdpB200.F_LOST_SALES | SdpB200.F_CASH_LIMITED_ACTIONS)
                .maxOrder((int) maxOrderQuantity).inventory(minInv, maxInv)
                .d("cash_min", minCash).d("cash_max", maxCash).i("quantiser", SdpB200.Q_DIV).d("q_mul", 10).d("q_div", 10.0)
                .d("price", price).d("vari_cost", v).d("fixed_cost", K).d("hold_cost", h).d("salvage", salvage)
                .d("overhead", overhead).d("gamma", discountFactor).d("reserve2", K);
        return b.perPeriod("reserve_t", overhead, pmf.length);   // (w - overhead - K) / v, CashConstraint.java:98
    }

    /** src/cash/overdraft/CashOverdraft.java:72-118: four-branch interest, Math.round(w*10)/10 with LONG division. */
    public static ModelBuilder cashOverdraft(Arena a, double[][][] pmf, double price, double v, double K, double salvage,
                                             double[] overhead, double r0, double r2, double r3, double limit,
                                             double interestFree, double maxOrderQuantity, double minInv, double maxInv,
                                             double minCash, double maxCash) {
        ModelBuilder b = new ModelBuilder(a, SdpB200.COST_CASH_OVERDRAFT, pmf).direction(false)
                .flags(SdpB200.F_CLAMP_INV | SdpB200.F_LOST_SALES).maxOrder((int) maxOrderQuantity).inventory(minInv, maxInv)
                .d("cash_min", minCash).d("cash_max", maxCash).i("quantiser", SdpB200.Q_LONGDIV).d("q_mul", 10).d("q_div", 10)
                .d("price", price).d("vari_cost", v).d("fixed_cost", K).d("salvage", salvage)
                .d("r0", r0).d("r2", r2).d("r3", r3).d("od_limit", limit).d("interest_free", interestFree);
        MemorySegment o = a.allocate(JAVA_DOUBLE, overhead.length);
        for (int t = 0; t < overhead.length; t++) o.setAtIndex(JAVA_DOUBLE, t, overhead[t]);
        b.setA("overhead_t", o);
        return b;
    }

    public ModelBuilder survival() { return i("recursion", SdpB200.REC_SURVIVAL); }   // RiskRecursion.java:64-108

    /** src/cash/overdraft/SingleProductLeadtime.java:72-119: cash + lead time 1, no order in the last period (:74-75),
     *  quantiser Math.round(w*100)/100.0 (:117).  Engine: GpuCashLeadtimeRecursion. */
    public static ModelBuilder cashLeadtime(Arena a, double[][][] pmf, double price, double v, double salvage,
                                            double[] overhead, double r0, double r2, double r3, double limit,
                                            double interestFree, double maxOrderQuantity, double minInv, double maxInv,
                                            double minCash, double maxCash) {
        return cashOverdraft(a, pmf, price, v, 0, salvage, overhead, r0, r2, r3, limit, interestFree, maxOrderQuantity,
                minInv, maxInv, minCash, maxCash)
                .flags(SdpB200.F_CLAMP_INV | SdpB200.F_LOST_SALES | SdpB200.F_NO_ORDER_LAST).i("lead_time", 1)
                .i("quantiser", SdpB200.Q_DIV).d("q_mul", 100).d("q_div", 100.0);
    }

    /** src/cash/singleItem/CashConstraintXR.java:71-110: (x, R = w + v x) coordinates, action = order-up-to level.
     *  max_order_idx is chosen so that no state's order-up-to range x..max(x, R/v) is capped (:71-75 has no cap). */
    public static ModelBuilder cashXR(Arena a, double[][][] pmf, double price, double v, double K, double h, double salvage,
                                      double overhead, double minInv, double maxInv, double minCash, double maxCash,
                                      double discountFactor) {
        int levels = (int) Math.ceil((maxCash + v * maxInv) / v - minInv) + 1;
        return new ModelBuilder(a, SdpB200.COST_CASH_XR, pmf).direction(false)
                .flags(SdpB200.F_CLAMP_INV | SdpB200.F_LOST_SALES).maxOrder(levels).inventory(minInv, maxInv)
                .d("cash_min", minCash).d("cash_max", maxCash).i("quantiser", SdpB200.Q_LONGDIV).d("q_mul", 1).d("q_div", 1)
                .d("price", price).d("vari_cost", v).d("fixed_cost", K).d("hold_cost", h).d("salvage", salvage)
                .d("overhead", overhead).d("gamma", discountFactor);
    }

    /** src/cash/risk/cashSurvival.java:102-147 with per-period price / cost / overhead arrays; engine GpuRiskRecursion. */
    public static ModelBuilder cashSurvival(Arena a, double[][][] pmf, double[] price, double[] v, double[] overhead,
                                            double salvage, double maxOrderQuantity, double minInv, double maxInv,
                                            double minCash, double maxCash) {
        ModelBuilder b = new ModelBuilder(a, SdpB200.COST_CASH_DEPOSIT, pmf).direction(false).survival()
                .flags(SdpB200.F_CLAMP_INV | SdpB200.F_LOST_SALES | SdpB200.F_CASH_LIMITED_ACTIONS)
                .maxOrder((int) maxOrderQuantity).inventory(minInv, maxInv).d("cash_min", minCash).d("cash_max", maxCash)
                .i("quantiser", SdpB200.Q_LONGDIV).d("q_mul", 1).d("q_div", 1).d("salvage", salvage);
        b.array("price_t", price); b.array("vari_cost_t", v); b.array("overhead_t", overhead);
        return b;
    }

    /**
     * The terminal boundary function (FinalCash.BoundaryFuncton, src/sdp/inventory/FinalCash.java:16-18; consumed at
     * src/sdp/cash/multiItem/CashRecursionV.java:125-128): `table[i]` is the value of grid state i of period T+1, in
     * the library's state order (inventory outermost, cash innermost; sdpb_state_of_index gives the coordinates).
     */
    public ModelBuilder boundFinalCash(double[] table) { array("terminal_value", table); return this; }

    /** Leadtime.java:63-67 does not clamp: make the inventory axis the reachable hull of the initial states. */
    public ModelBuilder reachableHull(double[] initStates, int nStates) {
        try {
            MemorySegment st = arena.allocate(JAVA_DOUBLE, initStates.length), lo = arena.allocate(JAVA_DOUBLE),
                    hi = arena.allocate(JAVA_DOUBLE);
            for (int k = 0; k < initStates.length; k++) st.setAtIndex(JAVA_DOUBLE, k, initStates[k]);
            int rc = (int) SdpB200.REACHABLE_HULL.invokeExact(m, st, nStates, lo, hi);
            if (rc != 0) throw new IllegalStateException("sdpb_reachable_hull: " + rc);
            return inventory(lo.get(JAVA_DOUBLE, 0), hi.get(JAVA_DOUBLE, 0));
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private void array(String f, double[] v) {
        MemorySegment s = arena.allocate(JAVA_DOUBLE, v.length);
        for (int t = 0; t < v.length; t++) s.setAtIndex(JAVA_DOUBLE, t, v[t]);
        setA(f, s);
    }

    public MemorySegment build() { return m; }

    // ---- plumbing ----
    public ModelBuilder direction(boolean min) { return i("direction", min ? SdpB200.MIN : SdpB200.MAX); }
    public ModelBuilder flags(int f) { return i("flags", f); }
    public ModelBuilder maxOrder(int q) { return i("max_order_idx", q); }
    public ModelBuilder inventory(double lo, double hi) { return d("inv_min", lo).d("inv_max", hi); }
    public ModelBuilder i(String f, int v) { setI(f, v); return this; }
    public ModelBuilder d(String f, double v) { setD(f, v); return this; }
    private ModelBuilder perPeriod(String f, double v, int T) {
        MemorySegment s = arena.allocate(JAVA_DOUBLE, T);
        for (int t = 0; t < T; t++) s.setAtIndex(JAVA_DOUBLE, t, v);
        setA(f, s);
        return this;
    }
    private long off(String f) { return SdpB200.MODEL.byteOffset(MemoryLayout.PathElement.groupElement(f)); }
    private void setI(String f, int v) { m.set(JAVA_INT, off(f), v); }
    private void setD(String f, double v) { m.set(JAVA_DOUBLE, off(f), v); }
    private void setA(String f, MemorySegment v) { m.set(ADDRESS, off(f), v); }
}
