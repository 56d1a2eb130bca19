package sdp.b200;

import java.lang.foreign.*;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for the three two-product engines whose states do not live on a grid, solved over the states the recursion
 * reaches from the initial state (sdpb_reached_solve):
 *   kind REACHED_MULTILEAD  sdp.cash.multiItem.CashRecursionMultiLead (CashRecursionMultiLead.java:28-95) with the
 *                           lambdas of src/cash/overdraft/MultiProductLeadtime.java:150-224
 *   kind REACHED_MULTI_XR   sdp.cash.multiItem.CashRecursionMultiXR (CashRecursionMultiXR.java:39-95) with the lambdas of
 *                           src/cash/multiItem/MultiItemCashXR.java:73-126
 *   kind REACHED_MULTI_YR   sdp.cash.multiItem.CashRecursionV (CashRecursionV.java:47-131, boundFinalCash as the boundary
 *                           function) with the lambdas of src/cash/multiItem/MultiItemYR.java:89-146
 * getExpectedValue(iniState) / getAction(iniState) of the reference become one call; `pmf` is GetPmfMulti's table
 * (rows {demand1, demand2, probability}, the same number of rows in every period).
 * NOT COMPILED in the build image (no JDK); the struct layout is checked against java/LAYOUT.txt by tests/test_abi.py.
 */
public final class GpuReachedRecursion implements AutoCloseable {
    private final Arena arena = Arena.ofShared();
    private final MemorySegment model;
    private final int T;
    private double value = Double.NaN;
    private final double[] action = new double[2];
    private long[] statesPerPeriod;

    public GpuReachedRecursion(int kind, double[][][] pmf, double[] price, double[] variCost, double[] salvage,
                               double[] overheadCost, double r0, double r1, double r2, double limit, double interestFree,
                               double depositeRate, int Qbound, double minInventory, double maxInventory, double minCash,
                               double maxCash, double discountFactor, double tieTolerance) {
        T = pmf.length;
        int nD = pmf[0].length;
        MemorySegment d1 = arena.allocate(JAVA_DOUBLE, (long) T * nD), d2 = arena.allocate(JAVA_DOUBLE, (long) T * nD),
                p = arena.allocate(JAVA_DOUBLE, (long) T * nD), ovh = arena.allocate(JAVA_DOUBLE, T);
        for (int t = 0; t < T; t++) {
            ovh.setAtIndex(JAVA_DOUBLE, t, overheadCost == null ? 0.0 : overheadCost[t]);
            for (int j = 0; j < nD; j++) {
                d1.setAtIndex(JAVA_DOUBLE, (long) t * nD + j, pmf[t][j][0]);
                d2.setAtIndex(JAVA_DOUBLE, (long) t * nD + j, pmf[t][j][1]);
                p.setAtIndex(JAVA_DOUBLE, (long) t * nD + j, pmf[t][j][2]);
            }
        }
        model = arena.allocate(SdpB200.REACHED_MODEL);
        model.fill((byte) 0);
        setI("struct_size", (int) SdpB200.REACHED_MODEL.byteSize());
        setI("kind", kind); setI("T", T); setI("q_bound", Qbound); setI("n_demands", nD);
        setA("d1", d1); setA("d2", d2); setA("p", p); setA("overhead_t", ovh);
        setD2("price", price); setD2("vari_cost", variCost); setD2("salvage", salvage);
        setD("r0", r0); setD("r1", r1); setD("r2", r2); setD("limit", limit); setD("interest_free", interestFree);
        setD("deposit_rate", depositeRate);
        setD("min_inv", minInventory); setD("max_inv", maxInventory); setD("min_cash", minCash); setD("max_cash", maxCash);
        setD("gamma", discountFactor); setD("tie_tolerance", tieTolerance);
    }

    /** getExpectedValue(iniState): iniState = (x1, x2, preQ1, preQ2, cash) | (x1, x2, R) | (x1, x2, cash). */
    public double getExpectedValue(double... iniState) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment st = a.allocate(JAVA_DOUBLE, iniState.length), v = a.allocate(JAVA_DOUBLE),
                    a1 = a.allocate(JAVA_DOUBLE), a2 = a.allocate(JAVA_DOUBLE), ns = a.allocate(JAVA_LONG, T),
                    ms = a.allocate(JAVA_DOUBLE);
            for (int k = 0; k < iniState.length; k++) st.setAtIndex(JAVA_DOUBLE, k, iniState[k]);
            int rc = (int) SdpB200.REACHED_SOLVE.invokeExact(model, -1, st, v, a1, a2, ns, ms);
            if (rc != 0) throw new IllegalStateException("sdpb_reached_solve: "
                    + ((MemorySegment) SdpB200.REACHED_LAST_ERROR.invokeExact()).reinterpret(4096).getString(0));
            value = v.get(JAVA_DOUBLE, 0);
            action[0] = a1.get(JAVA_DOUBLE, 0);
            action[1] = a2.get(JAVA_DOUBLE, 0);
            statesPerPeriod = new long[T];
            for (int t = 0; t < T; t++) statesPerPeriod[t] = ns.getAtIndex(JAVA_LONG, t);
            return value;
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** getAction(iniState) of the last solve: order quantities (MULTILEAD) or order-up-to levels. */
    public double[] getAction() {
        if (Double.isNaN(value)) throw new NullPointerException("getAction on a state that was never solved");
        return action.clone();
    }

    public long[] getStatesPerPeriod() { return statesPerPeriod; }

    @Override public void close() { arena.close(); }

    private long off(String f) { return SdpB200.REACHED_MODEL.byteOffset(MemoryLayout.PathElement.groupElement(f)); }
    private void setI(String f, int v) { model.set(JAVA_INT, off(f), v); }
    private void setD(String f, double v) { model.set(JAVA_DOUBLE, off(f), v); }
    private void setA(String f, MemorySegment v) { model.set(ADDRESS, off(f), v); }
    private void setD2(String f, double[] v) { model.set(JAVA_DOUBLE, off(f), v[0]); model.set(JAVA_DOUBLE, off(f) + 8, v[1]); }
}
