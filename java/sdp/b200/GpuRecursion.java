package sdp.b200;

import java.lang.foreign.*;
import java.util.Map;
import java.util.TreeMap;

import sdp.inventory.State;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for sdp.inventory.Recursion (src/sdp/inventory/Recursion.java:33-186) over libsdpb200.so:
 * same public members, but the three lambdas are replaced by a model descriptor (a GPU cannot call
 * Java closures).  NOT COMPILED in the build image (no JDK); see INTEGRATION.md.
 *
 * <pre>
 *   // CLSPTesting.java:113-119, unchanged apart from the constructor argument
 *   GpuRecursion recursion = new GpuRecursion(GpuRecursion.inventoryModel(OptDirection.MIN, pmf,
 *           fixedOrderingCost, variOrderingCost, holdingCost, penaltyCost, maxOrderQuantity,
 *           minInventory, maxInventory, stepSize));
 *   double finalValue = recursion.getExpectedValue(new State(1, iniInventory));
 *   double q = recursion.getAction(new State(1, iniInventory));
 * </pre>
 */
public final class GpuRecursion implements AutoCloseable {
    private final Arena arena = Arena.ofShared();
    private final MemorySegment handle;
    private boolean solved = false;
    private final java.util.List<double[]> queried = new java.util.ArrayList<>();

    /** Descriptor for the lambdas of CLSPTesting.java:78-106 (family A of SURVEY.md Appendix A). */
    public static MemorySegment inventoryModel(Arena arena, boolean min, double[][][] pmf, double K, double v,
                                               double h, double pi, int maxOrderQuantity, double minInventory,
                                               double maxInventory, double stepSize) {
        int T = pmf.length, n = 0;
        for (double[][] row : pmf) n += row.length;
        MemorySegment len = arena.allocate(JAVA_INT, T), d = arena.allocate(JAVA_DOUBLE, n), p = arena.allocate(JAVA_DOUBLE, n);
        for (int t = 0, k = 0; t < T; t++) {
            len.setAtIndex(JAVA_INT, t, pmf[t].length);
            for (double[] dp : pmf[t]) { d.setAtIndex(JAVA_DOUBLE, k, dp[0]); p.setAtIndex(JAVA_DOUBLE, k++, dp[1]); }
        }
        MemorySegment m = arena.allocate(SdpB200.MODEL);
        m.fill((byte) 0);
        set(m, "struct_size", (int) SdpB200.MODEL.byteSize());
        set(m, "cost_kind", SdpB200.COST_BACKORDER);
        set(m, "recursion", SdpB200.REC_EXPECT);
        set(m, "direction", min ? SdpB200.MIN : SdpB200.MAX);
        set(m, "T", T);
        set(m, "flags", SdpB200.F_CLAMP_INV);
        set(m, "max_order_idx", (int) (maxOrderQuantity / stepSize));
        setD(m, "gamma", 1.0);
        setA(m, "pmf_len", len); setA(m, "pmf_d", d); setA(m, "pmf_p", p);
        setD(m, "inv_min", minInventory); setD(m, "inv_max", maxInventory); setD(m, "step", stepSize);
        setD(m, "fixed_cost", K); setD(m, "vari_cost", v); setD(m, "hold_cost", h); setD(m, "penalty_cost", pi);
        return m;
    }

    public GpuRecursion(MemorySegment model) {
        try {
            MemorySegment out = arena.allocate(ADDRESS);
            int rc = (int) SdpB200.CREATE.invokeExact(model, MemorySegment.NULL, out);
            if (rc != 0) throw new IllegalStateException("sdpb_create: " + SdpB200.lastError(MemorySegment.NULL));
            handle = out.get(ADDRESS, 0);
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** Recursion.java:89 — the whole grid is solved on the first call, then answered from the table. */
    public double getExpectedValue(State state) {
        double[] vq = valueAndAction(state);
        if (state.getPeriod() == 1) queried.add(new double[]{state.getIniInventory()});
        return vq[0];
    }

    /** Recursion.java:165-167 — throws, like the reference's NullPointerException, before any solve. */
    public double getAction(State state) {
        if (!solved) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state)[1];
    }

    /** Recursion.java:177-186: rows [t, x, Q*] for the visited states only, sorted by (t, x). */
    public double[][] getOptTable() {
        try {
            ensureSolved();
            MemorySegment init = arena.allocate(JAVA_DOUBLE, queried.size());
            for (int i = 0; i < queried.size(); i++) init.setAtIndex(JAVA_DOUBLE, i, queried.get(i)[0]);
            SdpB200.check((int) SdpB200.REACH.invokeExact(handle, init, queried.size()), handle);
            MemorySegment n = arena.allocate(JAVA_LONG);
            SdpB200.check((int) SdpB200.OPT_TABLE.invokeExact(handle, MemorySegment.NULL, n), handle);
            int rows = (int) n.get(JAVA_LONG, 0);
            MemorySegment buf = arena.allocate(JAVA_DOUBLE, 3L * rows);
            SdpB200.check((int) SdpB200.OPT_TABLE.invokeExact(handle, buf, n), handle);
            double[][] arr = new double[rows][3];
            for (int i = 0; i < rows; i++)
                for (int k = 0; k < 3; k++) arr[i][k] = buf.getAtIndex(JAVA_DOUBLE, 3L * i + k);
            return arr;
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** Recursion.java:169-171. */
    public Map<State, Double> getCacheActions() {
        Map<State, Double> m = new TreeMap<>((a, b) -> a.getPeriod() != b.getPeriod()
                ? Integer.compare(a.getPeriod(), b.getPeriod())
                : Double.compare(a.getIniInventory(), b.getIniInventory()));
        for (double[] r : getOptTable()) m.put(new State((int) r[0], r[1]), r[2]);
        return m;
    }

    public void setTreeMapCacheAction() { /* Recursion.java:80-86: tables are already sorted */ }

    private void ensureSolved() throws Throwable {
        if (!solved) { SdpB200.check((int) SdpB200.SOLVE.invokeExact(handle), handle); solved = true; }
    }

    private double[] valueAndAction(State s) {
        try (Arena a = Arena.ofConfined()) {
            ensureSolved();
            MemorySegment st = a.allocate(JAVA_DOUBLE, 1), v = a.allocate(JAVA_DOUBLE), q = a.allocate(JAVA_DOUBLE);
            st.set(JAVA_DOUBLE, 0, s.getIniInventory());
            SdpB200.check((int) SdpB200.VALUE.invokeExact(handle, s.getPeriod(), st, 1, v, q), handle);
            return new double[]{v.get(JAVA_DOUBLE, 0), q.get(JAVA_DOUBLE, 0)};
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    @Override public void close() {
        try { SdpB200.DESTROY.invokeExact(handle); } catch (Throwable ignored) { }
        arena.close();
    }

    private static void set(MemorySegment m, String f, int v) {
        m.set(JAVA_INT, SdpB200.MODEL.byteOffset(MemoryLayout.PathElement.groupElement(f)), v);
    }
    private static void setD(MemorySegment m, String f, double v) {
        m.set(JAVA_DOUBLE, SdpB200.MODEL.byteOffset(MemoryLayout.PathElement.groupElement(f)), v);
    }
    private static void setA(MemorySegment m, String f, MemorySegment v) {
        m.set(ADDRESS, SdpB200.MODEL.byteOffset(MemoryLayout.PathElement.groupElement(f)), v);
    }
}
