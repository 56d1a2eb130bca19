package sdp.b200;

import java.lang.foreign.*;
import java.util.Map;
import java.util.TreeMap;

import sdp.inventory.State;

/**
 * Drop-in for sdp.inventory.Recursion (src/sdp/inventory/Recursion.java:33-186) over libsdpb200.so:
 * same public members, but the three lambdas are replaced by a model descriptor (a GPU cannot call
 * Java closures).  NOT COMPILED in the build image (no JDK); see INTEGRATION.md.
 *
 * <pre>
 *   // CLSPTesting.java:113-119, unchanged apart from the constructor argument
 *   GpuRecursion recursion = new GpuRecursion(ModelBuilder.inventory(arena, pmf, true, fixedOrderingCost,
 *           variOrderingCost, holdingCost, penaltyCost, maxOrderQuantity, minInventory, maxInventory).build());
 *   double finalValue = recursion.getExpectedValue(new State(1, iniInventory));
 *   double q = recursion.getAction(new State(1, iniInventory));
 * </pre>
 */
public final class GpuRecursion extends GpuEngine {
    public GpuRecursion(MemorySegment model) { super(model, 1); }
    /** The same solve partitioned over several GPUs of this process (sdpb_group_*). */
    public GpuRecursion(MemorySegment model, int[] devices) { super(model, 1, devices, 0); }

    /** Recursion.java:89 -- the whole grid is solved on the first call, then answered from the table. */
    public double getExpectedValue(State state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory())[0];
    }

    /** Recursion.java:165-167 -- throws, like the reference's NullPointerException, before any solve. */
    public double getAction(State state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory())[1];
    }

    /** Recursion.java:177-186: rows [t, x, Q*] for the visited states only, sorted by (t, x). */
    public double[][] getOptTable() { return optTable(); }

    /** Recursion.java:169-171. */
    public Map<State, Double> getCacheActions() {
        Map<State, Double> m = new TreeMap<>((a, b) -> a.getPeriod() != b.getPeriod()
                ? Integer.compare(a.getPeriod(), b.getPeriod())
                : Double.compare(a.getIniInventory(), b.getIniInventory()));
        for (double[] r : optTable()) m.put(new State((int) r[0], r[1]), r[2]);
        return m;
    }
}
