package sdp.b200;

import java.lang.foreign.*;

import sdp.inventory.LeadtimeState;

/**
 * Drop-in for sdp.inventory.LeadtimeRecursion (src/sdp/inventory/LeadtimeRecursion.java:28-102): state
 * (period, iniInventory, preQ), MIN only (:52,66), table rows [t, x, preQ, Q] (:99).
 * The lambdas of src/leadtime/Leadtime.java:50-81 do NOT clamp the inventory (:65-66): size the grid with
 * sdpb_reachable_hull (ModelBuilder.leadtime does); getOptTable() refuses a grid that clipped a visited state.
 * NOT COMPILED in the build image (no JDK).
 */
public final class GpuLeadtimeRecursion extends GpuEngine {
    public GpuLeadtimeRecursion(MemorySegment model) { super(model, 2); }
    public GpuLeadtimeRecursion(MemorySegment model, int[] devices) { super(model, 2, devices, 0); }

    /** LeadtimeRecursion.java:47-75. */
    public double getExpectedValue(LeadtimeState state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getPreQ())[0];
    }

    /** LeadtimeRecursion.java:77-79. */
    public double getAction(LeadtimeState state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getPreQ())[1];
    }

    /** LeadtimeRecursion.java:93-102: rows [t, x, preQ, Q]. */
    public double[][] getOptTable() { return optTable(); }
}
