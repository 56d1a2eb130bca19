package sdp.b200;

import java.lang.foreign.*;

import sdp.cash.CashStateXR;

/**
 * Drop-in for sdp.cash.CashRecursionXR (src/sdp/cash/CashRecursionXR.java:39-150): state (period, iniInventory,
 * iniR = w + v x), the action is the order-up-to level y (:82-124), rows [t, x, R, y].
 * Lambdas: src/cash/singleItem/CashConstraintXR.java:71-110.  The reference's action set x..max(x, R/v) has no cap
 * (:71-75); the dense grid has max_order_idx + 1 levels per state -- ModelBuilder.cashXR picks it so that no state is
 * capped, and getOptTable() refuses (SDPB_ERR_OFFGRID) if a visited state were.
 * NOT COMPILED in the build image (no JDK).
 */
public final class GpuCashRecursionXR extends GpuEngine {
    public GpuCashRecursionXR(MemorySegment model) { super(model, 2); }

    /** CashRecursionXR.java:79-125. */
    public double getExpectedValue(CashStateXR state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniR())[0];
    }

    /** CashRecursionXR.java:128-130: the optimal order-up-to level. */
    public double getAction(CashStateXR state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniR())[1];
    }

    /** CashRecursionXR.java:140-150. */
    public double[][] getOptTable() { return optTable(); }
}
