package sdp.b200;

import java.lang.foreign.*;

import sdp.cash.RiskState;

/**
 * Drop-in for sdp.cash.RiskRecursion (src/sdp/cash/RiskRecursion.java:31-135): survival probability, always MAX
 * (:100), terminal indicator 1[w + c >= 0] (:80-84), continuation 0 once the successor cash is negative (:87-95).
 * RiskState ignores its bankruptBefore argument (RiskState.java:15-18), so the flag is not a state dimension.
 * The descriptor must carry recursion = REC_SURVIVAL (ModelBuilder.survival()).
 * NOT COMPILED in the build image (no JDK).
 */
public final class GpuRiskRecursion extends GpuEngine {
    public GpuRiskRecursion(MemorySegment model) { super(model, 2); }

    /** RiskRecursion.java:64-108. */
    public double getSurvProb(RiskState state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash())[0];
    }

    /** RiskRecursion.java:111-113. */
    public double getAction(RiskState state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash())[1];
    }

    /** RiskRecursion.java:123-135: rows [t, x, w, Q]. */
    public double[][] getOptTable() { return optTable(); }
}
