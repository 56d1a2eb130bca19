package sdp.b200;

import java.lang.foreign.*;

import sdp.cash.CashState;

/**
 * Drop-in for sdp.cash.CashRecursion (src/sdp/cash/CashRecursion.java:39-220): state (period, iniInventory,
 * iniCash), MIN or MAX, discount factor applied as (p * gamma) * V (:120), table rows [t, x, w, Q] (:215).
 * getSurvProb (:143-194) needs a descriptor built with ModelBuilder.survival().
 * Lambdas covered by the descriptor: CashConstraint.java:95-133, CashOverdraft.java:72-118,
 * CashOverdraftLimit.java:62-99, CashOverdraftTesting.java:78-120, TestPaper.java:76-110, cashSurvival.java:102-147.
 * NOT COMPILED in the build image (no JDK).
 */
public class GpuCashRecursion extends GpuEngine {
    public GpuCashRecursion(MemorySegment model) { super(model, 2); }
    public GpuCashRecursion(MemorySegment model, int[] devices) { super(model, 2, devices, 0); }

    /** CashRecursion.java:79-140. */
    public double getExpectedValue(CashState state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash())[0];
    }

    /** CashRecursion.java:143-194: the same tables hold survival probabilities when the descriptor says REC_SURVIVAL. */
    public double getSurvProb(CashState state) { return getExpectedValue(state); }

    /** CashRecursion.java:197-199. */
    public double getAction(CashState state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash())[1];
    }

    /** CashRecursion.java:209-220: rows [t, x, w, Q] (the reference allocates [n][3] but stores 4-wide rows). */
    public double[][] getOptTable() { return optTable(); }
}
