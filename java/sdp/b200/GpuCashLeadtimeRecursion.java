package sdp.b200;

import java.lang.foreign.*;

import sdp.cash.CashLeadtimeState;

/**
 * Drop-in for sdp.cash.CashLeadtimeRecursion (src/sdp/cash/CashLeadtimeRecursion.java:28-105): state (period,
 * iniInventory, iniCash, preQ), MAX only (:53,70), no discount, rows [t, x, w, preQ, Q].
 * Lambdas: src/cash/overdraft/SingleProductLeadtime.java:72-119 (no order in the last period, :74-75).
 * The reference's key comparator is inconsistent (:37-41: it compares getIniCash() twice), which makes its memo map
 * miss and duplicate entries; values are unaffected because the function is pure, and the tables here are keyed
 * (period, inventory, preQ, cash) without duplicates.
 * NOT COMPILED in the build image (no JDK).
 */
public final class GpuCashLeadtimeRecursion extends GpuEngine {
    public GpuCashLeadtimeRecursion(MemorySegment model) { super(model, 3); }

    /** CashLeadtimeRecursion.java:48-79. */
    public double getExpectedValue(CashLeadtimeState state) {
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash(), state.getPreQ())[0];
    }

    /** CashLeadtimeRecursion.java:81-83. */
    public double getAction(CashLeadtimeState state) {
        if (!isSolved()) throw new NullPointerException("getAction on a state that was never solved");
        return valueAndAction(state.getPeriod(), state.getIniInventory(), state.getIniCash(), state.getPreQ())[1];
    }

    /** CashLeadtimeRecursion.java:97-105. */
    public double[][] getOptTable() { return optTable(); }
}
