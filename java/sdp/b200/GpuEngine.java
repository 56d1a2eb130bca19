package sdp.b200;

import java.lang.foreign.*;
import java.util.ArrayList;
import java.util.List;

import static java.lang.foreign.ValueLayout.*;

/**
 * What the six recursion facades share: one libsdpb200 handle (or one group of shards on several GPUs), solved
 * lazily on the first query like the reference's memoised recursion, and the table extraction behind
 * getOptTable().  A state is handed over as its coordinates in the reference constructor's order:
 * (inv) | (inv, preQ) | (inv, cash) | (inv, cash, preQ).
 * NOT COMPILED in the build image (no JDK); see INTEGRATION.md and java/LAYOUT.txt.
 */
abstract class GpuEngine implements AutoCloseable {
    protected final Arena arena = Arena.ofShared();
    protected final MemorySegment handle;   // sdpb_handle*, or NULL when the engine runs on a group
    protected final MemorySegment group;    // sdpb_group*, or NULL
    protected final int ndim;
    private boolean solved = false;
    private final List<double[]> queried = new ArrayList<>();

    /** One GPU (the current device). */
    protected GpuEngine(MemorySegment model, int ndim) { this(model, ndim, null, 0); }

    /** `devices` == null: one GPU.  Otherwise the state grid is partitioned over the listed CUDA ordinals. */
    protected GpuEngine(MemorySegment model, int ndim, int[] devices, int allow) {
        this(model, ndim, devices, allow, SdpB200.KERNEL_AUTO);
    }

    /**
     * `kernel`: SdpB200.KERNEL_AUTO (every kernel the library picks by itself is bit-identical to the Java loop), or a
     * request such as SdpB200.KERNEL_COLLAPSED (1-D inventory family only: G(y) per order-up-to level, ~1e-13 relative).
     */
    protected GpuEngine(MemorySegment model, int ndim, int[] devices, int allow, int kernel) {
        this.ndim = ndim;
        try {
            MemorySegment opt = arena.allocate(SdpB200.OPTIONS);
            opt.fill((byte) 0);
            opt.set(JAVA_INT, 0, (int) SdpB200.OPTIONS.byteSize());
            opt.set(JAVA_INT, SdpB200.OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("device")), -1);
            opt.set(JAVA_INT, SdpB200.OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("shard_count")), 1);
            opt.set(JAVA_INT, SdpB200.OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("allow")), allow);
            opt.set(JAVA_INT, SdpB200.OPTIONS.byteOffset(MemoryLayout.PathElement.groupElement("kernel")), kernel);
            MemorySegment out = arena.allocate(ADDRESS);
            if (devices == null || devices.length < 2) {
                int rc = (int) SdpB200.CREATE.invokeExact(model, opt, out);
                if (rc != 0) throw new IllegalStateException("sdpb_create: " + SdpB200.lastError(MemorySegment.NULL));
                handle = out.get(ADDRESS, 0);
                group = MemorySegment.NULL;
            } else {
                MemorySegment dev = arena.allocate(JAVA_INT, devices.length);
                for (int i = 0; i < devices.length; i++) dev.setAtIndex(JAVA_INT, i, devices[i]);
                int rc = (int) SdpB200.GROUP_CREATE.invokeExact(model, opt, dev, devices.length, out);
                if (rc != 0) throw new IllegalStateException("sdpb_group_create: "
                        + ((MemorySegment) SdpB200.GROUP_LAST_ERROR.invokeExact(MemorySegment.NULL)).reinterpret(4096).getString(0));
                group = out.get(ADDRESS, 0);
                handle = MemorySegment.NULL;
            }
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    protected void ensureSolved() throws Throwable {
        if (solved) return;
        if (group.equals(MemorySegment.NULL)) SdpB200.check((int) SdpB200.SOLVE.invokeExact(handle), handle);
        else if ((int) SdpB200.GROUP_SOLVE.invokeExact(group) != 0)
            throw new IllegalStateException(((MemorySegment) SdpB200.GROUP_LAST_ERROR.invokeExact(group)).reinterpret(4096).getString(0));
        solved = true;
    }

    protected boolean isSolved() { return solved; }

    /** (value, optimal action) of one state; period-1 states become roots of the visited-state table. */
    protected double[] valueAndAction(int period, double... coords) {
        try (Arena a = Arena.ofConfined()) {
            ensureSolved();
            MemorySegment st = a.allocate(JAVA_DOUBLE, ndim), v = a.allocate(JAVA_DOUBLE), q = a.allocate(JAVA_DOUBLE);
            for (int k = 0; k < ndim; k++) st.setAtIndex(JAVA_DOUBLE, k, coords[k]);
            if (group.equals(MemorySegment.NULL))
                SdpB200.check((int) SdpB200.VALUE.invokeExact(handle, period, st, 1, v, q), handle);
            else if ((int) SdpB200.GROUP_VALUE.invokeExact(group, period, st, 1, v, q) != 0)
                throw new IllegalStateException(((MemorySegment) SdpB200.GROUP_LAST_ERROR.invokeExact(group)).reinterpret(4096).getString(0));
            if (period == 1) queried.add(coords.clone());
            return new double[]{v.get(JAVA_DOUBLE, 0), q.get(JAVA_DOUBLE, 0)};
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /**
     * Rows [t, state dims..., Q*] over the states the reference's recursion visits from the queried period-1 states,
     * sorted by (t, state dims) -- the row set and order of getOptTable() in every reference engine.  sdpb_reach also
     * refuses (SDPB_ERR_OFFGRID) if the dense grid clipped a successor or capped an action set at a visited state.
     */
    protected double[][] optTable() {
        if (!group.equals(MemorySegment.NULL))
            throw new UnsupportedOperationException("the visited-state table needs a single-GPU engine (sdpb_reach)");
        try {
            ensureSolved();
            int n0 = queried.size();
            MemorySegment init = arena.allocate(JAVA_DOUBLE, (long) Math.max(1, n0) * ndim);
            for (int i = 0; i < n0; i++)
                for (int k = 0; k < ndim; k++) init.setAtIndex(JAVA_DOUBLE, (long) i * ndim + k, queried.get(i)[k]);
            if (n0 == 0) return new double[0][ndim + 2];
            SdpB200.check((int) SdpB200.REACH.invokeExact(handle, init, n0), handle);
            MemorySegment n = arena.allocate(JAVA_LONG);
            SdpB200.check((int) SdpB200.OPT_TABLE.invokeExact(handle, MemorySegment.NULL, n), handle);
            int rows = (int) n.get(JAVA_LONG, 0), w = ndim + 2;
            MemorySegment buf = arena.allocate(JAVA_DOUBLE, (long) w * Math.max(1, rows));
            SdpB200.check((int) SdpB200.OPT_TABLE.invokeExact(handle, buf, n), handle);
            double[][] arr = new double[rows][w];
            for (int i = 0; i < rows; i++)
                for (int k = 0; k < w; k++) arr[i][k] = buf.getAtIndex(JAVA_DOUBLE, (long) w * i + k);
            return arr;
        } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /**
     * Solve many independent single-GPU engines together (sdpb_solve_batch): the parameter sweeps of
     * CLSPTesting.java:57-61 build hundreds of small recursions one after the other; here their periods run side by side
     * as ONE CUDA graph (first call with a given list: plain solves; second: capture; later: replay).
     */
    public static void solveBatch(List<? extends GpuEngine> engines) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment hs = a.allocate(ADDRESS, Math.max(1, engines.size()));
            for (int i = 0; i < engines.size(); i++) {
                GpuEngine e = engines.get(i);
                if (!e.group.equals(MemorySegment.NULL))
                    throw new IllegalArgumentException("sdpb_solve_batch takes single-GPU engines");
                hs.setAtIndex(ADDRESS, i, e.handle);
            }
            int rc = (int) SdpB200.SOLVE_BATCH.invokeExact(hs, engines.size());
            if (rc != 0) throw new IllegalStateException("sdpb_solve_batch: " + SdpB200.lastError(engines.get(0).handle));
            for (GpuEngine e : engines) e.solved = true;
        } catch (RuntimeException r) { throw r; } catch (Throwable t) { throw new RuntimeException(t); }
    }

    /** Recursion.java:80-86 and siblings: the tables are produced in key order already. */
    public void setTreeMapCacheAction() { }

    @Override public void close() {
        try {
            if (!group.equals(MemorySegment.NULL)) SdpB200.GROUP_DESTROY.invokeExact(group);
            else SdpB200.DESTROY.invokeExact(handle);
        } catch (Throwable ignored) { }
        arena.close();
    }
}
