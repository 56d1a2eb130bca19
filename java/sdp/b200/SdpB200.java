package sdp.b200;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (java.lang.foreign, final in JDK 22) binding of include/sdpb200.h.
 *
 * NOT COMPILED OR RUN in the build image (no JDK there); written against the header so that a
 * maintainer with JDK >= 22 can drop it next to src/sdp/inventory/Recursion.java.  The Python ctypes
 * binding (stochastic-inventory_b200/_abi.py) is the one the tests exercise; both mirror the same
 * struct layouts, and sdpb_sizeof_model()/sdpb_sizeof_options() let either verify them at load time.
 */
public final class SdpB200 {
    public static final int COST_BACKORDER = 0, COST_CASH_DEPOSIT = 1, COST_CASH_OVERDRAFT = 2, COST_CASH_XR = 3;
    public static final int REC_EXPECT = 0, REC_SURVIVAL = 1;
    public static final int MIN = 0, MAX = 1;
    public static final int COST_CASH_OD_LIMIT = 4, COST_CASH_OD_TESTING = 5, COST_CASH_LOAN = 6, COST_CASH_TWO_PRODUCT = 7,
            COST_STAFF = 8;
    public static final int Q_DIV = 0, Q_LONGDIV = 1, Q_TRUNC = 2;
    public static final int ALLOW_CLIPPED_SUCCESSORS = 1, ALLOW_CAPPED_ACTIONS = 2;
    public static final int ABI_VERSION = 4;
    public static final int F_CLAMP_INV = 1, F_LOST_SALES = 2, F_GY_MODE = 4, F_NO_ORDER_LAST = 8,
            F_CASH_LIMITED_ACTIONS = 16;

    /** struct sdpb_model, field for field (include/sdpb200.h). */
    public static final StructLayout MODEL = MemoryLayout.structLayout(
            JAVA_INT.withName("struct_size"), JAVA_INT.withName("cost_kind"), JAVA_INT.withName("recursion"),
            JAVA_INT.withName("direction"), JAVA_INT.withName("T"), JAVA_INT.withName("lead_time"),
            JAVA_INT.withName("flags"), JAVA_INT.withName("max_order_idx"), JAVA_DOUBLE.withName("gamma"),
            ADDRESS.withName("pmf_len"), ADDRESS.withName("pmf_d"), ADDRESS.withName("pmf_p"),
            JAVA_DOUBLE.withName("inv_min"), JAVA_DOUBLE.withName("inv_max"), JAVA_DOUBLE.withName("step"),
            JAVA_DOUBLE.withName("cash_min"), JAVA_DOUBLE.withName("cash_max"),
            JAVA_INT.withName("quantiser"), JAVA_INT.withName("q_from_period"),
            JAVA_DOUBLE.withName("q_mul"), JAVA_DOUBLE.withName("q_div"),
            JAVA_DOUBLE.withName("fixed_cost"), JAVA_DOUBLE.withName("vari_cost"), JAVA_DOUBLE.withName("hold_cost"),
            JAVA_DOUBLE.withName("penalty_cost"), JAVA_DOUBLE.withName("price"), JAVA_DOUBLE.withName("salvage"),
            JAVA_DOUBLE.withName("deposit_rate"), JAVA_DOUBLE.withName("overhead_rate"),
            JAVA_DOUBLE.withName("overhead"), JAVA_DOUBLE.withName("r0"), JAVA_DOUBLE.withName("r2"),
            JAVA_DOUBLE.withName("r3"), JAVA_DOUBLE.withName("od_limit"), JAVA_DOUBLE.withName("interest_free"),
            ADDRESS.withName("price_t"), ADDRESS.withName("vari_cost_t"), ADDRESS.withName("overhead_t"),
            ADDRESS.withName("reserve_t"), JAVA_DOUBLE.withName("reserve2"),
            JAVA_DOUBLE.withName("price2"), JAVA_DOUBLE.withName("vari_cost2"), JAVA_DOUBLE.withName("salvage2"),
            ADDRESS.withName("pmf_d2"), JAVA_DOUBLE.withName("tie_tolerance"),
            ADDRESS.withName("apmf_len"), ADDRESS.withName("apmf_p"), ADDRESS.withName("min_level_t"),
            ADDRESS.withName("terminal_value"));

    /** struct sdpb_options, field for field. */
    public static final StructLayout OPTIONS = MemoryLayout.structLayout(
            JAVA_INT.withName("struct_size"), JAVA_INT.withName("device"), JAVA_INT.withName("shard_rank"),
            JAVA_INT.withName("shard_count"), JAVA_INT.withName("kernel"), JAVA_INT.withName("dedup"),
            ADDRESS.withName("stream"), JAVA_INT.withName("allow"), JAVA_INT.withName("strict_cash_bounds"),
            JAVA_INT.withName("profile"), JAVA_INT.withName("reserved"));

    /** enum sdpb_kernel_choice (include/sdpb200.h): the requests a host may make; sdpb_stats.kernel_used reports the rest. */
    public static final int KERNEL_AUTO = 0, KERNEL_GENERIC = 1, KERNEL_TILED = 2, KERNEL_LEAD_Q2 = 9, KERNEL_FUSED = 11,
            KERNEL_COLLAPSED = 15;

    /** struct sdpb_grid, field for field. */
    public static final StructLayout GRID = MemoryLayout.structLayout(
            JAVA_INT.withName("ndim"), JAVA_INT.withName("n_inv"), JAVA_INT.withName("n_cash"), JAVA_INT.withName("n_q"),
            JAVA_LONG.withName("n_states"), JAVA_LONG.withName("shard_lo"), JAVA_LONG.withName("shard_hi"),
            JAVA_INT.withName("n_actions"), JAVA_INT.withName("T"), JAVA_LONG.withName("cash_k_min"),
            JAVA_LONG.withName("window_lo"), JAVA_LONG.withName("window_hi"), JAVA_LONG.withName("device_bytes"));

    /** struct sdpb_reached_model, field for field (price / vari_cost / salvage are double[2]). */
    public static final StructLayout REACHED_MODEL = MemoryLayout.structLayout(
            JAVA_INT.withName("struct_size"), JAVA_INT.withName("kind"), JAVA_INT.withName("T"), JAVA_INT.withName("q_bound"),
            JAVA_INT.withName("n_demands"), JAVA_INT.withName("reserved"),
            ADDRESS.withName("d1"), ADDRESS.withName("d2"), ADDRESS.withName("p"), ADDRESS.withName("overhead_t"),
            MemoryLayout.sequenceLayout(2, JAVA_DOUBLE).withName("price"),
            MemoryLayout.sequenceLayout(2, JAVA_DOUBLE).withName("vari_cost"),
            MemoryLayout.sequenceLayout(2, JAVA_DOUBLE).withName("salvage"),
            JAVA_DOUBLE.withName("r0"), JAVA_DOUBLE.withName("r1"), JAVA_DOUBLE.withName("r2"), JAVA_DOUBLE.withName("limit"),
            JAVA_DOUBLE.withName("interest_free"), JAVA_DOUBLE.withName("deposit_rate"),
            JAVA_DOUBLE.withName("min_inv"), JAVA_DOUBLE.withName("max_inv"), JAVA_DOUBLE.withName("min_cash"),
            JAVA_DOUBLE.withName("max_cash"), JAVA_DOUBLE.withName("gamma"), JAVA_DOUBLE.withName("tie_tolerance"),
            JAVA_DOUBLE.withName("fixed_cost"), JAVA_DOUBLE.withName("hold_cost"), JAVA_DOUBLE.withName("min_cash_required"),
            JAVA_DOUBLE.withName("state_q"), ADDRESS.withName("n_demands_t"));
    public static final int REACHED_MULTILEAD = 0, REACHED_MULTI_XR = 1, REACHED_MULTI_YR = 2, REACHED_CASH_ROUNDED = 3;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
            SymbolLookup.libraryLookup(System.getProperty("sdpb200.lib", "libsdpb200.so"), Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(), d);
    }

    static final MethodHandle SIZEOF_MODEL = fn("sdpb_sizeof_model", FunctionDescriptor.of(JAVA_LONG));
    static final MethodHandle CREATE = fn("sdpb_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle DESTROY = fn("sdpb_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle LAST_ERROR = fn("sdpb_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle SOLVE = fn("sdpb_solve", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle VALUE = fn("sdpb_value",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle REACH = fn("sdpb_reach", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle OPT_TABLE = fn("sdpb_opt_table", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle SOLVE_ASYNC = fn("sdpb_solve_async", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle SOLVE_PERIOD_ASYNC = fn("sdpb_solve_period_async", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle SYNC = fn("sdpb_sync", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle PERIOD_TABLES = fn("sdpb_period_tables", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle DEVICE_TABLES = fn("sdpb_device_tables", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle SHARD_READS = fn("sdpb_shard_reads", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle SIMULATE = fn("sdpb_simulate",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_DOUBLE, ADDRESS));
    static final MethodHandle EVAL_TRIPLES = fn("sdpb_eval_triples", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT,
            ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));

    static final MethodHandle ABI_VERSION_FN = fn("sdpb_abi_version", FunctionDescriptor.of(JAVA_INT));
    static final MethodHandle SIZEOF_OPTIONS = fn("sdpb_sizeof_options", FunctionDescriptor.of(JAVA_LONG));
    static final MethodHandle SIZEOF_GRID = fn("sdpb_sizeof_grid", FunctionDescriptor.of(JAVA_LONG));
    static final MethodHandle GRID_INFO = fn("sdpb_grid_info", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle REACHABLE_HULL = fn("sdpb_reachable_hull",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle SOLVE_BATCH = fn("sdpb_solve_batch", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
    // one process, several GPUs (a JVM is exactly that): the shards exchange rows of V_t inside the library
    static final MethodHandle GROUP_CREATE = fn("sdpb_group_create",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle GROUP_DESTROY = fn("sdpb_group_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    static final MethodHandle GROUP_LAST_ERROR = fn("sdpb_group_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle GROUP_SOLVE = fn("sdpb_group_solve", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle GROUP_SHARD = fn("sdpb_group_shard", FunctionDescriptor.of(ADDRESS, ADDRESS, JAVA_INT));
    static final MethodHandle GROUP_VALUE = fn("sdpb_group_value",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));
    // the two-product recursions over the states they reach (CashRecursionMultiLead / MultiXR / V)
    static final MethodHandle REACHED_SOLVE = fn("sdpb_reached_solve",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    static final MethodHandle REACHED_LAST_ERROR = fn("sdpb_reached_last_error", FunctionDescriptor.of(ADDRESS));
    static final MethodHandle GROUP_PERIOD_TABLES = fn("sdpb_group_period_tables",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));

    static {
        try {
            if ((int) ABI_VERSION_FN.invokeExact() != ABI_VERSION
                    || (long) SIZEOF_MODEL.invokeExact() != MODEL.byteSize()
                    || (long) SIZEOF_OPTIONS.invokeExact() != OPTIONS.byteSize()
                    || (long) SIZEOF_GRID.invokeExact() != GRID.byteSize())
                throw new IllegalStateException("libsdpb200.so struct layout differs from SdpB200 (java/LAYOUT.txt lists "
                        + "the offsets the C compiler produced)");
        } catch (Throwable t) {
            throw new ExceptionInInitializerError(t);
        }
    }

    static String lastError(MemorySegment handle) throws Throwable {
        MemorySegment s = (MemorySegment) LAST_ERROR.invokeExact(handle);
        return s.reinterpret(4096).getString(0);
    }

    static void check(int rc, MemorySegment handle) throws Throwable {
        if (rc != 0) throw new IllegalStateException("sdpb error " + rc + ": " + lastError(handle));
    }

    private SdpB200() {}
}
