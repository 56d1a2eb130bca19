/* Plain-C client of libsdpb200.so: what a non-Python host (a JNI shim, a cgo wrapper) does.
 * Family A instance (CLSPTesting.java lambdas): T = 2, demand {1,2,3} w.p. {.2,.5,.3}, K = 4, v = 1,
 * h = 1, pi = 5, inventory -6..6, actions 0..4.  Prints V_1(0), Q_1(0) and the number of opt-table rows.
 * Built and run by tests/test_c_client.py; the expected numbers come from the oracle. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdpb200.h"

int main(void) {
    const int32_t len[2] = {3, 3};
    const double d[6] = {1, 2, 3, 1, 2, 3};
    const double p[6] = {0.2, 0.5, 0.3, 0.2, 0.5, 0.3};
    sdpb_model m;
    memset(&m, 0, sizeof m);
    m.struct_size = sizeof m;
    m.cost_kind = SDPB_COST_BACKORDER;
    m.recursion = SDPB_REC_EXPECT;
    m.direction = SDPB_MIN;
    m.T = 2;
    m.flags = SDPB_F_CLAMP_INV;
    m.max_order_idx = 4;
    m.gamma = 1.0;
    m.pmf_len = len; m.pmf_d = d; m.pmf_p = p;
    m.inv_min = -6; m.inv_max = 6; m.step = 1;
    m.q_mul = 1; m.q_div = 1;
    m.fixed_cost = 4; m.vari_cost = 1; m.hold_cost = 1; m.penalty_cost = 5;

    if (sdpb_abi_version() != SDPB_ABI_VERSION || sdpb_sizeof_model() != sizeof m) {
        fprintf(stderr, "ABI mismatch\n");
        return 2;
    }
    sdpb_handle* h = NULL;
    int rc = sdpb_create(&m, NULL, &h);
    if (rc != SDPB_OK) { fprintf(stderr, "create: %d %s\n", rc, sdpb_last_error(NULL)); return rc == SDPB_ERR_NO_DEVICE ? 77 : 1; }
    if ((rc = sdpb_solve(h)) != SDPB_OK) { fprintf(stderr, "solve: %s\n", sdpb_last_error(h)); return 1; }
    const double s0 = 0.0;
    double v, q;
    if ((rc = sdpb_value(h, 1, &s0, 1, &v, &q)) != SDPB_OK) { fprintf(stderr, "value: %s\n", sdpb_last_error(h)); return 1; }
    if (sdpb_reach(h, &s0, 1) != SDPB_OK) return 1;
    size_t rows = 0;
    if (sdpb_opt_table(h, NULL, &rows) != SDPB_OK) return 1;
    const double bad = 100.0;  /* outside the grid: the reference would throw from getAction */
    int rc_bad = sdpb_value(h, 1, &bad, 1, &v, NULL);
    sdpb_value(h, 1, &s0, 1, &v, &q);
    sdpb_stats st;
    sdpb_stats_get(h, &st);
    printf("%.17g %.17g %zu %d %.0f\n", v, q, rows, rc_bad, st.evals);
    sdpb_destroy(h);
    return 0;
}
