/* Plain-C client of the multi-GPU entry points of libsdpb200.so: what a single-process host (the JVM of the
 * reference, through Panama FFM) does to spread one solve over several GPUs -- no Python, no torch, no NCCL in the
 * data path.  A lead-time-2 instance (Leadtime.java lambdas, state (x, q1, q2)) is cut into `n` shards on the
 * devices given on the command line (default: 3 shards on device 0); every shard holds only its window of V_t and
 * the shards hand rows to each other through peer-mapped memory inside the library.
 * Prints: V_1 and Q_1 at the initial state, the sum of all V_1 and Q_1 entries, total evaluations, bytes of device
 * memory of the largest shard and of an unsharded handle.  Built and run by tests/test_c_client.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sdpb200.h"

int main(int argc, char** argv) {
    enum { T = 4, D = 6 };
    int32_t len[T];
    double d[T * D], p[T * D];
    const double probs[D] = {0.1, 0.2, 0.3, 0.2, 0.15, 0.05};
    for (int t = 0; t < T; t++) {
        len[t] = D;
        for (int j = 0; j < D; j++) { d[t * D + j] = j; p[t * D + j] = probs[j]; }
    }
    sdpb_model m;
    memset(&m, 0, sizeof m);
    m.struct_size = sizeof m;
    m.cost_kind = SDPB_COST_BACKORDER;
    m.direction = SDPB_MIN;
    m.T = T;
    m.lead_time = 2;
    m.flags = SDPB_F_CLAMP_INV;
    m.max_order_idx = 7;
    m.gamma = 1.0;
    m.pmf_len = len; m.pmf_d = d; m.pmf_p = p;
    m.inv_min = -15; m.inv_max = 15; m.step = 1;
    m.q_mul = 1; m.q_div = 1;
    m.vari_cost = 1; m.hold_cost = 2; m.penalty_cost = 10;

    int devices[64], n = 0;
    for (int i = 1; i < argc && n < 64; i++) devices[n++] = atoi(argv[i]);
    if (n == 0) { n = 3; devices[0] = devices[1] = devices[2] = 0; }

    sdpb_group* g = NULL;
    int rc = sdpb_group_create(&m, NULL, devices, n, &g);
    if (rc != SDPB_OK) {
        fprintf(stderr, "group_create: %d %s\n", rc, sdpb_group_last_error(NULL));
        return rc == SDPB_ERR_NO_DEVICE ? 77 : 1;
    }
    if ((rc = sdpb_group_solve(g)) != SDPB_OK) { fprintf(stderr, "solve: %s\n", sdpb_group_last_error(g)); return 1; }
    sdpb_grid gi;
    sdpb_grid_info(sdpb_group_shard(g, 0), &gi);
    const long long S = gi.n_states;
    double* V = malloc(sizeof(double) * S);
    double* Q = malloc(sizeof(double) * S);
    if ((rc = sdpb_group_period_tables(g, 1, V, Q)) != SDPB_OK) { fprintf(stderr, "tables: %s\n", sdpb_group_last_error(g)); return 1; }
    double sv = 0, sq = 0;
    for (long long i = 0; i < S; i++) { sv += V[i]; sq += Q[i]; }
    const double s0[3] = {0, 0, 0};
    double v, q;
    if ((rc = sdpb_group_value(g, 1, s0, 1, &v, &q)) != SDPB_OK) { fprintf(stderr, "value: %s\n", sdpb_group_last_error(g)); return 1; }
    sdpb_stats st;
    sdpb_group_stats(g, &st);
    long long max_bytes = 0;
    for (int r = 0; r < n; r++) {
        sdpb_grid_info(sdpb_group_shard(g, r), &gi);
        if (gi.device_bytes > max_bytes) max_bytes = gi.device_bytes;
    }
    sdpb_group_destroy(g);
    sdpb_handle* h = NULL;
    long long full_bytes = 0;
    if (sdpb_create(&m, NULL, &h) == SDPB_OK) { sdpb_grid_info(h, &gi); full_bytes = gi.device_bytes; sdpb_destroy(h); }
    printf("%.17g %.17g %.17g %.17g %.0f %lld %lld %lld\n", v, q, sv, sq, st.evals, S, max_bytes, full_bytes);
    free(V); free(Q);
    return 0;
}
