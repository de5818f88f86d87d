/* ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the rectangular linear-sum-assignment solve that the
 * reference calls per clip (sedt/matcher.py:10,95 ->
 * scipy.optimize.linear_sum_assignment).  scipy is a third-party dependency
 * that the reference does not pin (no requirements file); the version in this
 * image is scipy 1.18.1.  Its algorithm is the published one: D. F. Crouse,
 * "On implementing 2D rectangular assignment algorithms", IEEE TAES 52(4),
 * 2016 -- a shortest-augmenting-path (Jonker-Volgenant style) solve in fp64,
 * with the tall matrix transposed first and, on exact ties of the reduced
 * path cost, preference for a column that is still unassigned.
 *
 * tests/test_matcher_oracle.py pins this file against scipy itself on random,
 * tied, integer, empty and tall/wide problems (bit-exact index equality).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may
 * load this file.
 *
 * Build: gcc -O2 -shared -fPIC -o oracle/_build/liblsap_oracle.so oracle/lsap_oracle.c
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* cost: nr x nc row-major fp64.  Writes min(nr,nc) pairs (a[i], b[i]) with a
 * ascending.  Returns 0 on success, -1 if infeasible, -2 on NaN/-inf entries. */
static int lsap_solve_f64(int64_t nr, int64_t nc, const double *cost_in, int64_t *a, int64_t *b)
{
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    double *cost = (double *)cost_in;
    double *temp = NULL;
    if (transpose) {
        temp = (double *)malloc(sizeof(double) * nr * nc);
        for (int64_t i = 0; i < nr; i++)
            for (int64_t j = 0; j < nc; j++) temp[j * nr + i] = cost_in[i * nc + j];
        int64_t t = nr; nr = nc; nc = t;
        cost = temp;
    }
    for (int64_t i = 0; i < nr * nc; i++)
        if (cost[i] != cost[i] || cost[i] == -INFINITY) { free(temp); return -2; }

    double *u = (double *)calloc(nr, sizeof(double));
    double *v = (double *)calloc(nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * nc);
    int64_t *path = (int64_t *)malloc(sizeof(int64_t) * nc);
    int64_t *col4row = (int64_t *)malloc(sizeof(int64_t) * nr);
    int64_t *row4col = (int64_t *)malloc(sizeof(int64_t) * nc);
    int64_t *remaining = (int64_t *)malloc(sizeof(int64_t) * nc);
    char *SR = (char *)malloc(nr), *SC = (char *)malloc(nc);
    for (int64_t i = 0; i < nr; i++) col4row[i] = -1;
    for (int64_t j = 0; j < nc; j++) { row4col[j] = -1; path[j] = -1; }
    int rc = 0;

    for (int64_t cur = 0; cur < nr && rc == 0; cur++) {
        double minVal = 0;
        int64_t i = cur, num_remaining = nc, sink = -1;
        for (int64_t it = 0; it < nc; it++) remaining[it] = nc - it - 1;
        memset(SR, 0, nr); memset(SC, 0, nc);
        for (int64_t j = 0; j < nc; j++) spc[j] = INFINITY;
        while (sink == -1) {
            int64_t index = -1;
            double lowest = INFINITY;
            SR[i] = 1;
            for (int64_t it = 0; it < num_remaining; it++) {
                int64_t j = remaining[it];
                double r = minVal + cost[i * nc + j] - u[i] - v[j];
                if (r < spc[j]) { path[j] = i; spc[j] = r; }
                if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
            }
            minVal = lowest;
            if (minVal == INFINITY) { rc = -1; break; }
            int64_t j = remaining[index];
            if (row4col[j] == -1) sink = j; else i = row4col[j];
            SC[j] = 1;
            remaining[index] = remaining[--num_remaining];
        }
        if (rc) break;
        u[cur] += minVal;
        for (int64_t r = 0; r < nr; r++)
            if (SR[r] && r != cur) u[r] += minVal - spc[col4row[r]];
        for (int64_t j = 0; j < nc; j++)
            if (SC[j]) v[j] -= minVal - spc[j];
        int64_t j = sink;
        for (;;) {
            int64_t r = path[j];
            row4col[j] = r;
            int64_t t = col4row[r]; col4row[r] = j; j = t;
            if (r == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* rows of the transposed problem are original columns: emit sorted by
             * original row (= col4row value), i.e. argsort(col4row). */
            int64_t n = 0;
            for (int64_t c = 0; c < nc; c++)          /* nc = original nr */
                if (row4col[c] != -1) { a[n] = c; b[n] = row4col[c]; n++; }
        } else {
            for (int64_t r = 0; r < nr; r++) { a[r] = r; b[r] = col4row[r]; }
        }
    }
    free(temp); free(u); free(v); free(spc); free(path); free(col4row); free(row4col); free(remaining);
    free(SR); free(SC);
    return rc;
}

/* Batched entry: fp32 cost blocks [B][Q][ldk] (only the first K[b] columns of
 * clip b are used), converted to fp64 exactly as scipy does on input.
 * rows/cols: [B][min(Q,ldk)] int64, counts[b] = min(Q, K[b]). */
int lsap_oracle_batched_f32(int64_t B, int64_t Q, int64_t ldk, const float *cost, const int32_t *K,
                            int64_t *rows, int64_t *cols, int32_t *counts)
{
    int64_t cap = Q < ldk ? Q : ldk;
    double *c64 = (double *)malloc(sizeof(double) * (Q * ldk + 1));
    for (int64_t bidx = 0; bidx < B; bidx++) {
        int64_t k = K[bidx];
        for (int64_t i = 0; i < Q; i++)
            for (int64_t j = 0; j < k; j++) c64[i * k + j] = (double)cost[(bidx * Q + i) * ldk + j];
        int rc = lsap_solve_f64(Q, k, c64, rows + bidx * cap, cols + bidx * cap);
        if (rc) { free(c64); return rc; }
        counts[bidx] = (int32_t)(Q < k ? Q : k);
    }
    free(c64);
    return 0;
}

int lsap_oracle_f64(int64_t nr, int64_t nc, const double *cost, int64_t *a, int64_t *b)
{
    return lsap_solve_f64(nr, nc, cost, a, b);
}
