/* TEST INFRASTRUCTURE ONLY — CPU oracle for the matcher's assignment step.
 *
 * Restatement of the published algorithm behind scipy.optimize.linear_sum_assignment
 * (module scipy.optimize._lsap, C++ `rectangular_lsap`; D. F. Crouse, "On implementing 2D
 * rectangular assignment algorithms", IEEE TAES 52(4), 2016 — shortest augmenting path with
 * dual updates).  SciPy is a third-party dependency of the reference (requirements.txt:22
 * pins scipy==1.15.1; 1.18.1 is installed here), its source is not under /root/reference;
 * the reference call sites are src/d_fine/matcher.py:14,243.  Parity is pinned in
 * tests/test_oracle_lsap.py against the installed SciPy on random, tie-heavy and
 * rectangular matrices plus the known-answer vectors recorded in SURVEY.md §8c.
 *
 * Behaviour that matters for bit-exact indices:
 *   - costs are promoted to double;
 *   - if n_rows > n_cols the transpose is solved;
 *   - rows are inserted in order, candidate columns are scanned from a reverse-filled
 *     `remaining` list, ties on the reduced cost prefer an unassigned column;
 *   - output pairs are sorted by original row index.
 *
 * int lsap_solve(nr, nc, cost[nr*nc] row-major double, rows_out[min], cols_out[min])
 *   returns 0 ok, -1 infeasible, -2 invalid entry (NaN / -inf).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

static long augment(long nc, const double *cost, const double *u, const double *v, long *path,
                    const long *row4col, double *spc, long i, char *SR, char *SC, long *remaining,
                    double *p_min)
{
    double min_val = 0.0;
    long n_rem = nc;
    for (long it = 0; it < nc; it++) remaining[it] = nc - it - 1;
    memset(SC, 0, (size_t)nc);
    for (long j = 0; j < nc; j++) spc[j] = INFINITY;

    long sink = -1;
    while (sink == -1) {
        long index = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (long it = 0; it < n_rem; it++) {
            long j = remaining[it];
            double r = min_val + cost[i * nc + j] - u[i] - v[j];
            if (r < spc[j]) { path[j] = i; spc[j] = r; }
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        long j = remaining[index];
        if (row4col[j] == -1) sink = j; else i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[--n_rem];
    }
    *p_min = min_val;
    return sink;
}

static int cmp_idx(const void *a, const void *b, void *key)
{
    const long *k = (const long *)key;
    long x = k[*(const long *)a], y = k[*(const long *)b];
    return (x > y) - (x < y);
}

int lsap_solve(long nr, long nc, const double *cost_in, long *rows_out, long *cols_out)
{
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    double *cost = (double *)malloc(sizeof(double) * (size_t)(nr * nc));
    if (transpose) {
        for (long i = 0; i < nr; i++)
            for (long j = 0; j < nc; j++) cost[j * nr + i] = cost_in[i * nc + j];
        long t = nr; nr = nc; nc = t;
    } else {
        memcpy(cost, cost_in, sizeof(double) * (size_t)(nr * nc));
    }
    for (long k = 0; k < nr * nc; k++)
        if (isnan(cost[k]) || cost[k] == -INFINITY) { free(cost); return -2; }

    double *u = (double *)calloc((size_t)nr, sizeof(double));
    double *v = (double *)calloc((size_t)nc, sizeof(double));
    double *spc = (double *)malloc(sizeof(double) * (size_t)nc);
    long *path = (long *)malloc(sizeof(long) * (size_t)nc);
    long *col4row = (long *)malloc(sizeof(long) * (size_t)nr);
    long *row4col = (long *)malloc(sizeof(long) * (size_t)nc);
    long *remaining = (long *)malloc(sizeof(long) * (size_t)nc);
    char *SR = (char *)malloc((size_t)nr);
    char *SC = (char *)malloc((size_t)nc);
    for (long j = 0; j < nc; j++) { path[j] = -1; row4col[j] = -1; }
    for (long i = 0; i < nr; i++) col4row[i] = -1;

    int rc = 0;
    for (long cur = 0; cur < nr; cur++) {
        double min_val;
        memset(SR, 0, (size_t)nr);
        long sink = augment(nc, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, &min_val);
        if (sink < 0) { rc = -1; break; }
        u[cur] += min_val;
        for (long i = 0; i < nr; i++)
            if (SR[i] && i != cur) u[i] += min_val - spc[col4row[i]];
        for (long j = 0; j < nc; j++)
            if (SC[j]) v[j] -= min_val - spc[j];
        long j = sink;
        for (;;) {
            long i = path[j];
            row4col[j] = i;
            long t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            long *order = (long *)malloc(sizeof(long) * (size_t)nr);
            for (long i = 0; i < nr; i++) order[i] = i;
            qsort_r(order, (size_t)nr, sizeof(long), cmp_idx, col4row);
            for (long k = 0; k < nr; k++) { rows_out[k] = col4row[order[k]]; cols_out[k] = order[k]; }
            free(order);
        } else {
            for (long i = 0; i < nr; i++) { rows_out[i] = i; cols_out[i] = col4row[i]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(remaining); free(SR); free(SC);
    return rc;
}
