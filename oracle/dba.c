/* CPU oracle for the DTW-barycentre-averaging step.  TEST INFRASTRUCTURE ONLY.
 *
 * Plain C restatement of the two DBA algorithms the reference touches; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline leg may load the library built from this
 * file (oracle/dba.py).  Build: gcc -O2 -ffp-contract=off -shared -fPIC (see oracle/dba.py) --
 * contraction is OFF so every product and sum rounds exactly as the NumPy/numba code it restates.
 *
 * (1) be_oracle_dba_subgradient: tslearn 0.5.1.0 `dtw_barycenter_averaging_subgradient`
 *     (pinned at requirements.txt:185; called at ensembles/models.py:176-178 and :251-253 with
 *     max_iter=50, tol=1e-3).  tslearn is NOT vendored in the reference and not installable here,
 *     so this restates its published algorithm (tslearn/barycenters/dba.py: _init_avg,
 *     _mm_assignment, _subgradient_valence_warping, _subgradient_update_barycenter;
 *     tslearn/metrics/dtw_variants.py: njit_accumulated_matrix, _return_path) => PARITY UNPINNED
 *     against tslearn itself.  The DTW recursion and the path machinery underneath it ARE pinned:
 *     see (2).
 * (2) be_oracle_perform_dba / be_oracle_squared_dtw: the reference's own NumPy implementation,
 *     ensembles/dtwa.py:6-143 (exported at ensembles/__init__.py:3).  That file imports only
 *     NumPy and runs in this container, so tests/golden/make_golden_dba.py executes it and
 *     commits its outputs (tests/golden/dba_reference.npz): PINNED.
 *
 * Summation orders are part of the definition (the CUDA path reproduces them):
 *   - per series k and barycentre index i, wx_k[i] = sum of x_k[j] over the path cells (i, j) in
 *     BACKTRACK order (j decreasing), starting from 0.0; v_k[i] = number of such cells;
 *   - subgradient: delta[i] = ((((0 + v_0 c_i) - wx_0) + v_1 c_i) - wx_1) ...  (dba.py's loop over k);
 *   - perform_dba: centre[i] = (wx_0[i] + wx_1[i] + ...) / (v_0[i] + v_1[i] + ...), k increasing
 *     (dtwa.py accumulates one running sum across series; this groups it per series, which is
 *     a <= 1e-15 relative re-association -- the golden test states that tolerance).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TIE_TSLEARN 0 /* argmin([diag, top, left]): first minimum wins (dtw_variants.py:_return_path) */
#define TIE_DTWA 1    /* ensembles/dtwa.py:62-74 / :113-129: diag<=left ? (diag<=top ? diag : top) : (left<=top ? left : top) */

/* cum is (T1+1) x (T2+1), row-major; returns cum[T1][T2] (the SQUARED DTW distance).
 * tslearn: njit_accumulated_matrix (inf border, cum[0][0] = 0); dtwa.py:49-75 builds the same
 * table without the border (first row / column are running sums), which is what the inf border
 * gives. dir (T1 x T2, may be NULL): 0 diag, 1 top (i-1, j), 2 left (i, j-1), 3 = origin. */
static double dtw_table(const double* a, int T1, const double* x, int T2, double* cum, uint8_t* dir, int tie) {
    const int ld = T2 + 1;
    for (int j = 0; j <= T2; ++j) cum[j] = INFINITY;
    for (int i = 1; i <= T1; ++i) cum[(size_t)i * ld] = INFINITY;
    cum[0] = 0.0;
    for (int i = 0; i < T1; ++i) {
        const double ai = a[i];
        const double* up = cum + (size_t)i * ld;
        double* cur = cum + (size_t)(i + 1) * ld;
        for (int j = 0; j < T2; ++j) {
            const double diff = ai - x[j];
            const double d = diff * diff;
            const double diag = up[j], top = up[j + 1], left = cur[j];
            double m;
            int code;
            if (tie == TIE_TSLEARN) {
                m = diag; code = 0;
                if (top < m) { m = top; code = 1; }
                if (left < m) { m = left; code = 2; }
            } else {
                if (diag <= left) {
                    if (diag <= top) { m = diag; code = 0; } else { m = top; code = 1; }
                } else {
                    if (left <= top) { m = left; code = 2; } else { m = top; code = 1; }
                }
            }
            cur[j + 1] = m + d;
            if (dir) dir[(size_t)i * T2 + j] = (uint8_t)((i == 0 && j == 0) ? 3 : code);
        }
    }
    return cum[(size_t)T1 * ld + T2];
}

/* Backtrack from (T1-1, T2-1) (tslearn _return_path / dtwa.py:131-139): v[i], wx[i] as defined above. */
static void backtrack(const uint8_t* dir, int T1, int T2, const double* x, double* v, double* wx) {
    for (int i = 0; i < T1; ++i) { v[i] = 0.0; wx[i] = 0.0; }
    int i = T1 - 1, j = T2 - 1;
    for (;;) {
        v[i] += 1.0;
        wx[i] += x[j];
        const int code = dir[(size_t)i * T2 + j];
        if (code == 3) break;
        int ni = i - (code != 2), nj = j - (code != 1);
        /* a NaN table (NaN inputs) can record "diag" on a border; keep the walk inside the table */
        if (ni < 0) { ni = 0; nj = j - 1; }
        if (nj < 0) { nj = 0; ni = i - 1; }
        i = ni; j = nj;
    }
}

double be_oracle_squared_dtw(const double* s, int T1, const double* t, int T2, int tie) {
    double* cum = (double*)malloc(sizeof(double) * (size_t)(T1 + 1) * (T2 + 1));
    const double r = dtw_table(s, T1, t, T2, cum, NULL, tie);
    free(cum);
    return r;
}

/* path of dtw(s, t): writes the (i, j) pairs from (0,0) to the end into path_ij [2 * (T1+T2)], returns the length */
int be_oracle_dtw_path(const double* s, int T1, const double* t, int T2, int tie, int* path_ij, double* sq_cost) {
    double* cum = (double*)malloc(sizeof(double) * (size_t)(T1 + 1) * (T2 + 1));
    uint8_t* dir = (uint8_t*)malloc((size_t)T1 * T2);
    *sq_cost = dtw_table(s, T1, t, T2, cum, dir, tie);
    int n = 0, i = T1 - 1, j = T2 - 1;
    for (;;) {
        path_ij[2 * n] = i; path_ij[2 * n + 1] = j; ++n;
        const int code = dir[(size_t)i * T2 + j];
        if (code == 3) break;
        int ni = i - (code != 2), nj = j - (code != 1);
        if (ni < 0) { ni = 0; nj = j - 1; }
        if (nj < 0) { nj = 0; ni = i - 1; }
        i = ni; j = nj;
    }
    for (int a = 0, b = n - 1; a < b; ++a, --b) {
        int t0 = path_ij[2 * a], t1 = path_ij[2 * a + 1];
        path_ij[2 * a] = path_ij[2 * b]; path_ij[2 * a + 1] = path_ij[2 * b + 1];
        path_ij[2 * b] = t0; path_ij[2 * b + 1] = t1;
    }
    free(cum); free(dir);
    return n;
}

/* tslearn 0.5.1.0 dtw_barycenter_averaging_subgradient(X, max_iter, initial_step_size, final_step_size, tol),
 * weights = None (all ones), barycenter_size = None, metric_params = None, X [R, T] (d = 1).
 * init [T] or NULL (= _init_avg: the mean over series when barycenter_size == T).
 * Returns the number of iterations run; *cost_out = the last cost evaluated (cost of the barycentre
 * BEFORE the last update, as in the original loop). */
int be_oracle_dba_subgradient(const double* X, int R, int T, int max_iter, double initial_step_size,
                              double final_step_size, double tol, const double* init, double* bary,
                              double* cost_out) {
    double* cum = (double*)malloc(sizeof(double) * (size_t)(T + 1) * (T + 1));
    uint8_t* dir = (uint8_t*)malloc((size_t)T * T);
    double* v = (double*)malloc(sizeof(double) * (size_t)R * T);
    double* wx = (double*)malloc(sizeof(double) * (size_t)R * T);
    if (init) {
        memcpy(bary, init, sizeof(double) * T);
    } else { /* numpy.nanmean(X_, axis=0): pairwise summation degenerates to sequential for R < 8 rows
                (axis-0 reduction adds row by row), then divides by the count */
        for (int i = 0; i < T; ++i) {
            double s = 0.0;
            for (int k = 0; k < R; ++k) s += X[(size_t)k * T + i];
            bary[i] = s / (double)R;
        }
    }
    double cost_prev = INFINITY, cost = INFINITY, eta = initial_step_size;
    int it = 0;
    for (; it < max_iter;) {
        /* _mm_assignment */
        cost = 0.0;
        for (int k = 0; k < R; ++k) {
            const double* xk = X + (size_t)k * T;
            const double dist = sqrt(dtw_table(bary, T, xk, T, cum, dir, TIE_TSLEARN));
            cost += dist * dist * 1.0;
            backtrack(dir, T, T, xk, v + (size_t)k * T, wx + (size_t)k * T);
        }
        cost /= (double)R;
        /* _subgradient_valence_warping + _subgradient_update_barycenter */
        const double step = 2.0 * eta / (double)R;
        for (int i = 0; i < T; ++i) {
            double delta = 0.0;
            for (int k = 0; k < R; ++k) {
                delta += v[(size_t)k * T + i] * bary[i];
                delta -= wx[(size_t)k * T + i];
            }
            bary[i] -= step * delta;
        }
        eta -= (initial_step_size - final_step_size) / (double)max_iter;
        ++it;
        if (fabs(cost_prev - cost) < tol) break;
        else if (cost_prev < cost) { /* tslearn warns "DBA loss is increasing while it should not be." and goes on */ }
        else cost_prev = cost;
    }
    if (cost_out) *cost_out = cost;
    free(cum); free(dir); free(v); free(wx);
    return it;
}

/* ensembles/dtwa.py:6-20 performDBA(series, n_iterations) for len(series) <= 50 equal-length series
 * (more than 50 takes an unseeded random subset of medoid candidates, dtwa.py:26).
 * X [R, T] -> center [T]; returns the medoid index (dtwa.py:23-37). */
int be_oracle_perform_dba(const double* X, int R, int T, int n_iterations, double* center) {
    double* cum = (double*)malloc(sizeof(double) * (size_t)(T + 1) * (T + 1));
    uint8_t* dir = (uint8_t*)malloc((size_t)T * T);
    double* v = (double*)malloc(sizeof(double) * (size_t)T);
    double* wx = (double*)malloc(sizeof(double) * (size_t)T);
    double* sum = (double*)malloc(sizeof(double) * (size_t)T);
    double* cnt = (double*)malloc(sizeof(double) * (size_t)T);
    int medoid = -1;
    double best = 1e20;
    for (int c = 0; c < R; ++c) {
        double ss = 0.0; /* Python sum(map(...)): sequential from 0 */
        for (int k = 0; k < R; ++k) ss += dtw_table(X + (size_t)c * T, T, X + (size_t)k * T, T, cum, NULL, TIE_DTWA);
        if (medoid == -1 || ss < best) { best = ss; medoid = c; }
    }
    memcpy(center, X + (size_t)medoid * T, sizeof(double) * T);
    for (int it = 0; it < n_iterations; ++it) {
        for (int i = 0; i < T; ++i) { sum[i] = 0.0; cnt[i] = 0.0; }
        for (int k = 0; k < R; ++k) {
            const double* xk = X + (size_t)k * T;
            dtw_table(center, T, xk, T, cum, dir, TIE_DTWA);
            backtrack(dir, T, T, xk, v, wx);
            for (int i = 0; i < T; ++i) { sum[i] += wx[i]; cnt[i] += v[i]; }
        }
        for (int i = 0; i < T; ++i) center[i] = sum[i] / cnt[i];
    }
    free(cum); free(dir); free(v); free(wx); free(sum); free(cnt);
    return medoid;
}
