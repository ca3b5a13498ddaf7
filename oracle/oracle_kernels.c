/*
 * oracle_kernels.c -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C restatements of the two third-party algorithms the reference's
 * aggregation stage calls but does not vendor:
 *
 *   oracle_nms  : torchvision.ops.nms, CPU kernel semantics (torchvision 0.26.0,
 *                 requirements.txt:8 "torchvision>=0.10.0", unpinned).  Call sites in
 *                 the reference: yolox/models/tscd_head.py:1630,
 *                 yolox/models/post_process.py:58,73,510 (all through
 *                 torchvision.ops.batched_nms -> coordinate trick -> nms).
 *   oracle_lap  : scipy.optimize.linear_sum_assignment (scipy 1.18.1; imported at
 *                 yolox/models/tscd_matching.py:7, called at :935), i.e. the
 *                 shortest-augmenting-path rectangular LSAP solver of Crouse (2016)
 *                 in the form SciPy publishes it.
 *
 * Neither source is present under /root/reference; the published algorithms are
 * restated here and pinned black-box against the installed binaries by
 * tests/test_oracle_thirdparty.py.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file.
 *
 * Build: gcc -O2 -fPIC -shared -ffp-contract=off -o _build/liboracle.so oracle_kernels.c -lm
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* stable argsort, descending score, ties -> lower index first (torch.sort    */
/* stable=True, descending=True as used by torchvision's CPU nms kernel).     */
static void stable_argsort_desc(const float* key, int64_t n, int64_t* order, int64_t* tmp) {
    for (int64_t i = 0; i < n; ++i) order[i] = i;
    for (int64_t width = 1; width < n; width *= 2) {
        for (int64_t lo = 0; lo < n; lo += 2 * width) {
            int64_t mid = lo + width < n ? lo + width : n;
            int64_t hi = lo + 2 * width < n ? lo + 2 * width : n;
            int64_t a = lo, b = mid, o = lo;
            while (a < mid && b < hi) {
                /* take from the right run only if strictly greater: stability */
                if (key[order[b]] > key[order[a]]) tmp[o++] = order[b++];
                else tmp[o++] = order[a++];
            }
            while (a < mid) tmp[o++] = order[a++];
            while (b < hi) tmp[o++] = order[b++];
        }
        memcpy(order, tmp, (size_t)n * sizeof(int64_t));
    }
}

/*
 * Greedy NMS.  boxes: [n,4] float32 xyxy; scores: [n] float32.
 * keep (out): indices into boxes, descending-score order. Returns num kept.
 * Arithmetic is single precision with no FMA contraction; the IoU test is
 * `inter / (area_i + area_j - inter) > thr` with the comparison in double, as
 * the CPU kernel compares a float against the double `iou_threshold`.
 */
int64_t oracle_nms(const float* boxes, const float* scores, int64_t n, double thr, int64_t* keep) {
    if (n <= 0) return 0;
    int64_t* order = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    int64_t* tmp = (int64_t*)malloc((size_t)n * sizeof(int64_t));
    float* area = (float*)malloc((size_t)n * sizeof(float));
    unsigned char* dead = (unsigned char*)calloc((size_t)n, 1);
    stable_argsort_desc(scores, n, order, tmp);
    for (int64_t i = 0; i < n; ++i) {
        volatile float w = boxes[4 * i + 2] - boxes[4 * i + 0];
        volatile float h = boxes[4 * i + 3] - boxes[4 * i + 1];
        area[i] = w * h;
    }
    int64_t nk = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i];
        if (dead[i]) continue;
        keep[nk++] = i;
        float ix1 = boxes[4 * i], iy1 = boxes[4 * i + 1], ix2 = boxes[4 * i + 2], iy2 = boxes[4 * i + 3];
        float ia = area[i];
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (dead[j]) continue;
            float xx1 = ix1 > boxes[4 * j] ? ix1 : boxes[4 * j];
            float yy1 = iy1 > boxes[4 * j + 1] ? iy1 : boxes[4 * j + 1];
            float xx2 = ix2 < boxes[4 * j + 2] ? ix2 : boxes[4 * j + 2];
            float yy2 = iy2 < boxes[4 * j + 3] ? iy2 : boxes[4 * j + 3];
            volatile float w = xx2 - xx1;
            volatile float h = yy2 - yy1;
            float w0 = w > 0.f ? w : 0.f;
            float h0 = h > 0.f ? h : 0.f;
            volatile float inter = w0 * h0;
            volatile float uni = ia + area[j];
            uni = uni - inter;
            volatile float ovr = inter / uni;
            if ((double)ovr > thr) dead[j] = 1;
        }
    }
    free(order); free(tmp); free(area); free(dead);
    return nk;
}

/* ------------------------------------------------------------------------- */
/* Rectangular linear sum assignment (shortest augmenting path).              */
/* cost: [nr,nc] row-major double.  a,b (out): min(nr,nc) matched pairs with  */
/* `a` ascending.  Returns 0, or -1 if infeasible / invalid entries.          */
static int64_t augmenting_path(int64_t nc, const double* cost, const double* u, const double* v,
                               int64_t* path, const int64_t* row4col, double* spc, int64_t i,
                               unsigned char* SR, unsigned char* SC, int64_t* remaining,
                               int64_t nr, double* p_min) {
    double min_val = 0.0;
    int64_t num_remaining = nc;
    for (int64_t it = 0; it < nc; ++it) remaining[it] = nc - it - 1; /* reverse fill */
    memset(SR, 0, (size_t)nr);
    memset(SC, 0, (size_t)nc);
    for (int64_t j = 0; j < nc; ++j) spc[j] = INFINITY;
    int64_t sink = -1;
    while (sink == -1) {
        int64_t index = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int64_t it = 0; it < num_remaining; ++it) {
            int64_t j = remaining[it];
            double r = min_val + cost[i * nc + j] - u[i] - v[j];
            if (r < spc[j]) { path[j] = i; spc[j] = r; }
            /* among equal minima prefer one that is a new sink */
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) { lowest = spc[j]; index = it; }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1;
        int64_t j = remaining[index];
        if (row4col[j] == -1) sink = j; else i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[--num_remaining];
    }
    *p_min = min_val;
    return sink;
}

int oracle_lap(const double* cost_in, int64_t nr, int64_t nc, int64_t* a, int64_t* b) {
    if (nr == 0 || nc == 0) return 0;
    int transpose = nc < nr;
    double* cost = (double*)malloc((size_t)(nr * nc) * sizeof(double));
    if (transpose) {
        for (int64_t i = 0; i < nr; ++i)
            for (int64_t j = 0; j < nc; ++j) cost[j * nr + i] = cost_in[i * nc + j];
        int64_t t = nr; nr = nc; nc = t;
    } else {
        memcpy(cost, cost_in, (size_t)(nr * nc) * sizeof(double));
    }
    for (int64_t k = 0; k < nr * nc; ++k)
        if (cost[k] != cost[k] || cost[k] == -INFINITY) { free(cost); return -1; }

    double* u = (double*)calloc((size_t)nr, sizeof(double));
    double* v = (double*)calloc((size_t)nc, sizeof(double));
    double* spc = (double*)malloc((size_t)nc * sizeof(double));
    int64_t* path = (int64_t*)malloc((size_t)nc * sizeof(int64_t));
    int64_t* col4row = (int64_t*)malloc((size_t)nr * sizeof(int64_t));
    int64_t* row4col = (int64_t*)malloc((size_t)nc * sizeof(int64_t));
    unsigned char* SR = (unsigned char*)malloc((size_t)nr);
    unsigned char* SC = (unsigned char*)malloc((size_t)nc);
    int64_t* remaining = (int64_t*)malloc((size_t)nc * sizeof(int64_t));
    for (int64_t j = 0; j < nc; ++j) { path[j] = -1; row4col[j] = -1; }
    for (int64_t i = 0; i < nr; ++i) col4row[i] = -1;

    int rc = 0;
    for (int64_t cur = 0; cur < nr; ++cur) {
        double min_val;
        int64_t sink = augmenting_path(nc, cost, u, v, path, row4col, spc, cur, SR, SC, remaining, nr, &min_val);
        if (sink < 0) { rc = -1; break; }
        u[cur] += min_val;
        for (int64_t i = 0; i < nr; ++i)
            if (SR[i] && i != cur) u[i] += min_val - spc[col4row[i]];
        for (int64_t j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= min_val - spc[j];
        int64_t j = sink;
        for (;;) {
            int64_t i = path[j];
            row4col[j] = i;
            int64_t t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    if (rc == 0) {
        if (transpose) {
            /* rows of the transposed problem are original columns: emit sorted by original row */
            int64_t k = 0;
            for (int64_t orig_row = 0; orig_row < nc; ++orig_row) {
                int64_t i = row4col[orig_row];
                if (i >= 0) { a[k] = orig_row; b[k] = i; ++k; }
            }
        } else {
            for (int64_t i = 0; i < nr; ++i) { a[i] = i; b[i] = col4row[i]; }
        }
    }
    free(cost); free(u); free(v); free(spc); free(path); free(col4row); free(row4col);
    free(SR); free(SC); free(remaining);
    return rc;
}
