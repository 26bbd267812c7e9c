/*
 * spmv_oracle.c -- CPU oracle (TEST INFRASTRUCTURE, see spmv_oracle.h).
 *
 * Plain-C restatement of the reference's Matrix Market reader, format
 * converters, row partition and SpMV loops.  Citations are into
 * /root/reference/src unless noted.  Parity: PINNED (see header).
 */
#include "spmv_oracle.h"

#include <ctype.h>
#include <errno.h>
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

static __thread char orc_err[512];

const char *orc_last_error(void) { return orc_err; }

static int orc_fail(int code, const char *msg)
{
    snprintf(orc_err, sizeof orc_err, "%s", msg);
    return code;
}

/* ======================================================================== */
/* Matrix Market                                                             */
/* ======================================================================== */

typedef struct { const char *p, *end; } cursor;

static void skip_ws(cursor *c)
{
    while (c->p < c->end && isspace((unsigned char)*c->p)) c->p++;
}

/* One line without its terminator (std::getline, matrix-market.cpp:416-417). */
static int get_line(cursor *c, const char **b, const char **e)
{
    if (c->p >= c->end) return 0;
    *b = c->p;
    while (c->p < c->end && *c->p != '\n') c->p++;
    *e = c->p;
    if (c->p < c->end) c->p++;
    return 1;
}

static int next_word(cursor *c, char *buf, size_t cap, int lower)
{
    size_t n = 0;
    skip_ws(c);
    while (c->p < c->end && !isspace((unsigned char)*c->p)) {
        char ch = *c->p++;
        /* asciitolower (matrix-market.cpp:339-344) */
        if (lower && ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        if (n + 1 < cap) buf[n++] = ch;
    }
    buf[n] = 0;
    return n > 0;
}

/* operator>>(istream&, int32_t&): failure on overflow or missing digits. */
static int read_i32(cursor *c, int32_t *v)
{
    char tmp[64];
    size_t n = 0;
    char *endp;
    long long x;
    skip_ws(c);
    while (c->p + n < c->end && n + 1 < sizeof tmp &&
           (isdigit((unsigned char)c->p[n]) || ((c->p[n] == '-' || c->p[n] == '+') && n == 0)))
        n++;
    if (n == 0) return 0;
    memcpy(tmp, c->p, n);
    tmp[n] = 0;
    errno = 0;
    x = strtoll(tmp, &endp, 10);
    if (endp == tmp || errno == ERANGE || x > INT32_MAX || x < INT32_MIN) return 0;
    c->p += (endp - tmp);
    *v = (int32_t)x;
    return 1;
}

static int read_f64(cursor *c, double *v)
{
    char tmp[128];
    size_t n = 0;
    char *endp;
    skip_ws(c);
    while (c->p + n < c->end && n + 1 < sizeof tmp && !isspace((unsigned char)c->p[n])) n++;
    if (n == 0) return 0;
    memcpy(tmp, c->p, n);
    tmp[n] = 0;
    *v = strtod(tmp, &endp);
    if (endp == tmp) return 0;
    c->p += (endp - tmp);
    return 1;
}

int orc_mm_parse(const char *text, size_t len, orc_mm *out)
{
    cursor c = { text, text + len };
    cursor h;
    const char *lb, *le;
    char w[64];
    int32_t k;

    memset(out, 0, sizeof *out);
    orc_err[0] = 0;

    /* readHeader (matrix-market.cpp:414-436) */
    if (!get_line(&c, &lb, &le)) return orc_fail(ORC_ERR_PARSE, "Failed to parse header");
    h.p = lb; h.end = le;
    next_word(&h, w, sizeof w, 0);
    if (strcmp(w, "%%MatrixMarket") != 0)
        return orc_fail(ORC_ERR_PARSE, "Failed to parse header: Expected \"%%MatrixMarket\"");
    next_word(&h, w, sizeof w, 1); /* readObject :346-360 */
    if (strcmp(w, "matrix") != 0)
        return orc_fail(ORC_ERR_PARSE, "Failed to parse header: Expected \"matrix\"");
    next_word(&h, w, sizeof w, 1); /* readFormat :362-375 */
    if (strcmp(w, "coordinate") == 0) out->format = ORC_FORMAT_COORDINATE;
    else if (strcmp(w, "array") == 0) out->format = ORC_FORMAT_ARRAY;
    else return orc_fail(ORC_ERR_PARSE, "Expected \"coordinate\" or \"array\"");
    next_word(&h, w, sizeof w, 1); /* readField :377-394 */
    if (strcmp(w, "real") == 0) out->field = ORC_FIELD_REAL;
    else if (strcmp(w, "complex") == 0) out->field = ORC_FIELD_COMPLEX;
    else if (strcmp(w, "integer") == 0) out->field = ORC_FIELD_INTEGER;
    else if (strcmp(w, "pattern") == 0) out->field = ORC_FIELD_PATTERN;
    else return orc_fail(ORC_ERR_PARSE, "Expected \"real\", \"complex\", \"integer\", or \"pattern\"");
    next_word(&h, w, sizeof w, 1); /* readSymmetry :396-412 */
    if (strcmp(w, "general") == 0) out->symmetry = ORC_SYM_GENERAL;
    else if (strcmp(w, "symmetric") == 0) out->symmetry = ORC_SYM_SYMMETRIC;
    else if (strcmp(w, "skew-symmetric") == 0) out->symmetry = ORC_SYM_SKEW;
    else if (strcmp(w, "hermitian") == 0) out->symmetry = ORC_SYM_HERMITIAN;
    else return orc_fail(ORC_ERR_PARSE, "Expected \"general\", \"symmetric\", \"skew-symmetric\", or \"hermitian\"");

    /* readComments (matrix-market.cpp:438-447) */
    while (c.p < c.end && *c.p == '%') {
        get_line(&c, &lb, &le);
        out->num_comments++;
    }

    /* readSize (matrix-market.cpp:449-484) */
    if (!get_line(&c, &lb, &le)) return orc_fail(ORC_ERR_PARSE, "Failed to parse size");
    h.p = lb; h.end = le;
    if (!read_i32(&h, &out->rows))
        return orc_fail(ORC_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of rows");
    if (!read_i32(&h, &out->columns))
        return orc_fail(ORC_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of columns");
    if (out->format == ORC_FORMAT_ARRAY) {
        out->num_entries = 0;
        return ORC_OK;
    }
    if (!read_i32(&h, &out->num_entries))
        return orc_fail(ORC_ERR_OVERFLOW, "Failed to parse size: Integer overflow when reading number of non-zeros");
    if (out->num_entries < 0) return orc_fail(ORC_ERR_PARSE, "Failed to parse size");

    /* readEntries (matrix-market.cpp:508-528): whitespace separated records */
    out->i = (int32_t *)malloc(sizeof(int32_t) * (size_t)(out->num_entries + 1));
    out->j = (int32_t *)malloc(sizeof(int32_t) * (size_t)(out->num_entries + 1));
    out->a = (double *)malloc(sizeof(double) * (size_t)(out->num_entries + 1));
    if (!out->i || !out->j || !out->a) { orc_mm_free(out); return orc_fail(ORC_ERR_NOMEM, "out of memory"); }
    for (k = 0; k < out->num_entries; k++) {
        int ok = read_i32(&c, &out->i[k]) && read_i32(&c, &out->j[k]);
        if (ok) {
            switch (out->field) {
            case ORC_FIELD_REAL:
                ok = read_f64(&c, &out->a[k]);
                break;
            case ORC_FIELD_COMPLEX: { /* values_real(): real part (:253-258) */
                double im;
                ok = read_f64(&c, &out->a[k]) && read_f64(&c, &im);
                break;
            }
            case ORC_FIELD_INTEGER: { /* CoordinateEntryInteger::a is int (:53-58) */
                int32_t v;
                ok = read_i32(&c, &v);
                out->a[k] = (double)v;
                break;
            }
            default: /* pattern -> 1.0 (:267-272) */
                out->a[k] = 1.0;
            }
        }
        if (!ok) {
            char msg[128];
            snprintf(msg, sizeof msg, "Failed to parse entries: Expected %d entries, got %d entries.",
                     out->num_entries, k);
            orc_mm_free(out);
            return orc_fail(ORC_ERR_PARSE, msg);
        }
    }
    return ORC_OK;
}

void orc_mm_free(orc_mm *m)
{
    free(m->i); free(m->j); free(m->a);
    m->i = m->j = NULL; m->a = NULL;
}

void orc_row_lengths(int32_t rows, int32_t n, const int32_t *i, int32_t *len)
{
    int32_t k;
    memset(len, 0, sizeof(int32_t) * (size_t)rows);
    for (k = 0; k < n; k++) ++len[i[k] - 1];
}

int32_t orc_max_row_length(int32_t rows, int32_t n, const int32_t *i)
{
    int32_t *len = (int32_t *)calloc((size_t)rows + 1, sizeof(int32_t));
    int32_t r, m = 0;
    orc_row_lengths(rows, n, i, len);
    for (r = 0; r < rows; r++) if (len[r] > m) m = len[r];
    free(len);
    return m;
}

/* ---- stable two-key sort: bucket by primary key, order buckets by secondary */
typedef struct { int32_t key; int32_t idx; } keyidx;

static int cmp_keyidx(const void *pa, const void *pb)
{
    const keyidx *a = (const keyidx *)pa, *b = (const keyidx *)pb;
    if (a->key != b->key) return a->key < b->key ? -1 : 1;
    return a->idx < b->idx ? -1 : (a->idx > b->idx);
}

static void sort_two_keys(int32_t n, int32_t *prim, int32_t *sec, double *a)
{
    int32_t k, maxp = 0;
    int64_t *start;
    int32_t *p2, *s2;
    double *a2;
    if (n <= 1) return;
    for (k = 0; k < n; k++) if (prim[k] > maxp) maxp = prim[k];
    start = (int64_t *)calloc((size_t)maxp + 2, sizeof(int64_t));
    p2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    s2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    a2 = (double *)malloc(sizeof(double) * (size_t)n);
    for (k = 0; k < n; k++) start[prim[k] + 1]++;
    for (k = 0; k <= maxp; k++) start[k + 1] += start[k];
    {   /* stable scatter by primary key */
        int64_t *pos = (int64_t *)malloc(sizeof(int64_t) * ((size_t)maxp + 2));
        memcpy(pos, start, sizeof(int64_t) * ((size_t)maxp + 2));
        for (k = 0; k < n; k++) {
            int64_t d = pos[prim[k]]++;
            p2[d] = prim[k]; s2[d] = sec[k]; a2[d] = a[k];
        }
        free(pos);
    }
    for (k = 0; k <= maxp; k++) {
        int64_t b = start[k], e = start[k + 1], t;
        int sorted = 1;
        for (t = b + 1; t < e; t++) if (s2[t] < s2[t - 1]) { sorted = 0; break; }
        if (sorted) continue;
        if (e - b <= 48) { /* stable insertion sort */
            for (t = b + 1; t < e; t++) {
                int32_t sv = s2[t]; double av = a2[t];
                int64_t u = t - 1;
                while (u >= b && s2[u] > sv) { s2[u + 1] = s2[u]; a2[u + 1] = a2[u]; u--; }
                s2[u + 1] = sv; a2[u + 1] = av;
            }
        } else {
            int64_t m = e - b;
            keyidx *ki = (keyidx *)malloc(sizeof(keyidx) * (size_t)m);
            int32_t *st = (int32_t *)malloc(sizeof(int32_t) * (size_t)m);
            double *at = (double *)malloc(sizeof(double) * (size_t)m);
            for (t = 0; t < m; t++) { ki[t].key = s2[b + t]; ki[t].idx = (int32_t)t; }
            qsort(ki, (size_t)m, sizeof(keyidx), cmp_keyidx);
            for (t = 0; t < m; t++) { st[t] = s2[b + ki[t].idx]; at[t] = a2[b + ki[t].idx]; }
            memcpy(s2 + b, st, sizeof(int32_t) * (size_t)m);
            memcpy(a2 + b, at, sizeof(double) * (size_t)m);
            free(ki); free(st); free(at);
        }
    }
    memcpy(prim, p2, sizeof(int32_t) * (size_t)n);
    memcpy(sec, s2, sizeof(int32_t) * (size_t)n);
    memcpy(a, a2, sizeof(double) * (size_t)n);
    free(start); free(p2); free(s2); free(a2);
}

void orc_sort_row_major(int32_t n, int32_t *i, int32_t *j, double *a) { sort_two_keys(n, i, j, a); }
void orc_sort_column_major(int32_t n, int32_t *i, int32_t *j, double *a) { sort_two_keys(n, j, i, a); }

/* copy + sort_matrix_row_major, as every converter does first
 * (csr-matrix.cpp:201, ell-matrix.cpp:209, hybrid-matrix.cpp:363). */
static int sorted_copy(int32_t n, const int32_t *i, const int32_t *j, const double *a,
                       int32_t **si, int32_t **sj, double **sa)
{
    size_t m = (size_t)n + 1;
    *si = (int32_t *)malloc(sizeof(int32_t) * m);
    *sj = (int32_t *)malloc(sizeof(int32_t) * m);
    *sa = (double *)malloc(sizeof(double) * m);
    if (!*si || !*sj || !*sa) return orc_fail(ORC_ERR_NOMEM, "out of memory");
    if (n > 0) {
        memcpy(*si, i, sizeof(int32_t) * (size_t)n);
        memcpy(*sj, j, sizeof(int32_t) * (size_t)n);
        memcpy(*sa, a, sizeof(double) * (size_t)n);
    }
    orc_sort_row_major(n, *si, *sj, *sa);
    return ORC_OK;
}

static int check_entries(int32_t rows, int32_t columns, int32_t n, const int32_t *i, const int32_t *j)
{
    int32_t k;
    for (k = 0; k < n; k++)
        if (i[k] < 1 || i[k] > rows || j[k] < 1 || j[k] > columns)
            return orc_fail(ORC_ERR_FORMAT, "entry index outside the matrix");
    return ORC_OK;
}

/* ======================================================================== */
/* CSR                                                                       */
/* ======================================================================== */

int orc_csr_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t row_alignment, orc_csr *out)
{
    int32_t *si, *sj, r, k, l;
    double *sa;
    int rc;
    memset(out, 0, sizeof *out);
    if ((rc = check_entries(rows, columns, n, i, j))) return rc;
    if ((rc = sorted_copy(n, i, j, a, &si, &sj, &sa))) return rc;

    /* row lengths incl. alignment padding (csr-matrix.cpp:206-217) */
    out->row_ptr = (int32_t *)calloc((size_t)rows + 1, sizeof(int32_t));
    k = 0; l = 0;
    for (r = 0; r < rows; ++r) {
        while (l < n && si[l] - 1 == r) { l++; k++; }
        k = ((k + (row_alignment - 1)) / row_alignment) * row_alignment;
        out->row_ptr[r + 1] = k;
    }
    /* column indices and values, padding = (column 0, 0.0) (csr-matrix.cpp:219-237) */
    out->column_index = (int32_t *)calloc((size_t)out->row_ptr[rows] + 1, sizeof(int32_t));
    out->value = (double *)calloc((size_t)out->row_ptr[rows] + 1, sizeof(double));
    k = 0; l = 0;
    for (r = 0; r < rows; ++r) {
        while (l < n && si[l] - 1 == r) {
            out->column_index[k] = sj[l] - 1;
            out->value[k] = sa[l];
            ++k; ++l;
        }
        while (k < out->row_ptr[r + 1]) {
            out->column_index[k] = 0;
            out->value[k] = 0.0;
            ++k;
        }
    }
    out->rows = rows; out->columns = columns; out->num_entries = n;
    out->row_alignment = row_alignment;
    free(si); free(sj); free(sa);
    return ORC_OK;
}

void orc_csr_free(orc_csr *m)
{
    free(m->row_ptr); free(m->column_index); free(m->value);
    memset(m, 0, sizeof *m);
}

size_t orc_csr_size(const orc_csr *m)
{
    size_t stored = (size_t)m->row_ptr[m->rows];
    return sizeof(double) * stored + sizeof(int32_t) * ((size_t)m->rows + 1) + sizeof(int32_t) * stored;
}

static inline void csr_row(int32_t i, const int32_t *p, const int32_t *j, const double *a,
                           const double *x, double *y)
{
    /* csr_spmv_inner_loop (csr-matrix-spmv.cpp:21-33) */
    double z = 0.0;
    int32_t k;
    for (k = p[i]; k < p[i + 1]; ++k) z += a[k] * x[j[k]];
    y[i] += z;
}

void orc_csr_spmv(const orc_csr *A, const double *x, double *y)
{
    int32_t i;
    for (i = 0; i < A->rows; ++i) csr_row(i, A->row_ptr, A->column_index, A->value, x, y);
}

void orc_csr_spmv_omp(const orc_csr *A, const double *x, double *y, int num_threads)
{
    int32_t i;
    int32_t chunk = (A->rows + num_threads - 1) / (num_threads > 0 ? num_threads : 1);
    if (chunk < 1) chunk = 1;
    (void)chunk;
#pragma omp parallel for schedule(static, chunk) num_threads(num_threads)
    for (i = 0; i < A->rows; ++i) csr_row(i, A->row_ptr, A->column_index, A->value, x, y);
}

int32_t orc_csr_start_row(int32_t rows, int thread, int num_threads)
{
    /* csr-matrix.cpp:79-80 */
    int32_t rpt = (rows + num_threads - 1) / num_threads;
    int64_t s = (int64_t)thread * rpt;
    return (int32_t)(s < rows ? s : rows);
}

int32_t orc_csr_rows_per_thread(int32_t rows, int thread, int num_threads)
{
    return orc_csr_start_row(rows, thread + 1, num_threads) - orc_csr_start_row(rows, thread, num_threads);
}

int32_t orc_csr_nonzeros_per_thread(const int32_t *row_ptr, int32_t rows, int thread, int num_threads)
{
    /* csr-matrix.cpp:86-95 */
    return row_ptr[orc_csr_start_row(rows, thread + 1, num_threads)] -
           row_ptr[orc_csr_start_row(rows, thread, num_threads)];
}

void orc_csr_abs_rowsum(const orc_csr *A, const double *x, double *bound)
{
    int32_t i, k;
    for (i = 0; i < A->rows; ++i) {
        double s = 0.0;
        for (k = A->row_ptr[i]; k < A->row_ptr[i + 1]; ++k)
            s += fabs(A->value[k] * x[A->column_index[k]]);
        bound[i] = s;
    }
}

/* ======================================================================== */
/* COO                                                                       */
/* ======================================================================== */

int orc_coo_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a, orc_coo *out)
{
    int32_t k;
    int rc;
    memset(out, 0, sizeof *out);
    if ((rc = check_entries(rows, columns, n, i, j))) return rc;
    /* file order kept, 1-based -> 0-based (coo-matrix.cpp:226-239) */
    out->row_index = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    out->column_index = (int32_t *)malloc(sizeof(int32_t) * ((size_t)n + 1));
    out->value = (double *)malloc(sizeof(double) * ((size_t)n + 1));
    for (k = 0; k < n; k++) out->row_index[k] = i[k] - 1;
    for (k = 0; k < n; k++) out->column_index[k] = j[k] - 1;
    for (k = 0; k < n; k++) out->value[k] = a[k];
    out->rows = rows; out->columns = columns; out->num_entries = n;
    return ORC_OK;
}

void orc_coo_free(orc_coo *m)
{
    free(m->row_index); free(m->column_index); free(m->value);
    memset(m, 0, sizeof *m);
}

size_t orc_coo_size(const orc_coo *m)
{
    return (sizeof(double) + 2 * sizeof(int32_t)) * (size_t)m->num_entries;
}

/* coo_spmv (coo-matrix.cpp:248-285; same body in hybrid-matrix.cpp:491-528) */
static void coo_kernel(int T, int32_t rows, int32_t n, const int32_t *ri, const int32_t *ci,
                       const double *v, const double *x, double *y, double *ws, int32_t chunk)
{
    int32_t k, i;
    int t;
    if (T == 1) {
        for (k = 0; k < n; ++k) y[ri[k]] += v[k] * x[ci[k]];
        return;
    }
    if (chunk < 1) chunk = 1;
    /* "omp for schedule(static, chunk)": chunk c belongs to thread c % T */
    for (k = 0; k < n; ++k) {
        size_t thread = (size_t)((k / chunk) % T);
        ws[thread * (size_t)rows + ri[k]] += v[k] * x[ci[k]];
    }
    for (i = 0; i < rows; i++)
        for (t = 0; t < T; t++) y[i] += ws[(size_t)t * rows + i];
}

void orc_coo_spmv(int num_threads, const orc_coo *A, const double *x, double *y,
                  double *workspace, int32_t chunk_size)
{
    if (chunk_size <= 0) /* coo-matrix.cpp:321-323 */
        chunk_size = (A->num_entries + num_threads - 1) / num_threads;
    coo_kernel(num_threads, A->rows, A->num_entries, A->row_index, A->column_index, A->value,
               x, y, workspace, chunk_size);
}

void orc_coo_spmv_atomic(const orc_coo *A, const double *x, double *y)
{
    int32_t k;
    for (k = 0; k < A->num_entries; ++k)
        y[A->row_index[k]] += A->value[k] * x[A->column_index[k]];
}

/* Parallel port for the CPU baseline: same two loops, really threaded. */
static void coo_kernel_omp(int T, int32_t rows, int32_t n, const int32_t *ri, const int32_t *ci,
                           const double *v, const double *x, double *y, double *ws, int32_t chunk)
{
    if (T == 1) { coo_kernel(1, rows, n, ri, ci, v, x, y, ws, chunk); return; }
    if (chunk < 1) chunk = 1;
#pragma omp parallel num_threads(T)
    {
#ifdef _OPENMP
        size_t thread = (size_t)omp_get_thread_num();
#else
        size_t thread = 0;
#endif
        int32_t k, i;
        int t;
#pragma omp for schedule(static, chunk)
        for (k = 0; k < n; ++k) ws[thread * (size_t)rows + ri[k]] += v[k] * x[ci[k]];
#pragma omp for schedule(static, chunk)
        for (i = 0; i < rows; i++)
            for (t = 0; t < T; t++) y[i] += ws[(size_t)t * rows + i];
    }
}

void orc_coo_spmv_omp(const orc_coo *A, const double *x, double *y, double *workspace, int num_threads)
{
    int32_t chunk = (A->num_entries + num_threads - 1) / num_threads;
    coo_kernel_omp(num_threads, A->rows, A->num_entries, A->row_index, A->column_index, A->value,
                   x, y, workspace, chunk);
}

/* ======================================================================== */
/* ELLPACK                                                                   */
/* ======================================================================== */

int orc_ell_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t skip_padding, orc_ell *out, int32_t *first_row_empty)
{
    int32_t *si, *sj, r, k, l, W, slots;
    double *sa;
    int rc;
    memset(out, 0, sizeof *out);
    if (first_row_empty) *first_row_empty = 0;
    if ((rc = check_entries(rows, columns, n, i, j))) return rc;
    W = orc_max_row_length(rows, n, i); /* ell-matrix.cpp:199 */
    if (__builtin_mul_overflow(rows, W, &slots)) /* :200-205 */
        return orc_fail(ORC_ERR_OVERFLOW,
                        "Failed to convert to ELLPACK: Integer overflow when computing number of non-zeros");
    if ((rc = sorted_copy(n, i, j, a, &si, &sj, &sa))) return rc;
    out->column_index = (int32_t *)calloc((size_t)slots + 1, sizeof(int32_t));
    out->value = (double *)calloc((size_t)slots + 1, sizeof(double));
    k = 0; l = 0;
    for (r = 0; r < rows; ++r) { /* ell-matrix.cpp:218-234 */
        while (k < n && si[k] - 1 == r) {
            out->column_index[l] = sj[k] - 1;
            out->value[l] = sa[k];
            ++k; ++l;
        }
        while ((int64_t)l < (int64_t)(r + 1) * W) {
            int32_t pad;
            if (skip_padding) pad = INT32_MAX;
            else if (k > 0) pad = sj[k - 1] - 1;
            else { pad = 0; if (first_row_empty) *first_row_empty = 1; } /* reference: UB */
            out->column_index[l] = pad;
            out->value[l] = 0.0;
            ++l;
        }
    }
    out->rows = rows; out->columns = columns; out->num_entries = n;
    out->row_length = W; out->skip_padding = skip_padding;
    free(si); free(sj); free(sa);
    return ORC_OK;
}

void orc_ell_free(orc_ell *m)
{
    free(m->column_index); free(m->value);
    memset(m, 0, sizeof *m);
}

size_t orc_ell_size(const orc_ell *m)
{
    return (sizeof(double) + sizeof(int32_t)) * (size_t)m->rows * (size_t)m->row_length;
}

static inline void ell_row(int32_t i, int32_t W, const int32_t *j, const double *a,
                           const double *x, double *y, int skip)
{
    /* ell_spmv_inner_loop (ell-matrix.cpp:243-258), _skip_padding (:275-292) */
    double z = 0.0;
    int32_t l;
    for (l = 0; l < W; ++l) {
        size_t k = (size_t)i * W + l;
        if (skip && j[k] == INT32_MAX) break;
        z += a[k] * x[j[k]];
    }
    y[i] += z;
}

void orc_ell_spmv(const orc_ell *A, const double *x, double *y)
{
    int32_t i;
    for (i = 0; i < A->rows; ++i)
        ell_row(i, A->row_length, A->column_index, A->value, x, y, A->skip_padding);
}

void orc_ell_spmv_omp(const orc_ell *A, const double *x, double *y, int num_threads)
{
    int32_t i;
    int32_t chunk = (A->rows + num_threads - 1) / (num_threads > 0 ? num_threads : 1);
    if (chunk < 1) chunk = 1;
    (void)chunk;
#pragma omp parallel for schedule(static, chunk) num_threads(num_threads)
    for (i = 0; i < A->rows; ++i)
        ell_row(i, A->row_length, A->column_index, A->value, x, y, A->skip_padding);
}

/* ======================================================================== */
/* Hybrid                                                                    */
/* ======================================================================== */

int32_t orc_hyb_ell_row_length(int32_t rows, const int32_t *row_lengths)
{
    /* hybrid-matrix.cpp:329-344 */
    int32_t maxlen = 0, i, median = 0;
    int64_t below = 0;
    int64_t *hist;
    for (i = 0; i < rows; i++) if (row_lengths[i] > maxlen) maxlen = row_lengths[i];
    hist = (int64_t *)calloc((size_t)maxlen + 2, sizeof(int64_t));
    for (i = 0; i < rows; i++) hist[row_lengths[i]]++;
    while (below < (2 * (int64_t)rows) / 3) {
        below += hist[median];
        median++;
    }
    free(hist);
    return median == 0 ? 0 : median - 1;
}

int orc_hyb_from_entries(int32_t rows, int32_t columns, int32_t n,
                         const int32_t *i, const int32_t *j, const double *a,
                         int32_t skip_padding, orc_hyb *out)
{
    int32_t *si, *sj, *len, r, k, W, ne, nc, jj, slots;
    int64_t ncoo = 0;
    double *sa;
    int rc;
    memset(out, 0, sizeof *out);
    if ((rc = check_entries(rows, columns, n, i, j))) return rc;
    len = (int32_t *)calloc((size_t)rows + 1, sizeof(int32_t));
    orc_row_lengths(rows, n, i, len);
    W = orc_hyb_ell_row_length(rows, len);
    if (__builtin_mul_overflow(rows, W, &slots)) { /* hybrid-matrix.cpp:349-354 */
        free(len);
        return orc_fail(ORC_ERR_OVERFLOW,
                        "Failed to convert to HYBRID: Integer overflow when computing number of non-zeros");
    }
    for (r = 0; r < rows; r++) if (len[r] > W) ncoo += len[r] - W; /* :357-360 */
    if ((rc = sorted_copy(n, i, j, a, &si, &sj, &sa))) { free(len); return rc; }
    out->ell_column_index = (int32_t *)calloc((size_t)slots + 1, sizeof(int32_t));
    out->ell_value = (double *)calloc((size_t)slots + 1, sizeof(double));
    out->coo_row_index = (int32_t *)calloc((size_t)ncoo + 1, sizeof(int32_t));
    out->coo_column_index = (int32_t *)calloc((size_t)ncoo + 1, sizeof(int32_t));
    out->coo_value = (double *)calloc((size_t)ncoo + 1, sizeof(double));
    k = 0; ne = 0; nc = 0;
    for (r = 0; r < rows; ++r) { /* hybrid-matrix.cpp:378-410 */
        if (len[r] < W) {
            for (jj = 0; jj < len[r]; jj++) {
                out->ell_column_index[ne] = sj[k] - 1;
                out->ell_value[ne] = sa[k];
                ne++; k++;
            }
            for (jj = len[r]; jj < W; jj++) {
                out->ell_column_index[ne] = skip_padding ? INT32_MAX : (k > 0 ? sj[k - 1] - 1 : 0);
                out->ell_value[ne] = 0.0;
                ne++;
            }
        } else {
            for (jj = 0; jj < W; jj++) {
                out->ell_column_index[ne] = sj[k] - 1;
                out->ell_value[ne] = sa[k];
                ne++; k++;
            }
            for (jj = W; jj < len[r]; jj++) {
                out->coo_row_index[nc] = si[k] - 1;
                out->coo_column_index[nc] = sj[k] - 1;
                out->coo_value[nc] = sa[k];
                nc++; k++;
            }
        }
    }
    out->rows = rows; out->columns = columns; out->num_entries = n;
    out->ell_row_length = W; out->num_ell_entries = ne; out->ell_skip_padding = skip_padding;
    out->num_coo_entries = nc;
    free(len); free(si); free(sj); free(sa);
    return ORC_OK;
}

void orc_hyb_free(orc_hyb *m)
{
    free(m->ell_column_index); free(m->ell_value);
    free(m->coo_row_index); free(m->coo_column_index); free(m->coo_value);
    memset(m, 0, sizeof *m);
}

size_t orc_hyb_size(const orc_hyb *m)
{
    return 12u * (size_t)m->num_ell_entries + 16u * (size_t)m->num_coo_entries;
}

size_t orc_hyb_size_reference(const orc_hyb *m)
{
    /* value_size + index_size as written in hybrid-matrix.cpp:70-86 */
    return 8u * ((size_t)m->num_ell_entries + (size_t)m->num_coo_entries) +
           4u * ((size_t)m->num_ell_entries + (size_t)m->num_coo_entries);
}

void orc_hyb_spmv(int num_threads, const orc_hyb *A, const double *x, double *y, double *workspace)
{
    int32_t i;
    int32_t chunk = (A->rows + num_threads - 1) / num_threads; /* hybrid-matrix.cpp:543-545 */
    for (i = 0; i < A->rows; ++i)
        ell_row(i, A->ell_row_length, A->ell_column_index, A->ell_value, x, y, A->ell_skip_padding);
    coo_kernel(num_threads, A->rows, A->num_coo_entries, A->coo_row_index, A->coo_column_index,
               A->coo_value, x, y, workspace, chunk);
}

void orc_hyb_spmv_omp(const orc_hyb *A, const double *x, double *y, double *workspace, int num_threads)
{
    int32_t i;
    int32_t chunk = (A->rows + num_threads - 1) / num_threads;
    if (chunk < 1) chunk = 1;
#pragma omp parallel for schedule(static, chunk) num_threads(num_threads)
    for (i = 0; i < A->rows; ++i)
        ell_row(i, A->ell_row_length, A->ell_column_index, A->ell_value, x, y, A->ell_skip_padding);
    coo_kernel_omp(num_threads, A->rows, A->num_coo_entries, A->coo_row_index, A->coo_column_index,
                   A->coo_value, x, y, workspace, chunk);
}

/* ======================================================================== */
/* Row partitioners for the multi-GPU mode                                   */
/* ======================================================================== */

void orc_partition_rows_ref(int64_t rows, int P, int64_t *starts)
{
    int64_t rpt = (rows + P - 1) / P;
    int p;
    for (p = 0; p <= P; p++) {
        int64_t s = (int64_t)p * rpt;
        starts[p] = s < rows ? s : rows;
    }
}

void orc_partition_rows_nnz(int64_t rows, const int64_t *row_ptr, int P, int64_t *starts)
{
    int64_t nnz = row_ptr[rows];
    int p;
    starts[0] = 0;
    for (p = 1; p < P; p++) {
        /* floor(p*nnz/P) without overflow for nnz < 2^56 */
        int64_t target = (int64_t)(((__int128)nnz * p) / P);
        int64_t lo = 0, hi = rows; /* first r in [0, rows] with row_ptr[r] >= target */
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            if (row_ptr[mid] >= target) hi = mid; else lo = mid + 1;
        }
        starts[p] = lo;
    }
    starts[P] = rows;
}

/* P_weighted: start_p = first r with 1024*row_ptr[r] + w*r >= floor(p*(1024*nnz + w*rows)/P) (our design). */
void orc_partition_rows_weighted(int64_t rows, const int64_t *row_ptr, int P, int64_t w, int64_t *starts)
{
    const unsigned __int128 total = (unsigned __int128)1024 * (uint64_t)row_ptr[rows] + (unsigned __int128)w * (uint64_t)rows;
    int p;
    starts[0] = 0;
    for (p = 1; p < P; p++) {
        const unsigned __int128 target = total * (unsigned)p / (unsigned)P;
        int64_t lo = 0, hi = rows;
        while (lo < hi) {
            int64_t mid = lo + (hi - lo) / 2;
            const unsigned __int128 c = (unsigned __int128)1024 * (uint64_t)row_ptr[mid] + (unsigned __int128)w * (uint64_t)mid;
            if (c >= target) hi = mid; else lo = mid + 1;
        }
        starts[p] = lo;
    }
    starts[P] = rows;
}

/* find_new_order_GP after the METIS call (matrix/matrix-market-reorder.cpp:246-266): histogram of the parts, prefix
 * sum, permutation[new] = old in ascending old index inside a part, then new_order[old] = new.  (The call itself,
 * METIS_PartGraphKway :236-237, is third-party code absent from the reference tree and from this image: METIS 5, not
 * pinned by the reference -- there is nothing to restate; the product's partitioner is checked by properties.) */
int orc_order_from_parts(int32_t nvtxs, int32_t nparts, const int32_t *part, int32_t *new_order)
{
    int32_t i;
    int32_t *offset = (int32_t *)calloc((size_t)nparts + 1, sizeof(int32_t));
    int32_t *permutation = (int32_t *)malloc(((size_t)nvtxs + 1) * sizeof(int32_t));
    if (!offset || !permutation) { free(offset); free(permutation); return -1; }
    for (i = 0; i < nvtxs; i++) offset[part[i] + 1]++;
    offset[0] = 0;
    for (i = 1; i <= nparts; i++) offset[i] += offset[i - 1];
    for (i = 0; i < nvtxs; i++) {
        permutation[offset[part[i]]] = i;
        offset[part[i]]++;
    }
    for (i = 0; i < nvtxs; i++) new_order[i] = -1;
    for (i = 0; i < nvtxs; i++) new_order[permutation[i]] = i;
    free(offset);
    free(permutation);
    return 0;
}

/* ======================================================================== */
/* R-MAT edges of the synthetic power-law matrices (BASELINE configs 3, 4)   */
/* ======================================================================== */
/* Not a reference function: the reference has no generators.  This restates OUR device generator
 * (spmv_cache_trace_b200/csrc/generators.cu; numpy twin: oracle/generators_ref.py::rmat_entries) in C so
 * that the full-size parity tests can build their host matrix in seconds instead of minutes of numpy. */
static uint64_t orc_splitmix64(uint64_t z)
{
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* keys[e - e0] = (row << 32) | col of edge draw e, e in [e0, e1) */
void orc_rmat_keys(int scale, uint64_t seed, double a, double b, double c, uint64_t e0, uint64_t e1, uint64_t *keys)
{
    const double ta = a, tb = a + b, tc = a + b + c;
    int64_t k, n = (int64_t)(e1 - e0);
#pragma omp parallel for schedule(static)
    for (k = 0; k < n; ++k) {
        const uint64_t e = e0 + (uint64_t)k;
        uint64_t row = 0, col = 0;
        int level;
        for (level = 0; level < scale; ++level) {
            const uint64_t h = orc_splitmix64(seed + e * 0x9E3779B97F4A7C15ull + (uint64_t)level * 0xBF58476D1CE4E5B9ull);
            const double u = (double)(h >> 11) * 0x1.0p-53;
            const int q = u < ta ? 0 : (u < tb ? 1 : (u < tc ? 2 : 3));
            row = (row << 1) | (uint64_t)(q >> 1);
            col = (col << 1) | (uint64_t)(q & 1);
        }
        keys[k] = (row << 32) | col;
    }
}

/* From the sorted, de-duplicated keys: column index and value(i, j) = 2*u(hash(key)) - 1 of every entry, and the
 * number of entries of every row (rows + 1 counters, the last one unused). */
void orc_rmat_unpack(int64_t n, const uint64_t *keys, int32_t *col, double *val)
{
    int64_t k;
#pragma omp parallel for schedule(static)
    for (k = 0; k < n; ++k) {
        const uint64_t h = orc_splitmix64(keys[k] ^ 0xD1B54A32D192ED03ull);
        col[k] = (int32_t)(uint32_t)keys[k];
        val[k] = 2.0 * ((double)(h >> 11) * 0x1.0p-53) - 1.0;
    }
}
