/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's XR-Linear beam-search prediction.
 *
 * Nothing under pecos_b200/ may link, import or call this file; it is used by tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg as the checker.  Parity status: PINNED -- tests/test_oracle_cpu.py checks this
 * restatement bit-for-bit against oracle/_ref (the reference's own libpecos.cpp compiled unmodified) on the reference's
 * toy fixture (tests/golden) and on random trees (contiguous, permuted and pruned), for every post-processor.
 *
 * Plain C, single thread, works directly on the CSC weight / code matrices of each layer (no chunk layout), with the
 * arithmetic order of the reference's default BINARY_SEARCH_CHUNKED path.  Each function cites the code it restates
 * (paths relative to the reference checkout):
 *
 *   xlo_predict ................ HierarchicalMLModel::predict       pecos/core/xmc/inference.hpp:2446-2488
 *   xlo_predict_selected ....... HierarchicalMLModel::predict_on_selected_outputs  inference.hpp:2507-2571, :2129-2180, :1302-1358
 *   xlo_predict_from ........... c_xlinear_single_layer_predict_*   pecos/core/libpecos.cpp:201-235 (given previous beam)
 *   layer_predict .............. MLModel::predict_internal          pecos/core/xmc/inference.hpp:2029-2080
 *   candidates (prolongation) .. prolongate_predictions             pecos/core/xmc/inference.hpp:1155-1219
 *   score_sparse ............... chunk_ops<csr, bin_search>         pecos/core/xmc/inference.hpp:769-813 (+ :506-518)
 *   score_dense ................ chunk_ops<drm, bin_search>         pecos/core/xmc/inference.hpp:815-839
 *   transform / combine ........ PostProcessor<T>::get              pecos/core/xmc/inference.hpp:192-240, :1360-1384
 *   cmp_desc_then_pos .......... sorted_csr comparator              pecos/core/xmc/inference.hpp:1265-1273
 *   label mapping .............. rearrangement_t (perm_inv)         pecos/core/xmc/inference.hpp:1745-1784
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    uint32_t rows, cols;
    const uint64_t* col_ptr;
    const uint32_t* row_idx;
    const float* val;
} xlo_csc_t;

typedef struct {
    uint32_t rows, cols;
    const uint64_t* row_ptr; /* NULL => dense */
    const uint32_t* col_idx;
    const float* val;        /* csr values or row-major dense matrix */
} xlo_query_t;

enum { XLO_NOOP = 0, XLO_SIGMOID = 1, XLO_LOG_SIGMOID = 2, XLO_LP_HINGE = 3, XLO_LOG_LP_HINGE = 4 };

typedef struct {
    uint64_t* indptr;  /* rows + 1 */
    uint32_t* indices;
    float* data;
    uint64_t nnz;
    uint32_t rows, cols;
    /* algorithmic-byte counters of SURVEY.md 8(d), summed over layers: chunks, sum R (needs chunk view: not counted
       here), matched rows, entries, out cols, query nnz, beam out */
} xlo_result_t;

static float xlo_transform(float v, int kind, int p) {
    /* float/double promotion pattern of the reference lambdas (inference.hpp:208-238) */
    switch (kind) {
        case XLO_SIGMOID: return (float)(1.0 / (1.0 + expf(-v)));
        case XLO_LOG_SIGMOID: return (float)(-log(1.0 + expf(-v)));
        case XLO_LP_HINGE: {
            float z = (float)fmax(0.0, 1.0 - v);
            return (float)exp(-pow((double)z, (double)(size_t)p));
        }
        case XLO_LOG_LP_HINGE: {
            float z = (float)fmax(0.0, 1.0 - v);
            return (float)(-pow((double)z, (double)(size_t)p));
        }
        default: return v;
    }
}

static float xlo_combine(float x, float parent, int kind) {
    switch (kind) {
        case XLO_SIGMOID:
        case XLO_LP_HINGE: return x * parent;
        case XLO_LOG_SIGMOID:
        case XLO_LOG_LP_HINGE: return x + parent;
        default: return x;
    }
}

/* first position t in sorted idx[0..n) with idx[t] >= key (std::lower_bound) */
static uint32_t lower_bound_u32(const uint32_t* idx, uint32_t n, uint32_t key) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) {
        uint32_t mid = lo + (hi - lo) / 2;
        if (idx[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

/* One output column of a chunk against a sparse query: matched features in ascending index order, un-fused
 * multiply then add, bias row (feature index rows-1, only when bias > 0) last. */
static float score_sparse(const xlo_csc_t* W, uint32_t col, const uint32_t* qidx, const float* qval, uint32_t qn, float bias) {
    const uint64_t b = W->col_ptr[col], e = W->col_ptr[col + 1];
    const int use_bias = bias > 0.0f;
    volatile float acc = 0.0f; /* volatile: forbid any re-association / contraction by the compiler */
    for (uint64_t i = b; i < e; ++i) {
        const uint32_t r = W->row_idx[i];
        if (use_bias && r == W->rows - 1) continue; /* handled below */
        const uint32_t t = lower_bound_u32(qidx, qn, r);
        if (t < qn && qidx[t] == r) {
            volatile float prod = qval[t] * W->val[i];
            acc = acc + prod;
        }
    }
    if (use_bias) {
        for (uint64_t i = b; i < e; ++i) {
            if (W->row_idx[i] == W->rows - 1) {
                volatile float prod = bias * W->val[i];
                acc = acc + prod;
            }
        }
    }
    return acc;
}

/* Dense query: bias row first, then every stored row of the column in ascending order (zeros included). */
static float score_dense(const xlo_csc_t* W, uint32_t col, const float* x, float bias) {
    const uint64_t b = W->col_ptr[col], e = W->col_ptr[col + 1];
    const int use_bias = bias > 0.0f;
    volatile float acc = 0.0f;
    if (use_bias) {
        for (uint64_t i = b; i < e; ++i) {
            if (W->row_idx[i] == W->rows - 1) {
                volatile float prod = bias * W->val[i];
                acc = acc + prod;
            }
        }
    }
    for (uint64_t i = b; i < e; ++i) {
        const uint32_t r = W->row_idx[i];
        if (use_bias && r == W->rows - 1) continue;
        volatile float prod = x[r] * W->val[i];
        acc = acc + prod;
    }
    return acc;
}

typedef struct {
    float val;
    uint32_t pos;
    uint32_t label;
} xlo_cand_t;

static int cmp_desc_then_pos(const void* pa, const void* pb) {
    const xlo_cand_t* a = (const xlo_cand_t*)pa;
    const xlo_cand_t* b = (const xlo_cand_t*)pb;
    if (a->val == b->val) return (a->pos > b->pos) - (a->pos < b->pos);
    return (a->val > b->val) ? -1 : 1;
}

/* Sorted-by-row copy of a CSC column is assumed (scipy writes sorted indices; the reference would stable-sort). */

/* Layers [0, depth) starting from a given beam.  codes == NULL: prev_layer_pred = ones(Q x C[0].cols) and the first
 * layer does not combine (HierarchicalMLModel::predict, inference.hpp:2462-2463; c_xlinear_single_layer_predict with
 * csr_codes == NULL, libpecos.cpp:213-219).  codes != NULL: a CSR matrix Q x C[0].cols (row_ptr / col_idx / val) whose
 * entries, in stored order, are the beam entering layer 0, and layer 0 combines with their values
 * (libpecos.cpp:209-212: no_prev_pred = false). */
int xlo_predict_from(int depth, const xlo_csc_t* W, const xlo_csc_t* C, const float* bias, const int* pp_kind,
                     const int* pp_p, const uint32_t* only_topk, const xlo_query_t* X, const xlo_query_t* codes,
                     xlo_result_t* out) {
    const uint32_t Q = X->rows;
    /* beam per query: ids / vals, ragged */
    uint64_t* beam_ptr = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)Q + 1));
    uint32_t* beam_id;
    float* beam_val;
    if (codes) {
        const uint64_t n = codes->row_ptr[Q];
        beam_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n ? n : 1));
        beam_val = (float*)malloc(sizeof(float) * (size_t)(n ? n : 1));
        memcpy(beam_ptr, codes->row_ptr, sizeof(uint64_t) * ((size_t)Q + 1));
        memcpy(beam_id, codes->col_idx, sizeof(uint32_t) * (size_t)n);
        memcpy(beam_val, codes->val, sizeof(float) * (size_t)n);
    } else {
        const uint32_t nc = depth > 0 ? C[0].cols : 1; /* fill_ones(X.rows, C->cols) */
        beam_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(Q ? (size_t)Q * nc : 1));
        beam_val = (float*)malloc(sizeof(float) * (size_t)(Q ? (size_t)Q * nc : 1));
        for (uint32_t q = 0; q < Q; ++q) {
            beam_ptr[q] = (uint64_t)q * nc;
            for (uint32_t j = 0; j < nc; ++j) { beam_id[(size_t)q * nc + j] = j; beam_val[(size_t)q * nc + j] = 1.0f; }
        }
        beam_ptr[Q] = (uint64_t)Q * nc;
    }
    const int combine_first = codes != NULL;
    uint32_t out_cols = 1;

    for (int d = 0; d < depth; ++d) {
        const xlo_csc_t* Wd = &W[d];
        const xlo_csc_t* Cd = &C[d];
        const uint32_t k = only_topk[d];
        /* pass 1: sizes */
        uint64_t* new_ptr = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)Q + 1));
        new_ptr[0] = 0;
        uint64_t max_cand = 0;
        for (uint32_t q = 0; q < Q; ++q) {
            uint64_t n = 0;
            for (uint64_t j = beam_ptr[q]; j < beam_ptr[q + 1]; ++j) n += Cd->col_ptr[beam_id[j] + 1] - Cd->col_ptr[beam_id[j]];
            if (n > max_cand) max_cand = n;
            new_ptr[q + 1] = new_ptr[q] + (n < k ? n : k);
        }
        uint32_t* new_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(new_ptr[Q] ? new_ptr[Q] : 1));
        float* new_val = (float*)malloc(sizeof(float) * (size_t)(new_ptr[Q] ? new_ptr[Q] : 1));
        xlo_cand_t* cand = (xlo_cand_t*)malloc(sizeof(xlo_cand_t) * (size_t)(max_cand ? max_cand : 1));
        for (uint32_t q = 0; q < Q; ++q) {
            uint32_t n = 0;
            const uint32_t* qidx = NULL; const float* qval = NULL; uint32_t qn = 0;
            const float* xd = NULL;
            if (X->row_ptr) {
                qidx = X->col_idx + X->row_ptr[q]; qval = X->val + X->row_ptr[q];
                qn = (uint32_t)(X->row_ptr[q + 1] - X->row_ptr[q]);
            } else {
                xd = X->val + (size_t)q * X->cols;
            }
            for (uint64_t j = beam_ptr[q]; j < beam_ptr[q + 1]; ++j) {
                const uint32_t parent = beam_id[j];
                for (uint64_t c = Cd->col_ptr[parent]; c < Cd->col_ptr[parent + 1]; ++c) {
                    const uint32_t label = Cd->row_idx[c]; /* == c when the tree is contiguously ordered */
                    float raw = X->row_ptr ? score_sparse(Wd, label, qidx, qval, qn, bias[d]) : score_dense(Wd, label, xd, bias[d]);
                    float v = xlo_transform(raw, pp_kind[d], pp_p[d]);
                    if (d > 0 || combine_first) v = xlo_combine(v, beam_val[j], pp_kind[d]);
                    cand[n].val = v; cand[n].pos = n; cand[n].label = label;
                    ++n;
                }
            }
            qsort(cand, n, sizeof(xlo_cand_t), cmp_desc_then_pos);
            const uint32_t keep = (uint32_t)(new_ptr[q + 1] - new_ptr[q]);
            for (uint32_t i = 0; i < keep; ++i) { new_id[new_ptr[q] + i] = cand[i].label; new_val[new_ptr[q] + i] = cand[i].val; }
        }
        free(cand);
        free(beam_ptr); free(beam_id); free(beam_val);
        beam_ptr = new_ptr; beam_id = new_id; beam_val = new_val;
        out_cols = Cd->rows; /* W.cols when contiguous, perm.size() when rearranged: both == C.rows */
    }
    out->indptr = beam_ptr;
    out->indices = beam_id;
    out->data = beam_val;
    out->nnz = beam_ptr[Q];
    out->rows = Q;
    out->cols = out_cols;
    return 0;
}

int xlo_predict(int depth, const xlo_csc_t* W, const xlo_csc_t* C, const float* bias, const int* pp_kind,
                const int* pp_p, const uint32_t* only_topk, const xlo_query_t* X, xlo_result_t* out) {
    return xlo_predict_from(depth, W, C, bias, pp_kind, pp_p, only_topk, X, NULL, out);
}

/* ---------------------------------------------------------------------------------------------------------------
 * predict_on_selected_outputs (next scope row): score exactly the given (query, label) pairs through the
 * hierarchy, no top-k.  Restates HierarchicalMLModel::predict_on_selected_outputs (inference.hpp:2507-2571), per layer
 * MLModel::predict_on_selected_outputs_internal (:2129-2180) with prolongate_sparse_predictions (:1302-1358).
 *   - selected set of layer l-1 = parents (through C[l]) of the selected set of layer l, as a SORTED set (smat_x_smat with
 *     sorted output, :2531-2540);
 *   - entries of a row at layer l: for every entry of the previous layer's row, in order, its children in C's column
 *     order that belong to the layer's selected set;
 *   - score = CSC path (bias term + sparse dot product, :1019-1035) == score_sparse bit for bit (float add commutes);
 *     transform, then combine with the parent's value except at layer 0.
 * selected: CSR pattern rows x n_labels (values ignored, column indices unique per row).
 * --------------------------------------------------------------------------------------------------------------- */
static int cmp_u32(const void* a, const void* b) {
    const uint32_t x = *(const uint32_t*)a, y = *(const uint32_t*)b;
    return (x > y) - (x < y);
}

static int sorted_contains(const uint32_t* a, uint32_t n, uint32_t key) {
    const uint32_t t = lower_bound_u32(a, n, key);
    return t < n && a[t] == key;
}

int xlo_predict_selected(int depth, const xlo_csc_t* W, const xlo_csc_t* C, const float* bias, const int* pp_kind,
                         const int* pp_p, const xlo_query_t* X, const xlo_query_t* selected, xlo_result_t* out) {
    const uint32_t Q = X->rows;
    /* child -> parent maps */
    uint32_t** parent_of = (uint32_t**)calloc((size_t)depth, sizeof(uint32_t*));
    for (int d = 0; d < depth; ++d) {
        parent_of[d] = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(C[d].rows ? C[d].rows : 1));
        for (uint32_t r = 0; r < C[d].rows; ++r) parent_of[d][r] = 0xFFFFFFFFu;
        for (uint32_t p = 0; p < C[d].cols; ++p)
            for (uint64_t j = C[d].col_ptr[p]; j < C[d].col_ptr[p + 1]; ++j) parent_of[d][C[d].row_idx[j]] = p;
    }
    uint64_t* res_ptr = (uint64_t*)malloc(sizeof(uint64_t) * ((size_t)Q + 1));
    memcpy(res_ptr, selected->row_ptr, sizeof(uint64_t) * ((size_t)Q + 1));
    const uint64_t total = res_ptr[Q];
    uint32_t* res_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(total ? total : 1));
    float* res_val = (float*)malloc(sizeof(float) * (size_t)(total ? total : 1));
    for (uint32_t q = 0; q < Q; ++q) {
        const uint32_t n_leaf = (uint32_t)(selected->row_ptr[q + 1] - selected->row_ptr[q]);
        /* selected sets per layer (sorted, unique) */
        uint32_t** sel = (uint32_t**)calloc((size_t)depth, sizeof(uint32_t*));
        uint32_t* sel_n = (uint32_t*)calloc((size_t)depth, sizeof(uint32_t));
        sel[depth - 1] = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(n_leaf ? n_leaf : 1));
        memcpy(sel[depth - 1], selected->col_idx + selected->row_ptr[q], sizeof(uint32_t) * n_leaf);
        qsort(sel[depth - 1], n_leaf, sizeof(uint32_t), cmp_u32);
        sel_n[depth - 1] = n_leaf;
        for (int d = depth - 1; d > 0; --d) {
            sel[d - 1] = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(sel_n[d] ? sel_n[d] : 1));
            uint32_t m = 0;
            for (uint32_t i = 0; i < sel_n[d]; ++i) {
                const uint32_t par = parent_of[d][sel[d][i]];
                if (par != 0xFFFFFFFFu) sel[d - 1][m++] = par;
            }
            qsort(sel[d - 1], m, sizeof(uint32_t), cmp_u32);
            uint32_t u = 0;
            for (uint32_t i = 0; i < m; ++i) if (i == 0 || sel[d - 1][i] != sel[d - 1][i - 1]) sel[d - 1][u++] = sel[d - 1][i];
            sel_n[d - 1] = u;
        }
        const uint32_t* qidx = NULL; const float* qval = NULL; uint32_t qn = 0; const float* xd = NULL;
        if (X->row_ptr) {
            qidx = X->col_idx + X->row_ptr[q]; qval = X->val + X->row_ptr[q];
            qn = (uint32_t)(X->row_ptr[q + 1] - X->row_ptr[q]);
        } else {
            xd = X->val + (size_t)q * X->cols;
        }
        uint32_t prev_n = 1;
        uint32_t* prev_id = (uint32_t*)malloc(sizeof(uint32_t));
        float* prev_val = (float*)malloc(sizeof(float));
        prev_id[0] = 0; prev_val[0] = 1.0f;
        for (int d = 0; d < depth; ++d) {
            uint32_t* cur_id = (uint32_t*)malloc(sizeof(uint32_t) * (size_t)(sel_n[d] ? sel_n[d] : 1));
            float* cur_val = (float*)malloc(sizeof(float) * (size_t)(sel_n[d] ? sel_n[d] : 1));
            uint32_t k = 0;
            for (uint32_t i = 0; i < prev_n; ++i) {
                const uint32_t par = prev_id[i];
                for (uint64_t j = C[d].col_ptr[par]; j < C[d].col_ptr[par + 1]; ++j) {
                    const uint32_t label = C[d].row_idx[j];
                    if (!sorted_contains(sel[d], sel_n[d], label) || k >= sel_n[d]) continue;
                    const float raw = X->row_ptr ? score_sparse(&W[d], label, qidx, qval, qn, bias[d]) : score_dense(&W[d], label, xd, bias[d]);
                    float v = xlo_transform(raw, pp_kind[d], pp_p[d]);
                    if (d > 0) v = xlo_combine(v, prev_val[i], pp_kind[d]);
                    cur_id[k] = label; cur_val[k] = v; ++k;
                }
            }
            free(prev_id); free(prev_val);
            prev_id = cur_id; prev_val = cur_val; prev_n = k;
        }
        /* the reference copies the selected row's length; entries it could not reach stay unspecified (zero here) */
        for (uint32_t i = 0; i < n_leaf; ++i) {
            res_id[res_ptr[q] + i] = i < prev_n ? prev_id[i] : 0u;
            res_val[res_ptr[q] + i] = i < prev_n ? prev_val[i] : 0.0f;
        }
        free(prev_id); free(prev_val);
        for (int d = 0; d < depth; ++d) free(sel[d]);
        free(sel); free(sel_n);
    }
    for (int d = 0; d < depth; ++d) free(parent_of[d]);
    free(parent_of);
    out->indptr = res_ptr; out->indices = res_id; out->data = res_val;
    out->nnz = total; out->rows = Q; out->cols = selected->cols;
    return 0;
}

void xlo_free_result(xlo_result_t* r) {
    free(r->indptr); free(r->indices); free(r->data);
    r->indptr = NULL; r->indices = NULL; r->data = NULL;
}
