/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the reference's HNSW dense search.
 *
 * Nothing under pecos_b200/ may link, import or call this file.  Parity status: PINNED -- tests/test_oracle_cpu.py
 * checks it against the golden vectors recorded from the reference on its own fixtures (tests/golden/hnsw_toy) and,
 * when oracle/_ref exists, bit-for-bit against the reference library on random indices.
 *
 * Restated code (paths relative to the reference checkout):
 *   hno_search_one ......... HNSW::predict_single      pecos/core/ann/hnsw.hpp:927-971
 *   search level 0 ......... HNSW::search_level        pecos/core/ann/hnsw.hpp:849-924
 *   heap_push / heap_pop ... heap_t (std::push_heap / std::pop_heap of libstdc++; ties between equal distances are
 *                            resolved by those algorithms, so they are restated step by step)   hnsw.hpp:374-408
 *   hno_distance ........... FeatVecDense{IP,L2}Simd::distance   pecos/core/ann/feat_vectors.hpp:134-162
 *                            do_dot_product_simd / do_l2_distance_simd, per ISA clone
 *                                                       pecos/core/ann/distance_impl/x86.hpp:37-296
 *                            isa 0 = avx512f (16 partial sums, fold 16->4, then ((s0+s1)+s2)+s3; scalar tail FUSED
 *                                    multiply-add as compiled by GCC 13 for that clone -- checked by disassembly),
 *                            isa 1 = avx (8 partial sums), isa 2 = sse (4 partial sums), isa 3 = plain scalar loop.
 *   level-0 node record .... GraphL0 [deg][maxM0 ids][len][d floats]   hnsw.hpp:92-178
 *   SPARSE indices (csr) ... FeatVecSparse record [deg][maxM0 ids][len][len floats][len u32 indices] at byte offset
 *                            mem_start_of_node[node]                      feat_vectors.hpp:97-131, hnsw.hpp:122-176
 *   hno_sparse_dot ......... do_dot_product_sparse_block<4>               distance_impl/common.hpp:15-86   (the `avx` clone,
 *                            x86.hpp:305-428, intersects the same 4x4 blocks with SSE compares and adds the matched
 *                            products of a block in ascending position: the same sum for strictly ascending indices)
 *   hno_sparse_distance .... FeatVecSparse{IP,L2}Simd::distance           feat_vectors.hpp:186-210.  NOTE the reference's
 *                            sparse L2 calls do_l2_distance_simd(x, x, len) for the squared norms (feat_vectors.hpp:189-190),
 *                            i.e. the distance of x to itself = 0: its "l2" is -2 * <x, y>.  Restated literally.
 *   upper levels ........... GraphL1                                     hnsw.hpp:180-220
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    float dist;
    uint32_t node;
} hno_pair_t;

typedef struct {
    uint32_t num_node, maxM, maxM0, efC, max_level, init_node;
    uint32_t feat_dim, l0_max_degree, l0_node_mem_size; /* bytes */
    const uint8_t* l0_buffer;
    uint32_t l1_max_level, l1_max_degree, l1_node_mem_size, l1_level_mem_size; /* u32 units */
    const uint32_t* l1_buffer;
    int metric; /* 0 = ip, 1 = l2 */
    int isa;    /* see header */
    int sparse; /* 1: FeatVecSparse records located by mem_start (variable size) */
    const uint64_t* mem_start; /* [num_node + 1] byte offsets into l0_buffer (sparse only) */
} hno_index_t;

static float hno_dot_or_l2(const float* x, const float* y, size_t len, int metric, int isa) {
    const int lanes = isa == 0 ? 16 : (isa == 1 ? 8 : (isa == 2 ? 4 : 0));
    if (lanes == 0) {
        volatile float sum = 0.0f;
        for (size_t i = 0; i < len; ++i) {
            if (metric == 0) { volatile float p = x[i] * y[i]; sum = sum + p; }
            else { volatile float d = x[i] - y[i]; volatile float p = d * d; sum = sum + p; }
        }
        return sum;
    }
    volatile float acc[16];
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    const size_t len16 = len / 16, len4 = len / 4;
    size_t i = 0;
    for (; i < 16 * len16; ++i) {
        const int j = (int)(i % (size_t)lanes);
        if (metric == 0) { volatile float p = x[i] * y[i]; acc[j] = acc[j] + p; }
        else { volatile float d = x[i] - y[i]; volatile float p = d * d; acc[j] = acc[j] + p; }
    }
    volatile float s[4];
    if (lanes == 16) {
        for (int j = 0; j < 4; ++j) {
            volatile float a = acc[j] + acc[4 + j];
            volatile float b = acc[8 + j] + acc[12 + j];
            s[j] = a + b;
        }
    } else if (lanes == 8) {
        for (int j = 0; j < 4; ++j) s[j] = acc[j] + acc[4 + j];
    } else {
        for (int j = 0; j < 4; ++j) s[j] = acc[j];
    }
    for (; i < 4 * len4; ++i) {
        const int j = (int)(i % 4);
        if (metric == 0) { volatile float p = x[i] * y[i]; s[j] = s[j] + p; }
        else { volatile float d = x[i] - y[i]; volatile float p = d * d; s[j] = s[j] + p; }
    }
    volatile float sum = s[0] + s[1];
    sum = sum + s[2];
    sum = sum + s[3];
    for (; i < len; ++i) {
        if (isa == 0) {
            if (metric == 0) sum = fmaf(x[i], y[i], sum);
            else { volatile float d = x[i] - y[i]; sum = fmaf(d, d, sum); }
        } else {
            if (metric == 0) { volatile float p = x[i] * y[i]; sum = sum + p; }
            else { volatile float d = x[i] - y[i]; volatile float p = d * d; sum = sum + p; }
        }
    }
    return sum;
}

float hno_distance(const float* x, const float* y, uint32_t len, int metric, int isa) {
    const float v = hno_dot_or_l2(x, y, len, metric, isa);
    if (metric == 0) return (float)(1.0 - v); /* `1.0 - do_dot_product_simd(...)` narrowed to VAL_T */
    return v;
}

/* ---- libstdc++ heap algorithms, comparator passed as "a is worse-ordered than b" ------------------------------- */
/* max-heap (topk_queue): comp(a,b) = a.dist < b.dist ; min-heap (cand_queue): comp(a,b) = a.dist > b.dist */
static int comp_less(hno_pair_t a, hno_pair_t b) { return a.dist < b.dist; }
static int comp_greater(hno_pair_t a, hno_pair_t b) { return a.dist > b.dist; }
typedef int (*hno_comp_t)(hno_pair_t, hno_pair_t);

static void std_push_heap_(hno_pair_t* first, long hole, long top, hno_pair_t value, hno_comp_t comp) {
    long parent = (hole - 1) / 2;
    while (hole > top && comp(first[parent], value)) {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void std_adjust_heap_(hno_pair_t* first, long hole, long len, hno_pair_t value, hno_comp_t comp) {
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (comp(first[child], first[child - 1])) child--;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    std_push_heap_(first, hole, top, value, comp);
}

static void heap_push(hno_pair_t* h, long* n, hno_pair_t v, hno_comp_t comp) {
    h[*n] = v;
    *n += 1;
    std_push_heap_(h, *n - 1, 0, v, comp);
}

static void heap_pop(hno_pair_t* h, long* n, hno_comp_t comp) {
    if (*n > 1) {
        hno_pair_t value = h[*n - 1];
        h[*n - 1] = h[0];
        std_adjust_heap_(h, 0, *n - 1, value, comp);
    }
    *n -= 1;
}

static const uint8_t* l0_record(const hno_index_t* ix, uint32_t node) {
    return ix->l0_buffer + (ix->sparse ? (size_t)ix->mem_start[node] : (size_t)node * ix->l0_node_mem_size);
}
static const uint32_t* l0_neighborhood(const hno_index_t* ix, uint32_t node) { return (const uint32_t*)l0_record(ix, node); }
static const float* l0_vector(const hno_index_t* ix, uint32_t node) {
    return (const float*)(l0_record(ix, node) + (size_t)(1 + ix->l0_max_degree) * 4 + 4);
}

/* do_dot_product_sparse_block<4> (distance_impl/common.hpp:15-86) */
float hno_sparse_dot(size_t s_a, const float* x, const uint32_t* A, size_t s_b, const float* y, const uint32_t* B) {
    size_t i_a = 0, i_b = 0;
    volatile float ret = 0.0f;
    const size_t st_a = (s_a / 4) * 4, st_b = (s_b / 4) * 4;
    if (i_a < st_a && i_b < st_b) {
        for (;;) {
            for (int i = 0; i < 4; ++i) {
                for (int j = 0; j < 4; ++j) {
                    if (A[i_a + i] == B[i_b + j]) {
                        volatile float p = x[i_a + i] * y[i_b + j];
                        ret = ret + p;
                        break;
                    }
                }
            }
            const uint32_t a_last = A[i_a + 3], b_last = B[i_b + 3];
            if (a_last <= b_last) { i_a += 4; if (i_a == st_a) break; }
            if (a_last >= b_last) { i_b += 4; if (i_b == st_b) break; }
        }
    }
    if (i_a < s_a && i_b < s_b) {
        for (;;) {
            while (A[i_a] < B[i_b]) { if (++i_a == s_a) return ret; }
            while (B[i_b] < A[i_a]) { if (++i_b == s_b) return ret; }
            if (A[i_a] == B[i_b]) {
                volatile float p = x[i_a] * y[i_b];
                ret = ret + p;
                if (++i_a == s_a || ++i_b == s_b) return ret;
            }
        }
    }
    return ret;
}

/* FeatVecSparse{IP,L2}Simd::distance (feat_vectors.hpp:186-210) */
float hno_sparse_distance(size_t s_a, const float* x, const uint32_t* A, size_t s_b, const float* y, const uint32_t* B, int metric,
                          int isa) {
    const float dot = hno_sparse_dot(s_a, x, A, s_b, y, B);
    if (metric == 0) return (float)(1.0 - dot);
    volatile float x_sq = hno_dot_or_l2(x, x, s_a, 1, isa); /* do_l2_distance_simd(x, x, s_a): zero for finite values */
    volatile float y_sq = hno_dot_or_l2(y, y, s_b, 1, isa);
    volatile float sq = x_sq + y_sq;
    return (float)(sq - 2.0 * dot);
}

typedef struct {
    const float* dense;      /* dense query row, or */
    uint32_t nnz;            /* sparse query row */
    const float* val;
    const uint32_t* idx;
} hno_query_t;

static float query_distance(const hno_index_t* ix, const hno_query_t* q, uint32_t node) {
    if (!ix->sparse) return hno_distance(q->dense, l0_vector(ix, node), ix->feat_dim, ix->metric, ix->isa);
    const uint8_t* fv = l0_record(ix, node) + (size_t)(1 + ix->l0_max_degree) * 4;
    const uint32_t len = *(const uint32_t*)fv;
    const float* val = (const float*)(fv + 4);
    const uint32_t* idx = (const uint32_t*)(fv + 4 + (size_t)len * 4);
    return hno_sparse_distance(q->nnz, q->val, q->idx, len, val, idx, ix->metric, ix->isa);
}
static const uint32_t* l1_neighborhood(const hno_index_t* ix, uint32_t node, uint32_t level) {
    return ix->l1_buffer + (size_t)node * ix->l1_node_mem_size + (size_t)(level - 1) * ix->l1_level_mem_size;
}

/* Search the index for every query row; out arrays are Q x topk and must be zero-initialised by the caller
 * (libpecos.cpp:554-558 only writes the entries that exist).  counters (may be NULL): per query
 * {distance evaluations, level-0 expansions, upper-level hops (neighbourhood reads on levels >= 1)}. */
static int search_core(const hno_index_t* ix, const float* Q, const uint64_t* q_indptr, const uint32_t* q_idx, const float* q_val,
                       uint32_t nq, uint32_t efS, uint32_t topk, uint32_t* out_idx, float* out_val, uint64_t* counters) {
    const uint32_t d = ix->feat_dim;
    const uint32_t ef = efS > topk ? efS : topk;
    uint8_t* visited = (uint8_t*)calloc(ix->num_node ? ix->num_node : 1, 1);
    hno_pair_t* topq = (hno_pair_t*)malloc(sizeof(hno_pair_t) * ((size_t)ef + 2));
    hno_pair_t* candq = (hno_pair_t*)malloc(sizeof(hno_pair_t) * ((size_t)ix->num_node + 2));
    uint32_t* touched = (uint32_t*)malloc(sizeof(uint32_t) * ((size_t)ix->num_node + 1));
    if (!visited || !topq || !candq || !touched) return 1;

    for (uint32_t qi = 0; qi < nq; ++qi) {
        hno_query_t qrow = {0, 0, 0, 0};
        if (ix->sparse) {
            qrow.nnz = (uint32_t)(q_indptr[qi + 1] - q_indptr[qi]);
            qrow.val = q_val + q_indptr[qi];
            qrow.idx = q_idx + q_indptr[qi];
        } else {
            qrow.dense = Q + (size_t)qi * d;
        }
        const hno_query_t* q = &qrow;
        uint64_t n_dist = 0, n_expand = 0, n_hops = 0;
        uint32_t curr = ix->init_node;
        float curr_dist = query_distance(ix, q, curr);
        ++n_dist;
        for (uint32_t level = ix->max_level; level >= 1; --level) {
            int changed = 1;
            while (changed) {
                changed = 0;
                const uint32_t* nb = l1_neighborhood(ix, curr, level);
                const uint32_t deg = nb[0];
                ++n_hops;
                for (uint32_t j = 0; j < deg; ++j) {
                    const uint32_t next = nb[1 + j];
                    const float nd = query_distance(ix, q, next);
                    ++n_dist;
                    if (nd < curr_dist) { curr_dist = nd; curr = next; changed = 1; }
                }
            }
        }
        long ntop = 0, ncand = 0, ntouched = 0;
        float ub = query_distance(ix, q, curr);
        ++n_dist;
        hno_pair_t p0 = {ub, curr};
        heap_push(topq, &ntop, p0, comp_less);
        heap_push(candq, &ncand, p0, comp_greater);
        visited[curr] = 1; touched[ntouched++] = curr;
        while (ncand > 0) {
            const hno_pair_t c = candq[0];
            if (c.dist > ub) break;
            heap_pop(candq, &ncand, comp_greater);
            const uint32_t* nb = l0_neighborhood(ix, c.node);
            const uint32_t deg = nb[0];
            ++n_expand;
            for (uint32_t j = 0; j < deg; ++j) {
                const uint32_t next = nb[1 + j];
                if (visited[next]) continue;
                visited[next] = 1; touched[ntouched++] = next;
                const float nd = query_distance(ix, q, next);
                ++n_dist;
                if ((uint32_t)ntop < ef || nd < ub) {
                    hno_pair_t pn = {nd, next};
                    heap_push(candq, &ncand, pn, comp_greater);
                    heap_push(topq, &ntop, pn, comp_less);
                    if ((uint32_t)ntop > ef) heap_pop(topq, &ntop, comp_less);
                    if (ntop > 0) ub = topq[0].dist;
                }
            }
        }
        if (topk < efS) while ((uint32_t)ntop > topk) heap_pop(topq, &ntop, comp_less);
        /* std::sort_heap: repeated pop_heap leaves ascending order */
        long n = ntop;
        while (n > 1) {
            hno_pair_t value = topq[n - 1];
            topq[n - 1] = topq[0];
            std_adjust_heap_(topq, 0, n - 1, value, comp_less);
            --n;
        }
        for (long k = 0; k < ntop; ++k) {
            out_idx[(size_t)qi * topk + k] = topq[k].node;
            out_val[(size_t)qi * topk + k] = topq[k].dist;
        }
        for (long t = 0; t < ntouched; ++t) visited[touched[t]] = 0;
        if (counters) { counters[3 * (size_t)qi] = n_dist; counters[3 * (size_t)qi + 1] = n_expand; counters[3 * (size_t)qi + 2] = n_hops; }
    }
    free(visited); free(topq); free(candq); free(touched);
    return 0;
}

int hno_search(const hno_index_t* ix, const float* Q, uint32_t nq, uint32_t efS, uint32_t topk, uint32_t* out_idx,
               float* out_val, uint64_t* counters) {
    if (ix->sparse) return 2;
    return search_core(ix, Q, 0, 0, 0, nq, efS, topk, out_idx, out_val, counters);
}

/* sparse (csr) queries against a sparse index: rows are [q_indptr[i], q_indptr[i+1]) of q_idx / q_val, indices ascending */
int hno_search_csr(const hno_index_t* ix, const uint64_t* q_indptr, const uint32_t* q_idx, const float* q_val, uint32_t nq,
                   uint32_t efS, uint32_t topk, uint32_t* out_idx, float* out_val, uint64_t* counters) {
    if (!ix->sparse) return 2;
    return search_core(ix, 0, q_indptr, q_idx, q_val, nq, efS, topk, out_idx, out_val, counters);
}
