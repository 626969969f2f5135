// Chunk-major scoring for layers of MANY SMALL, HEAVILY REUSED chunks (e.g. the 64 / 512 / 4,096 eight-column chunks of the
// 3M-label tree's middle layers, each visited by 500 .. 30,000 (query, beam slot) pairs of a 100k-query batch).
//
// Included by xlinear_engine.cu (inside its anonymous namespace, after the small device helpers).
//
// STATUS: EXPERIMENTAL, OFF BY DEFAULT (kernel mode 5).  Written at the end of round 1 after the round's GPU budget was
// spent: it compiles for sm_100a but has not run on a GPU yet; tests/test_chunk_major_gpu.py is opt-in
// (PB200_UNVALIDATED=1).  DESIGN.md section 8 item 4 explains why this is the next step.
//
// Why: the query-warp kernel spends one L1 line (~2 cycles of the SM's L1 pipeline) on every (feature, chunk) probe --
// 2,560 per query and layer -- and ~600 warp-instructions per pair on compaction, prefix sums and conflict rounds.  Here
// the pairs of a layer are bucketed by chunk (the reference's b_sort_by_chunk, pecos/core/xmc/inference.hpp:985-993); a
// CTA takes one chunk and up to kCmPairs of its pairs, stages the chunk -- entries, row extents and an open-addressing
// hash of its row ids -- in shared memory, and then every LANE walks ITS OWN pair: the query's features in ascending
// order, one hash probe each, the matched row's entries added to the lane's private accumulators, bias row last.  That
// is literally the reference's marching loop (inference.hpp:788-811): same arithmetic order by construction, no
// compaction, no prefix sums, no conflict resolution, and every access except the query read and the result write hits
// shared memory.  Query features are staged through shared memory in coalesced 32-feature rounds (row stride 33 words, so
// the lane-per-row reads are bank-conflict free).
//
// Pipeline per layer: xl_cm_count_kernel (slot positions + pairs per chunk) -> xl_cm_scan_kernel (bucket and work-item
// offsets) -> xl_cm_scatter_kernel (pair lists) -> xl_cm_scores_kernel.  Eligibility (host): sparse queries, chunk width
// <= kCmCols, staged chunk <= kCmChunkBytes, at least kCmMinReuse pairs per chunk on average.
#pragma once

constexpr int kCmWarps = 8;                 // warps per CTA
constexpr int kCmPairs = kCmWarps * 32;     // pairs per work item
constexpr int kCmFeat = 16;                 // query features staged per pair and round
constexpr int kCmCols = 16;                 // widest chunk served (accumulators: kCmCols x 32 floats per warp)
constexpr uint32_t kCmChunkBytes = 64u << 10;  // staged chunk budget (hash + extents + entries)
constexpr uint32_t kCmMinReuse = 32;        // average pairs per chunk below which the per-chunk staging does not pay
constexpr uint32_t kCmEmpty = 0xFFFFFFFFu;

struct CmWork {
    uint32_t* slot_pos;     // [rows x beam_stride] first candidate position of every beam slot
    uint32_t* count;        // [n_chunks] pairs per chunk, reused as the scatter cursor
    uint32_t* bucket_ptr;   // [n_chunks + 1]
    uint32_t* item_ptr;     // [n_chunks + 1] work items (<= kCmPairs pairs each) per chunk
    uint32_t* pair_q;       // [pairs] query of a pair, grouped by chunk
    uint32_t* pair_pos;     // [pairs] candidate position of the pair's first column inside the query's row
};

__host__ __device__ inline uint32_t cm_hash_slots(uint32_t r_max) {  // power of two >= 2 * rows (load factor <= 0.5)
    uint32_t h = 16;
    while (h < 2u * r_max) h <<= 1;
    return h;
}

__host__ __device__ inline size_t cm_smem_bytes(uint32_t r_max, uint32_t e_max) {
    return static_cast<size_t>(cm_hash_slots(r_max)) * 8           // hash keys + values
           + static_cast<size_t>(r_max + 1) * 8                     // row extents
           + static_cast<size_t>(e_max + 1) * 8                     // entries
           + static_cast<size_t>(kCmWarps) * (32 * (kCmFeat + 1) * 8    // staged query features (idx + val), stride 33
                                              + kCmCols * 32 * 4)      // accumulators [col][lane]
           + 16;
}

// one warp per query: candidate position of every beam slot (prefix of the chunk widths) and pairs per chunk
__global__ void __launch_bounds__(128)
xl_cm_count_kernel(const LayerDev L, const QueryDev X, const uint32_t* __restrict__ beam_id,
                   const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, const uint32_t rows, CmWork w,
                   unsigned long long* stats) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= rows) return;
    const uint32_t cnt = beam_cnt[q];
    if (stats && lane == 0 && cnt > 0) atomicAdd(&stats[5], static_cast<unsigned long long>(X.row_ptr[q + 1] - X.row_ptr[q]));
    uint32_t run = 0;
    for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
        const uint32_t j = j0 + lane;
        uint32_t width = 0, p = 0;
        bool scored = false;
        if (j < cnt) {
            p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
            const uint4 h = *reinterpret_cast<const uint4*>(&L.chunks[p]);  // {col_begin, n_cols, nnz_rows, has_bias}
            width = h.y;
            scored = !(h.w & kChunkAbsent) && h.y > 0;
        }
        const uint32_t incl = warp_incl_scan(width, lane);
        if (j < cnt) {
            w.slot_pos[static_cast<uint64_t>(q) * beam_stride + j] = run + incl - width;
            if (scored) atomicAdd(&w.count[p], 1u);
        }
        run += __shfl_sync(kFull, incl, 31);
    }
}

// single CTA: exclusive scans of the pair counts (bucket offsets) and of the work items per chunk; count[] becomes the
// scatter cursor
__global__ void __launch_bounds__(1024)
xl_cm_scan_kernel(const uint32_t n_chunks, CmWork w) {
    __shared__ uint32_t s_pairs[32], s_items[32];
    __shared__ uint32_t carry_pairs, carry_items;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { carry_pairs = 0; carry_items = 0; }
    __syncthreads();
    for (uint32_t c0 = 0; c0 < n_chunks; c0 += 1024) {
        const uint32_t c = c0 + threadIdx.x;
        const uint32_t n = (c < n_chunks) ? w.count[c] : 0u;
        const uint32_t it = (n + kCmPairs - 1) / kCmPairs;
        const uint32_t in_p = warp_incl_scan(n, lane), in_i = warp_incl_scan(it, lane);
        if (lane == 31) { s_pairs[warp] = in_p; s_items[warp] = in_i; }
        __syncthreads();
        if (warp == 0) {
            const uint32_t a = s_pairs[lane], b = s_items[lane];
            const uint32_t ia = warp_incl_scan(a, lane), ib = warp_incl_scan(b, lane);
            s_pairs[lane] = ia - a;
            s_items[lane] = ib - b;
        }
        __syncthreads();
        const uint32_t ex_p = carry_pairs + s_pairs[warp] + in_p - n;
        const uint32_t ex_i = carry_items + s_items[warp] + in_i - it;
        if (c < n_chunks) {
            w.bucket_ptr[c] = ex_p;
            w.item_ptr[c] = ex_i;
            w.count[c] = ex_p;  // cursor
        }
        __syncthreads();
        if (threadIdx.x == 1023) { carry_pairs = ex_p + n; carry_items = ex_i + it; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { w.bucket_ptr[n_chunks] = carry_pairs; w.item_ptr[n_chunks] = carry_items; }
}

// one warp per query: append (query, position) to the pair list of every scored slot's chunk (order inside a bucket is
// irrelevant: a pair's result location is fixed by its query and position)
__global__ void __launch_bounds__(128)
xl_cm_scatter_kernel(const LayerDev L, const uint32_t* __restrict__ beam_id, const uint32_t* __restrict__ beam_cnt,
                     const uint32_t beam_stride, const uint32_t rows, CmWork w) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= rows) return;
    const uint32_t cnt = beam_cnt[q];
    for (uint32_t j = lane; j < cnt; j += 32) {
        const uint32_t p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
        const uint4 h = *reinterpret_cast<const uint4*>(&L.chunks[p]);
        if ((h.w & kChunkAbsent) || h.y == 0) continue;
        const uint32_t at = atomicAdd(&w.count[p], 1u);
        w.pair_q[at] = q;
        w.pair_pos[at] = w.slot_pos[static_cast<uint64_t>(q) * beam_stride + j];
    }
}

template <bool STATS>
__global__ void __launch_bounds__(kCmWarps * 32)
xl_cm_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, float* __restrict__ cand,
                    const uint64_t cand_stride_q, unsigned long long* stats, const uint32_t r_cap, const uint32_t e_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t H = cm_hash_slots(r_cap);
    uint32_t* hk = reinterpret_cast<uint32_t*>(smem_raw);            // [H] feature id or kCmEmpty
    uint32_t* hv = hk + H;                                           // [H] chunk row
    uint2* ext_s = reinterpret_cast<uint2*>(hv + H);                 // [r_cap + 1] {first entry, end} per chunk row
    uint2* ent_s = ext_s + (r_cap + 1);                              // [e_cap + 1] {col offset, weight}
    unsigned char* per_warp = reinterpret_cast<unsigned char*>(ent_s + (e_cap + 1));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int kStride = kCmFeat + 1;
    uint32_t* st_idx = reinterpret_cast<uint32_t*>(per_warp + static_cast<size_t>(warp) * (32 * kStride * 8 + kCmCols * 32 * 4));
    float* st_val = reinterpret_cast<float*>(st_idx + 32 * kStride);
    float* acc = st_val + 32 * kStride;                              // [kCmCols][32]

    // ---- which (chunk, slice) is this CTA's work item
    __shared__ uint32_t s_chunk, s_first, s_last;
    if (threadIdx.x == 0) {
        const uint32_t n_items = w.item_ptr[L.n_chunks];
        uint32_t c = kCmEmpty;
        if (blockIdx.x < n_items) {
            uint32_t lo = 0, hi = L.n_chunks;  // largest c with item_ptr[c] <= blockIdx.x (empty chunks share offsets)
            while (hi - lo > 1) {
                const uint32_t mid = (lo + hi) >> 1;
                if (w.item_ptr[mid] <= blockIdx.x) lo = mid; else hi = mid;
            }
            c = lo;
            const uint32_t slice = blockIdx.x - w.item_ptr[c];
            s_first = w.bucket_ptr[c] + slice * kCmPairs;
            s_last = min(s_first + kCmPairs, w.bucket_ptr[c + 1]);
        }
        s_chunk = c;
    }
    __syncthreads();
    const uint32_t c = s_chunk;
    if (c == kCmEmpty) return;

    // ---- stage the chunk: entries, row extents, hash of the row ids
    const ChunkHeader h = L.chunks[c];
    const uint32_t R = h.nnz_rows;
    const uint32_t R4 = (R + 3u) & ~3u;
    const uint32_t* ridx = L.meta + h.meta_off;
    const uint2* ext_g = reinterpret_cast<const uint2*>(L.rowext + h.meta_off);
    const uint2* ent_g = L.entries + h.ent_off;
    (void)R4;
    for (uint32_t i = threadIdx.x; i < H; i += blockDim.x) hk[i] = kCmEmpty;
    for (uint32_t i = threadIdx.x; i < R; i += blockDim.x) ext_s[i] = ext_g[i];
    __syncthreads();
    const uint32_t E = R ? ext_s[R - 1].y : 0u;
    for (uint32_t i = threadIdx.x; i < E; i += blockDim.x) ent_s[i] = ent_g[i];
    const uint32_t shift = 32u - static_cast<uint32_t>(__ffs(static_cast<int>(H)) - 1);  // H = 2^k: top k bits of the product
    for (uint32_t r = threadIdx.x; r < R; r += blockDim.x) {
        const uint32_t f = ridx[r];
        uint32_t slot = (f * 2654435761u) >> shift;
        while (atomicCAS(&hk[slot], kCmEmpty, f) != kCmEmpty) slot = (slot + 1u) & (H - 1u);  // row ids are distinct
        hv[slot] = r;
    }
    __syncthreads();

    // ---- one pair per lane
    const uint32_t pidx = s_first + static_cast<uint32_t>(warp) * 32u + lane;
    const bool have = pidx < s_last;
    if (__ballot_sync(kFull, have) == 0u) return;
    uint32_t q = 0, pos = 0;
    uint64_t qb = 0;
    uint32_t qn = 0;
    if (have) {
        q = w.pair_q[pidx];
        pos = w.pair_pos[pidx];
        qb = X.row_ptr[q] - X.nnz_base;
        qn = static_cast<uint32_t>(X.row_ptr[q + 1] - X.nnz_base - qb);
    }
    const uint32_t n_cols = h.n_cols;
    for (uint32_t col = 0; col < n_cols; ++col) acc[col * 32 + lane] = 0.0f;
    uint32_t qn_max = qn;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qn_max = max(qn_max, __shfl_xor_sync(kFull, qn_max, d));

    unsigned long long st_match = 0, st_ent = 0;
    uint32_t prev_f = kCmEmpty;
    for (uint32_t t0 = 0; t0 < qn_max; t0 += kCmFeat) {
        // coalesced staging: for every pair of the warp, its next kCmFeat features (idx, val) -> row i of the stage
        __syncwarp();
        constexpr int kPerIter = 32 / kCmFeat;  // pairs staged per warp-wide load
        const int sub = lane / kCmFeat, fl = lane % kCmFeat;
        for (int i0 = 0; i0 < 32; i0 += kPerIter) {
            const int i = i0 + sub;
            const uint64_t b_i = __shfl_sync(kFull, qb, i);
            const uint32_t n_i = __shfl_sync(kFull, qn, i);
            if (t0 + fl < n_i) {
                st_idx[i * kStride + fl] = X.col_idx[b_i + t0 + fl];
                st_val[i * kStride + fl] = X.val[b_i + t0 + fl];
            }
        }
        __syncwarp();
        const uint32_t n_here = (qn > t0) ? min(static_cast<uint32_t>(kCmFeat), qn - t0) : 0u;
        for (uint32_t k = 0; k < n_here; ++k) {
            const uint32_t f = st_idx[lane * kStride + k];
            const bool dup = (f == prev_f);  // a repeated column index only counts once (the first occurrence)
            prev_f = f;
            if (dup || f >= L.w_rows) continue;
            uint32_t slot = (f * 2654435761u) >> shift;
            uint32_t row = kCmEmpty;
            for (;;) {
                const uint32_t key = hk[slot];
                if (key == f) { row = hv[slot]; break; }
                if (key == kCmEmpty) break;
                slot = (slot + 1u) & (H - 1u);
            }
            if (row == kCmEmpty) continue;
            const float x = st_val[lane * kStride + k];
            const uint2 lh = ext_s[row];
            for (uint32_t e = lh.x; e < lh.y; ++e) {
                const uint2 en = ent_s[e];
                float* a = acc + en.x * 32 + lane;
                *a = __fadd_rn(*a, __fmul_rn(x, __uint_as_float(en.y)));
            }
            if (STATS) { st_match += 1; st_ent += lh.y - lh.x; }
        }
    }
    if (have && (h.has_bias & 1u)) {  // bias row last (inference.hpp:806-811)
        const uint2 lh = ext_s[R - 1u];
        for (uint32_t e = lh.x; e < lh.y; ++e) {
            const uint2 en = ent_s[e];
            float* a = acc + en.x * 32 + lane;
            *a = __fadd_rn(*a, __fmul_rn(L.bias, __uint_as_float(en.y)));
        }
        if (STATS) { st_match += 1; st_ent += lh.y - lh.x; }
    }
    if (have) {
        float* dst = cand + static_cast<uint64_t>(q) * cand_stride_q + pos;
        for (uint32_t col = 0; col < n_cols; ++col) dst[col] = acc[col * 32 + lane];
    }
    if (STATS) {
        unsigned long long pairs = have ? 1ull : 0ull, rows_sum = have ? R : 0ull, cols_sum = have ? n_cols : 0ull;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            pairs += __shfl_xor_sync(kFull, pairs, d);
            rows_sum += __shfl_xor_sync(kFull, rows_sum, d);
            cols_sum += __shfl_xor_sync(kFull, cols_sum, d);
            st_match += __shfl_xor_sync(kFull, st_match, d);
            st_ent += __shfl_xor_sync(kFull, st_ent, d);
        }
        if (lane == 0) {
            atomicAdd(&stats[0], pairs);
            atomicAdd(&stats[1], rows_sum);
            atomicAdd(&stats[2], st_match);
            atomicAdd(&stats[3], st_ent);
            atomicAdd(&stats[4], cols_sum);
        }
    }
}
