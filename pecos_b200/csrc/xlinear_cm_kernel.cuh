// Chunk-major scoring: the (query, beam slot) pairs of a layer are bucketed by weight chunk, a CTA stages ONE chunk in
// shared memory and every LANE walks ITS OWN pair.
//
// Included by xlinear_engine.cu (inside its anonymous namespace, after the small device helpers).
//
// Why (profiles/r02_*): the query-major kernels spend 600 - 3,900 warp-instructions per pair on match compaction, prefix
// sums and on putting colliding entries of a 32-entry group back into feature order, and every probe / extent / entry
// access is a scattered global load (one L1 line each).  Here
//   * pairs are bucketed by chunk on the device (count -> scan -> scatter; the reference's b_sort_by_chunk,
//     pecos/core/xmc/inference.hpp:985-993);
//   * a work item = one chunk + up to W x 32 of its pairs.  The CTA stages the chunk's FEATURE MAP -- one bit per
//     feature plus a 16-bit row prefix per 32 features: "is feature f a row of this chunk, and which one" is one
//     shared-memory word + a popcount, no probing loop, no divergence -- and its row pointers (u16) and entries
//     (u8 column + f32 weight, split arrays);
//   * a lane walks its pair's query features in ascending order (staged through shared memory in coalesced rounds of 16),
//     compacts the hits of a round in place, then adds the hit rows' entries to the lane's PRIVATE accumulators
//     acc[column][lane] -- literally the reference's marching loop (inference.hpp:788-811): ascending feature order,
//     separate multiply and add, bias row last.  No compaction across lanes, no conflict resolution: the order is right
//     by construction, and a column is only ever touched by the lane that owns the pair.
//
// Bit-identical to the query-major kernels (tests/test_chunk_major_gpu.py); eligibility is decided per layer on the
// host (cm_plan): sparse queries, feature map + chunk fit in shared memory, chunk rows / entries < 65536, width <= 256,
// enough pairs per chunk to amortise the staging.
#pragma once

constexpr int kCmFeat = 16;                 // query features staged per pair and round
constexpr int kCmMaxWarps = 16;
constexpr int kCmMinWarps = 4;
constexpr uint32_t kCmSmemBudget = 224u << 10;  // dynamic shared memory a CTA may take (227 KB is the sm_100a maximum)
constexpr uint32_t kCmMinReuse = 24;        // average pairs per chunk below which the per-chunk staging does not pay
constexpr uint32_t kCmEmpty = 0xFFFFFFFFu;

struct CmWork {
    uint32_t* slot_pos;     // [rows x beam_stride] first candidate position of every beam slot
    uint32_t* count;        // [n_chunks] pairs per chunk, reused as the scatter cursor
    uint32_t* bucket_ptr;   // [n_chunks + 1]
    uint32_t* item_ptr;     // [n_chunks + 1] work items (<= item_pairs pairs each) per chunk
    uint32_t* pair_q;       // [pairs] query of a pair, grouped by chunk
    uint32_t* pair_pos;     // [pairs] candidate position of the pair's first column inside the query's row
    uint32_t item_pairs;    // pairs per work item = 32 x warps of the score kernel
};

struct CmPlan {  // host-side launch plan of one layer
    bool eligible = false;
    uint32_t warps = 0;
    uint32_t fm_words = 0;   // feature-map cells staged
    uint32_t r_cap = 0;      // rows of the largest chunk
    uint32_t e_cap = 0;      // entries of the largest chunk
    uint32_t acc_cols = 0;   // accumulator rows per warp (= widest chunk)
    size_t smem = 0;
};

__host__ __device__ inline size_t cm_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }

__host__ __device__ inline size_t cm_chunk_bytes(uint32_t fm_words, uint32_t r_cap, uint32_t e_cap) {
    return cm_align16(static_cast<size_t>(fm_words) * 4)      // bits
           + cm_align16(static_cast<size_t>(fm_words) * 2)    // row prefix per cell (u16)
           + cm_align16(static_cast<size_t>(r_cap + 2) * 2)   // row pointers (u16, relative to the chunk's first entry)
           + cm_align16(static_cast<size_t>(e_cap + 1) * 4)   // entry weights
           + cm_align16(static_cast<size_t>(e_cap + 1));      // entry columns (u8)
}

__host__ __device__ inline size_t cm_warp_bytes(uint32_t acc_cols) {
    return static_cast<size_t>(32) * (kCmFeat + 1) * 8        // staged query features / compacted hits (idx + val), stride 17
           + static_cast<size_t>(acc_cols) * 32 * 4;          // accumulators [col][lane]
}

inline CmPlan cm_plan(uint32_t fm_words, uint32_t r_max, uint32_t e_max, uint32_t c_max, uint32_t n_chunks, uint64_t pairs) {
    CmPlan p;
    if (fm_words == 0 || n_chunks == 0 || c_max == 0 || c_max > 256u || r_max >= 65535u || e_max >= 65535u) return p;
    if (pairs < static_cast<uint64_t>(kCmMinReuse) * n_chunks) return p;
    const size_t chunk = cm_chunk_bytes(fm_words, r_max, e_max);
    const size_t per_warp = cm_warp_bytes(c_max);
    if (chunk + kCmMinWarps * per_warp + 64 > kCmSmemBudget) return p;
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(kCmMaxWarps, (kCmSmemBudget - chunk - 64) / per_warp));
    // no point in more lanes than the average bucket holds
    const uint64_t avg = pairs / n_chunks;
    while (warps > kCmMinWarps && static_cast<uint64_t>(warps - 1) * 32 >= avg) --warps;
    p.eligible = true;
    p.warps = warps;
    p.fm_words = fm_words;
    p.r_cap = r_max;
    p.e_cap = e_max;
    p.acc_cols = c_max;
    p.smem = chunk + warps * per_warp + 64;
    return p;
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
        const uint32_t it = (n + w.item_pairs - 1) / w.item_pairs;
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
__global__ void __launch_bounds__(kCmMaxWarps * 32)
xl_cm_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, float* __restrict__ cand,
                    const uint64_t cand_stride_q, unsigned long long* stats, const uint32_t fm_words, const uint32_t r_cap,
                    const uint32_t e_cap, const uint32_t acc_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* sp = smem_raw;
    uint32_t* bits_s = reinterpret_cast<uint32_t*>(sp);            sp += cm_align16(static_cast<size_t>(fm_words) * 4);
    unsigned short* pre_s = reinterpret_cast<unsigned short*>(sp); sp += cm_align16(static_cast<size_t>(fm_words) * 2);
    unsigned short* rp_s = reinterpret_cast<unsigned short*>(sp);  sp += cm_align16(static_cast<size_t>(r_cap + 2) * 2);
    float* ew_s = reinterpret_cast<float*>(sp);                    sp += cm_align16(static_cast<size_t>(e_cap + 1) * 4);
    unsigned char* ec_s = sp;                                      sp += cm_align16(static_cast<size_t>(e_cap + 1));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int kStride = kCmFeat + 1;
    unsigned char* mine = sp + static_cast<size_t>(warp) * cm_warp_bytes(acc_cols);
    uint32_t* st_idx = reinterpret_cast<uint32_t*>(mine);
    float* st_val = reinterpret_cast<float*>(st_idx + 32 * kStride);
    float* acc = st_val + 32 * kStride;                            // [acc_cols][32]

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
            s_first = w.bucket_ptr[c] + slice * w.item_pairs;
            s_last = min(s_first + w.item_pairs, w.bucket_ptr[c + 1]);
        }
        s_chunk = c;
    }
    __syncthreads();
    const uint32_t c = s_chunk;
    if (c == kCmEmpty) return;

    // ---- this lane's pair: issue its loads before the staging traffic
    const uint32_t pidx = s_first + static_cast<uint32_t>(warp) * 32u + lane;
    const bool have = pidx < s_last;
    uint32_t q = 0, pos = 0;
    uint64_t qb = 0;
    uint32_t qn = 0;
    if (have) {
        q = w.pair_q[pidx];
        pos = w.pair_pos[pidx];
        qb = X.row_ptr[q] - X.nnz_base;
        qn = static_cast<uint32_t>(X.row_ptr[q + 1] - X.nnz_base - qb);
    }

    // ---- stage the chunk: feature map (bits + 16-bit row prefix), row pointers, entries (split columns / weights)
    const ChunkHeader h = L.chunks[c];
    const uint32_t R = h.nnz_rows;
    const uint32_t R4 = (R + 3u) & ~3u;
    const uint32_t* rp_g = L.meta + h.meta_off + R4;               // row_ptr[R + 1], relative to the chunk's first entry
    const uint2* ent_g = L.entries + h.ent_off;
    const uint2* fm_g = L.featmap + static_cast<uint64_t>(c) * L.fm_words;
    const uint32_t nthreads = blockDim.x;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < fm_words; i += nthreads) {  // unrolled: four independent 8-byte loads in flight per thread
        const uint2 cell = __ldg(fm_g + i);
        bits_s[i] = cell.x;
        pre_s[i] = static_cast<unsigned short>(cell.y);
    }
    for (uint32_t i = threadIdx.x; i <= R; i += nthreads) rp_s[i] = static_cast<unsigned short>(__ldg(rp_g + i));
    const uint32_t E = R ? __ldg(rp_g + R) : 0u;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < E; i += nthreads) {
        const uint2 en = __ldg(ent_g + i);
        ec_s[i] = static_cast<unsigned char>(en.x);
        ew_s[i] = __uint_as_float(en.y);
    }
    const uint32_t n_cols = h.n_cols;
    for (uint32_t col = 0; col < n_cols; ++col) acc[col * 32 + lane] = 0.0f;
    __syncthreads();

    // ---- one pair per lane
    if (__ballot_sync(kFull, have) == 0u) return;
    uint32_t qn_max = qn;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qn_max = max(qn_max, __shfl_xor_sync(kFull, qn_max, d));

    unsigned long long st_match = 0, st_ent = 0;
    uint32_t prev_f = kCmEmpty;
    uint32_t* my_idx = st_idx + lane * kStride;
    float* my_val = st_val + lane * kStride;
    float* my_acc = acc + lane;
    for (uint32_t t0 = 0; t0 < qn_max; t0 += kCmFeat) {
        // coalesced staging: for every pair of the warp, its next kCmFeat features (idx, val) -> row i of the stage
        __syncwarp();
        constexpr int kPerIter = 32 / kCmFeat;  // pairs staged per warp-wide load
        const int sub = lane / kCmFeat, fl = lane % kCmFeat;
#pragma unroll 4
        for (int i0 = 0; i0 < 32; i0 += kPerIter) {
            const int i = i0 + sub;
            const uint64_t b_i = __shfl_sync(kFull, qb, i);
            const uint32_t n_i = __shfl_sync(kFull, qn, i);
            if (t0 + fl < n_i) {
                st_idx[i * kStride + fl] = __ldg(X.col_idx + b_i + t0 + fl);
                st_val[i * kStride + fl] = __ldg(X.val + b_i + t0 + fl);
            }
        }
        __syncwarp();
        const uint32_t n_here = (qn > t0) ? min(static_cast<uint32_t>(kCmFeat), qn - t0) : 0u;
        // phase 1: probe the feature map, compact the hits of this round IN PLACE (slot cnt <= k was already consumed)
        uint32_t cnt = 0;
        for (uint32_t k = 0; k < n_here; ++k) {
            const uint32_t f = my_idx[k];
            const bool dup = (f == prev_f);  // a repeated column index only counts once (the first occurrence)
            prev_f = f;
            if (dup || f >= L.w_rows) continue;
            const uint32_t word = bits_s[f >> 5];
            const uint32_t bit = f & 31u;
            if ((word >> bit) & 1u) {
                const float x = my_val[k];
                my_idx[cnt] = static_cast<uint32_t>(pre_s[f >> 5]) + __popc(word & ((1u << bit) - 1u));
                my_val[cnt] = x;
                ++cnt;
            }
        }
        // phase 2: the hit rows' entries, in feature order, into this lane's accumulators
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t row = my_idx[i];
            const float x = my_val[i];
            const uint32_t eb = rp_s[row], ee = rp_s[row + 1];
            for (uint32_t e = eb; e < ee; ++e) {
                float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
                *a = __fadd_rn(*a, __fmul_rn(x, ew_s[e]));
            }
            if (STATS) st_ent += ee - eb;
        }
        if (STATS) st_match += cnt;
    }
    if (have && (h.has_bias & 1u)) {  // bias row last (inference.hpp:806-811)
        const uint32_t eb = rp_s[R - 1u], ee = rp_s[R];
        for (uint32_t e = eb; e < ee; ++e) {
            float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
            *a = __fadd_rn(*a, __fmul_rn(L.bias, ew_s[e]));
        }
        if (STATS) { st_match += 1; st_ent += ee - eb; }
    }
    if (have) {
        float* dst = cand + static_cast<uint64_t>(q) * cand_stride_q + pos;
        for (uint32_t col = 0; col < n_cols; ++col) dst[col] = my_acc[col * 32];
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
