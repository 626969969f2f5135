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

constexpr int kCmFeat = 8;                  // query features staged per pair and round (two rounds in flight per warp)
constexpr int kCmMaxWarps = 16;
constexpr int kCmMinWarps = 4;
constexpr uint32_t kCmSmemBudget = 224u << 10;  // dynamic shared memory a CTA may take (227 KB is the sm_100a maximum)
constexpr uint32_t kCmMinReuse = 24;        // average pairs per chunk below which the per-chunk staging does not pay
constexpr uint32_t kCmMinItems = 148;       // fewer work items than SMs: the query-major kernels fill the GPU better
constexpr uint32_t kCmDirectRows = 16384;   // feature spaces up to this size get a direct feature -> entry-range table
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
    bool direct = false;     // feature -> {first entry, end} table (small feature spaces) instead of bits + prefix + row pointers
    uint32_t warps = 0;
    uint32_t fm_words = 0;   // feature-map cells staged (bits variant) / w_rows (direct variant)
    uint32_t r_cap = 0;      // rows of the largest chunk
    uint32_t e_cap = 0;      // entries of the largest chunk
    uint32_t acc_cols = 0;   // accumulator rows per warp (= widest chunk)
    size_t smem = 0;
};

__host__ __device__ inline size_t cm_align16(size_t x) { return (x + 15) & ~static_cast<size_t>(15); }

__host__ __device__ inline size_t cm_chunk_bytes(bool direct, uint32_t fm_words, uint32_t r_cap, uint32_t e_cap) {
    const size_t lookup = direct ? cm_align16(static_cast<size_t>(fm_words) * 4)            // {u16 begin, u16 end} per feature
                                 : cm_align16(static_cast<size_t>(fm_words) * 4)            // bits
                                       + cm_align16(static_cast<size_t>(fm_words) * 2)      // row prefix per cell (u16)
                                       + cm_align16(static_cast<size_t>(r_cap + 2) * 2);    // row pointers (u16)
    return lookup + cm_align16(static_cast<size_t>(e_cap + 1) * 4)   // entry weights
           + cm_align16(static_cast<size_t>(e_cap + 1));             // entry columns (u8)
}

__host__ __device__ inline size_t cm_warp_bytes(uint32_t acc_cols) {
    return static_cast<size_t>(2) * 32 * (kCmFeat + 1) * 8    // two staging buffers: query features / compacted hits, stride 9
           + static_cast<size_t>(acc_cols) * 32 * 4;          // accumulators [col][lane]
}

// force: take the kernel wherever it FITS, ignoring the reuse / occupancy heuristics (kernel mode 5, tests)
inline CmPlan cm_plan(uint32_t fm_words, uint32_t w_rows, uint32_t r_max, uint32_t e_max, uint32_t c_max, uint32_t n_chunks,
                      uint64_t pairs, bool force) {
    CmPlan p;
    if (fm_words == 0 || n_chunks == 0 || c_max == 0 || c_max > 256u || r_max >= 65535u || e_max >= 65535u || pairs == 0) return p;
    if (!force && pairs < static_cast<uint64_t>(kCmMinReuse) * n_chunks) return p;
    const bool direct = w_rows <= kCmDirectRows;
    const uint32_t words = direct ? w_rows : fm_words;
    const size_t chunk = cm_chunk_bytes(direct, words, r_max, e_max);
    const size_t per_warp = cm_warp_bytes(c_max);
    if (chunk + kCmMinWarps * per_warp + 64 > kCmSmemBudget) return p;
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(kCmMaxWarps, (kCmSmemBudget - chunk - 64) / per_warp));
    // no point in more lanes than the average bucket holds; and keep at least two waves of work items when possible
    const uint64_t avg = pairs / n_chunks;
    while (warps > kCmMinWarps && (static_cast<uint64_t>(warps - 1) * 32 >= avg || pairs / (32ull * warps) < 2ull * kCmMinItems)) --warps;
    if (!force && pairs / (32ull * warps) < kCmMinItems) return p;
    p.eligible = true;
    p.direct = direct;
    p.warps = warps;
    p.fm_words = words;
    p.r_cap = r_max;
    p.e_cap = e_max;
    p.acc_cols = c_max;
    p.smem = chunk + warps * per_warp + 64;
    return p;
}

__device__ __forceinline__ void cm_cp_async4(void* smem_dst, const void* gmem_src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cm_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

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

template <bool STATS, bool DIRECT>
__global__ void __launch_bounds__(kCmMaxWarps * 32)
xl_cm_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, float* __restrict__ cand,
                    const uint64_t cand_stride_q, unsigned long long* stats, const uint32_t fm_words, const uint32_t r_cap,
                    const uint32_t e_cap, const uint32_t acc_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned char* sp = smem_raw;
    // lookup structure: DIRECT: tab_s[f] = {first entry, end} (u16 | u16 << 16) of feature f's row, 0 = no row;
    //                   else:   bits_s / pre_s = the chunk's feature map, rp_s = row pointers
    uint32_t* bits_s = reinterpret_cast<uint32_t*>(sp);            sp += cm_align16(static_cast<size_t>(fm_words) * 4);
    unsigned short* pre_s = reinterpret_cast<unsigned short*>(sp);
    unsigned short* rp_s = nullptr;
    if (!DIRECT) {
        sp += cm_align16(static_cast<size_t>(fm_words) * 2);
        rp_s = reinterpret_cast<unsigned short*>(sp);
        sp += cm_align16(static_cast<size_t>(r_cap + 2) * 2);
    }
    float* ew_s = reinterpret_cast<float*>(sp);                    sp += cm_align16(static_cast<size_t>(e_cap + 1) * 4);
    unsigned char* ec_s = sp;                                      sp += cm_align16(static_cast<size_t>(e_cap + 1));
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    constexpr int kStride = kCmFeat + 1;
    constexpr int kBuf = 32 * kStride;                             // words per staging array
    unsigned char* mine = sp + static_cast<size_t>(warp) * cm_warp_bytes(acc_cols);
    uint32_t* st_idx = reinterpret_cast<uint32_t*>(mine);          // [2][32][kStride]
    float* st_val = reinterpret_cast<float*>(st_idx + 2 * kBuf);   // [2][32][kStride]
    float* acc = st_val + 2 * kBuf;                                // [acc_cols][32]

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

    // ---- this lane's pair
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
    uint32_t qn_max = qn;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) qn_max = max(qn_max, __shfl_xor_sync(kFull, qn_max, d));

    // query features travel global -> shared memory by cp.async, two rounds in flight per warp: round r + 1 is copied while
    // round r is processed.  Row i of a staging buffer = the next kCmFeat features of the warp's pair i (stride 9 words:
    // the lane-per-row reads are bank-conflict free).
    constexpr int kPerIter = 32 / kCmFeat;  // pairs covered by one warp-wide copy instruction
    const int sub = lane / kCmFeat, fl = lane % kCmFeat;
    auto stage_round = [&](uint32_t t0, int buf) {
        uint32_t* di = st_idx + buf * kBuf;
        float* dv = st_val + buf * kBuf;
#pragma unroll
        for (int i0 = 0; i0 < 32; i0 += kPerIter) {
            const int i = i0 + sub;
            const uint64_t b_i = __shfl_sync(kFull, qb, i);
            const uint32_t n_i = __shfl_sync(kFull, qn, i);
            if (t0 + fl < n_i) {
                cm_cp_async4(di + i * kStride + fl, X.col_idx + b_i + t0 + fl);
                cm_cp_async4(dv + i * kStride + fl, X.val + b_i + t0 + fl);
            }
        }
        cm_cp_async_commit();
    };
    if (qn_max > 0) stage_round(0, 0);  // in flight during the chunk staging below

    // ---- stage the chunk
    const ChunkHeader h = L.chunks[c];
    const uint32_t R = h.nnz_rows;
    const uint32_t R4 = (R + 3u) & ~3u;
    const uint32_t* ridx_g = L.meta + h.meta_off;                  // sorted feature ids of the chunk's rows
    const uint32_t* rp_g = ridx_g + R4;                            // row_ptr[R + 1], relative to the chunk's first entry
    const uint2* ent_g = L.entries + h.ent_off;
    const uint32_t nthreads = blockDim.x;
    if (DIRECT) {
        for (uint32_t i = threadIdx.x; i < fm_words; i += nthreads) bits_s[i] = 0u;
        __syncthreads();
#pragma unroll 2
        for (uint32_t r = threadIdx.x; r < R; r += nthreads) {
            const uint32_t f = __ldg(ridx_g + r);
            const uint32_t eb = __ldg(rp_g + r), ee = __ldg(rp_g + r + 1);
            if (f < fm_words) bits_s[f] = eb | (ee << 16);           // empty rows (eb == ee) read as "no row": nothing to add
        }
    } else {
        const uint2* fm_g = L.featmap + static_cast<uint64_t>(c) * L.fm_words;
#pragma unroll 4
        for (uint32_t i = threadIdx.x; i < fm_words; i += nthreads) {  // unrolled: four independent 8-byte loads per thread
            const uint2 cell = __ldg(fm_g + i);
            bits_s[i] = cell.x;
            pre_s[i] = static_cast<unsigned short>(cell.y);
        }
        for (uint32_t i = threadIdx.x; i <= R; i += nthreads) rp_s[i] = static_cast<unsigned short>(__ldg(rp_g + i));
    }
    const uint32_t E = R ? __ldg(rp_g + R) : 0u;
#pragma unroll 4
    for (uint32_t i = threadIdx.x; i < E; i += nthreads) {
        const uint2 en = __ldg(ent_g + i);
        ec_s[i] = static_cast<unsigned char>(en.x);
        ew_s[i] = __uint_as_float(en.y);
    }
    const uint32_t n_cols = h.n_cols;
    float* my_acc = acc + lane;
    for (uint32_t col = 0; col < n_cols; ++col) my_acc[col * 32] = 0.0f;
    uint32_t bias_range = 0;  // {first entry, end} of the bias row
    if (h.has_bias & 1u) bias_range = __ldg(rp_g + R - 1u) | (E << 16);
    __syncthreads();

    // ---- one pair per lane
    if (__ballot_sync(kFull, have) == 0u) { cm_cp_async_wait<0>(); return; }
    unsigned long long st_match = 0, st_ent = 0;
    uint32_t prev_f = kCmEmpty;
    // apply the entries [range & 0xFFFF, range >> 16) of one row to this lane's accumulators
    auto apply_row = [&](uint32_t range, float x) {
        const uint32_t ee = range >> 16;
        for (uint32_t e = range & 0xFFFFu; e < ee; ++e) {
            float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
            *a = __fadd_rn(*a, __fmul_rn(x, ew_s[e]));
        }
    };
    int buf = 0;
    for (uint32_t t0 = 0; t0 < qn_max; t0 += kCmFeat, buf ^= 1) {
        if (t0 + kCmFeat < qn_max) { stage_round(t0 + kCmFeat, buf ^ 1); cm_cp_async_wait<1>(); }
        else cm_cp_async_wait<0>();
        __syncwarp();
        uint32_t* my_idx = st_idx + buf * kBuf + lane * kStride;
        float* my_val = st_val + buf * kBuf + lane * kStride;
        const uint32_t n_here = (qn > t0) ? min(static_cast<uint32_t>(kCmFeat), qn - t0) : 0u;
        // phase 1: look the features up, compact the hits of this round IN PLACE as {entry range, x} (slot cnt <= k was
        // already consumed)
        uint32_t cnt = 0;
#pragma unroll
        for (uint32_t k = 0; k < static_cast<uint32_t>(kCmFeat); ++k) {
            if (k < n_here) {
                const uint32_t f = my_idx[k];
                const bool dup = (f == prev_f);  // a repeated column index only counts once (the first occurrence)
                prev_f = f;
                uint32_t range = 0;
                if (!dup && f < L.w_rows) {
                    if (DIRECT) {
                        range = bits_s[f];
                    } else {
                        const uint32_t word = bits_s[f >> 5];
                        const uint32_t bit = f & 31u;
                        if ((word >> bit) & 1u) {
                            const uint32_t row = static_cast<uint32_t>(pre_s[f >> 5]) + __popc(word & ((1u << bit) - 1u));
                            range = static_cast<uint32_t>(rp_s[row]) | (static_cast<uint32_t>(rp_s[row + 1]) << 16);
                        }
                    }
                }
                if ((range >> 16) > (range & 0xFFFFu)) {
                    const float x = my_val[k];
                    my_idx[cnt] = range;
                    my_val[cnt] = x;
                    ++cnt;
                }
            }
        }
        // phase 2: ONE entry per iteration and lane (the hit rows of the round are walked as one flat entry stream, so lanes
        // whose rows have different lengths stay busy), in feature order, into this lane's accumulators
        if (cnt) {
            uint32_t i = 1;
            uint32_t range = my_idx[0];
            float x = my_val[0];
            uint32_t e = range & 0xFFFFu, ee = range >> 16;
            if (STATS) st_ent += ee - e;
            for (;;) {
                float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
                *a = __fadd_rn(*a, __fmul_rn(x, ew_s[e]));
                if (++e == ee) {
                    if (i == cnt) break;
                    range = my_idx[i];
                    x = my_val[i];
                    ++i;
                    e = range & 0xFFFFu;
                    ee = range >> 16;
                    if (STATS) st_ent += ee - e;
                }
            }
        }
        if (STATS) st_match += cnt;
        __syncwarp();
    }
    if (have && (h.has_bias & 1u)) {  // bias row last (inference.hpp:806-811)
        apply_row(bias_range, L.bias);
        if (STATS) { st_match += 1; st_ent += (bias_range >> 16) - (bias_range & 0xFFFFu); }
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
