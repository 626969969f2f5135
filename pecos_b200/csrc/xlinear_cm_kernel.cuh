// Chunk-major scoring: the (query, beam slot) pairs of a layer are bucketed by weight chunk, a CTA stages ONE chunk in
// shared memory and every LANE walks ITS OWN pair.
//
// Included by xlinear_engine.cu (inside its anonymous namespace, after the small device helpers).
//
// Why (profiles/r02_*): the query-major kernels spend 600 - 3,900 warp-instructions per pair on match compaction, prefix
// sums and on putting colliding entries of a 32-entry group back into feature order, and every probe / extent / entry
// access is a scattered global load (one L1 line each).  Here
//   * at LOAD time every chunk of an eligible layer is packed into a self-contained IMAGE in HBM (xl_cm_build_images_kernel):
//     header | lookup structure | entry weights (f32) | entry columns (u8).  Lookup structure: for feature spaces up to
//     kCmDirectRows a direct table feature -> {first entry, end} (u16 | u16 << 16); else the chunk's feature map (one bit
//     per feature + a 16-bit row prefix per 32 features: "is f a row, and which one" = one shared-memory word + popcount)
//     and 16-bit row pointers;
//   * per call the pairs are bucketed by chunk on the device (count -> scan -> scatter; the reference's b_sort_by_chunk,
//     pecos/core/xmc/inference.hpp:985-993);
//   * the score kernel is PERSISTENT: one CTA per SM takes a contiguous 1/grid share of the chunk-sorted pair list, so it
//     meets only a few chunks; a chunk's image arrives by ONE bulk asynchronous copy (cp.async.bulk + mbarrier, the TMA
//     engine's non-tensor form) and serves every pair of the run; warps take 32-pair slices of the run;
//   * a lane walks its pair's query features in ascending order (staged global -> shared by cp.async, two rounds of 8 in
//     flight), compacts the hits of a round in place as {entry range, x}, then streams the hit rows' entries -- ONE entry
//     per iteration, warp-uniform trip count, so rows of different lengths do not serialise -- into the lane's PRIVATE
//     accumulators acc[column][lane].  That is the reference's marching loop (inference.hpp:788-811): ascending feature
//     order, separate multiply and add, bias row last; the order is right by construction and a column is only ever
//     touched by the lane that owns the pair: no compaction across lanes, no conflict resolution.
//
// Bit-identical to the query-major kernels (tests/test_chunk_major_gpu.py).  Eligibility: shape (cm_shape, at load: width
// <= 256, rows / entries per chunk < 65535, image + 4 warps fit in shared memory, images within PB200_CMIMG_MB) and call
// (cm_plan: sparse queries, enough pairs per chunk and per SM).
#pragma once

constexpr int kCmFeat = 8;                  // query features staged per pair and round (two rounds in flight per warp)
constexpr int kCmMaxWarps = 16;
constexpr int kCmMinWarps = 4;
constexpr uint32_t kCmMaxDup = 400;         // a column cap may at most quadruple the (virtual) chunks of a layer
constexpr uint32_t kCmSmemBudget = 224u << 10;  // dynamic shared memory a CTA may take (227 KB is the sm_100a maximum)
constexpr uint32_t kCmMinReuse = 24;        // average pairs per chunk below which the per-chunk staging does not pay
constexpr uint32_t kCmMinPairs = 148u * 48u; // fewer pairs than this: the query-major kernels fill the GPU better
constexpr uint32_t kCmDirectRows = 16384;   // feature spaces up to this size get a direct feature -> entry-range table
constexpr uint32_t kCmEmpty = 0xFFFFFFFFu;

struct CmWork {
    uint32_t* slot_pos;     // [rows x beam_stride] first candidate position of every beam slot
    uint32_t* count;        // [n_chunks] pairs per chunk, reused as the scatter cursor
    uint32_t* bucket_ptr;   // [n_chunks + 1]
    uint32_t* item_ptr;     // [n_chunks + 1] (unused by the persistent score kernel; kept by the scan)
    uint32_t* pair_q;       // [pairs] query of a pair, grouped by chunk
    uint32_t* pair_pos;     // [pairs] candidate position of the pair's first column inside the query's row
    uint32_t item_pairs;
};

struct CmPlan {  // per call
    bool eligible = false;
    uint32_t warps = 0, grid = 0;
    size_t smem = 0;
};

__host__ __device__ inline uint32_t cm_align16(uint32_t x) { return (x + 15u) & ~15u; }

__host__ __device__ inline size_t cm_warp_bytes(uint32_t acc_cols, uint32_t stages) {
    return static_cast<size_t>(stages) * 32 * (kCmFeat + 1) * 8   // staging ring: query features / compacted hits, stride 9
           + static_cast<size_t>(acc_cols) * 32 * 4;              // accumulators [col][lane]
}

// col_cap: a chunk wider than col_cap columns is cut into ceil(n_cols / col_cap) column ranges of (nearly) equal width, each
// with its OWN image (only its entries) -- a "virtual chunk"; a (query, chunk) pair is then scored once per range.  The lookups
// are repeated for the cut chunks, the accumulate work is not, and both the image and the per-warp accumulators shrink to the
// cap, so more warps fit (occupancy is what the kernel is short of on wide chunks: profiles/r02_e, r02_i).  The cap is chosen
// at load so that only the few widest chunks of a layer are cut (cm_choose_cap).  e_max = most entries of one virtual chunk,
// n_vc = number of virtual chunks.
inline CmShape cm_shape(uint32_t fm_words, uint32_t w_rows, uint32_t r_max, uint32_t e_max, uint32_t col_cap, uint32_t n_chunks,
                        uint32_t n_vc) {
    CmShape s;
    if (fm_words == 0 || n_chunks == 0 || col_cap == 0 || col_cap > 256u || r_max >= 65535u || e_max >= 65535u) return s;
    s.direct = w_rows <= kCmDirectRows;
    s.words = s.direct ? w_rows : fm_words;
    s.col_cap = col_cap;
    s.n_vc = n_vc;
    s.r_cap = r_max; s.e_cap = e_max; s.acc_cols = col_cap;
    // narrow chunks do little arithmetic per round of query features: keep four rounds of cp.async in flight per warp to
    // cover the global-memory latency; wide chunks (long accumulate phases) get by with two
    s.stages = s.acc_cols <= 16u ? 4u : 2u;
    uint32_t off = 16;  // header {bias range, n_cols, R, E}
    s.off_lookup = off; off += cm_align16(s.words * 4u);
    if (!s.direct) {
        s.off_pre = off; off += cm_align16(s.words * 2u);
        s.off_rp = off;  off += cm_align16((r_max + 2u) * 2u);
    }
    s.off_ew = off; off += cm_align16((e_max + 1u) * 4u);
    s.off_ec = off; off += cm_align16(e_max + 1u);
    s.img_bytes = (off + 127u) & ~127u;
    if (s.img_bytes + kCmMinWarps * cm_warp_bytes(s.acc_cols, s.stages) + 64 > kCmSmemBudget) return s;
    s.warps_fit = static_cast<uint32_t>(std::min<size_t>(kCmMaxWarps, (kCmSmemBudget - s.img_bytes - 64) / cm_warp_bytes(s.acc_cols, s.stages)));
    s.ok = true;
    return s;
}

// force: take the kernel wherever the layer has images, ignoring the reuse / occupancy heuristics (kernel mode 5, tests)
inline CmPlan cm_plan(const CmShape& s, uint32_t n_chunks, uint64_t pairs, uint32_t n_sm, bool force) {
    CmPlan p;
    if (!s.ok || pairs == 0) return p;
    // measured (profiles/r02_d, r02_e): with the 94 KB feature-map image of a large feature space the kernel does not beat the
    // query-major kernels (S layers 1-4: 2.0 / 2.4 / 6.3 ms vs 1.9 / 1.9 / 2.0 ms) -- only direct-table layers take it by default
    if (!force && !s.direct) return p;
    pairs = pairs * s.n_vc / std::max<uint32_t>(n_chunks, 1u);  // cut chunks are visited once per column range
    n_chunks = s.n_vc;
    if (!force && (pairs < static_cast<uint64_t>(kCmMinReuse) * n_chunks || pairs < kCmMinPairs)) return p;
    const size_t per_warp = cm_warp_bytes(s.acc_cols, s.stages);
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(kCmMaxWarps, (kCmSmemBudget - s.img_bytes - 64) / per_warp));
    // no point in more lanes than a CTA's share of the pair list holds
    const uint64_t share = (pairs + n_sm - 1) / n_sm;
    while (warps > kCmMinWarps && static_cast<uint64_t>(warps - 1) * 32 >= share) --warps;
    p.eligible = true;
    p.warps = warps;
    p.grid = n_sm;
    p.smem = s.img_bytes + warps * per_warp + 64;
    return p;
}

__device__ __forceinline__ void cm_cp_async4(void* smem_dst, const void* gmem_src) {
    const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cm_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cm_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// bulk asynchronous copy global -> shared (TMA engine, non-tensor form; SASS UBLKCP) completing on an mbarrier
__device__ __forceinline__ void cm_mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void cm_bulk_load(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void cm_mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}

// LOAD TIME: one CTA per VIRTUAL chunk (chunk p, column range h; S.vc_ptr[p] = first virtual chunk of chunk p) packs its image (see CmShape) from the layer's
// device arrays: the rows' entries whose column falls into the range (contiguous inside a row: entries are stored in
// ascending column order), columns re-based to the range.
__global__ void __launch_bounds__(256)
xl_cm_build_images_kernel(const LayerDev L, const CmShape S, unsigned char* __restrict__ images) {
    const uint32_t vc = blockIdx.x;
    uint32_t c;
    {
        uint32_t lo = 0, hi = L.n_chunks;  // largest c with vc_ptr[c] <= vc
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (S.vc_ptr[mid] <= vc) lo = mid; else hi = mid;
        }
        c = lo;
    }
    const uint32_t hh = vc - S.vc_ptr[c];
    const uint32_t n_ranges = S.vc_ptr[c + 1] - S.vc_ptr[c];
    const ChunkHeader h = L.chunks[c];
    unsigned char* img = images + static_cast<uint64_t>(vc) * S.img_bytes;
    uint32_t* hdr = reinterpret_cast<uint32_t*>(img);
    uint32_t* lookup = reinterpret_cast<uint32_t*>(img + S.off_lookup);
    float* ew = reinterpret_cast<float*>(img + S.off_ew);
    unsigned char* ec = img + S.off_ec;
    for (uint32_t i = threadIdx.x; i < S.img_bytes / 4u; i += blockDim.x) reinterpret_cast<uint32_t*>(img)[i] = 0u;
    __syncthreads();
    if (h.has_bias & kChunkAbsent) return;
    if (n_ranges == 0) return;
    const uint32_t width = (h.n_cols + n_ranges - 1u) / n_ranges;  // columns per range of THIS chunk
    const uint32_t lo = hh * width;
    if (lo >= h.n_cols) return;                                    // empty range: no pair is ever bucketed here
    const uint32_t hi = min(h.n_cols, lo + width);
    const uint32_t R = h.nnz_rows;
    const uint32_t R4 = (R + 3u) & ~3u;
    const uint32_t* ridx = L.meta + h.meta_off;
    const uint32_t* rp = ridx + R4;
    const uint2* ent = L.entries + h.ent_off;
    unsigned short* rps = S.direct ? nullptr : reinterpret_cast<unsigned short*>(img + S.off_rp);
    if (!S.direct) {
        unsigned short* pre = reinterpret_cast<unsigned short*>(img + S.off_pre);
        const uint2* fm = L.featmap + static_cast<uint64_t>(c) * L.fm_words;
        for (uint32_t i = threadIdx.x; i < S.words; i += blockDim.x) {
            const uint2 cell = fm[i];
            lookup[i] = cell.x;
            pre[i] = static_cast<unsigned short>(cell.y);
        }
    }
    // rows in blocks of 256: sub-range of the row inside [lo, hi), exclusive scan of the lengths, copy
    __shared__ uint32_t s_scan[256];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t r0 = 0; r0 < R; r0 += 256u) {
        const uint32_t r = r0 + threadIdx.x;
        uint32_t b = 0, e = 0;
        if (r < R) {
            b = rp[r];
            e = rp[r + 1];
            while (b < e && ent[b].x < lo) ++b;
            uint32_t t = b;
            while (t < e && ent[t].x < hi) ++t;
            e = t;
        }
        const uint32_t len = e - b;
        s_scan[threadIdx.x] = len;
        __syncthreads();
        for (uint32_t d = 1; d < 256u; d <<= 1) {  // Hillis-Steele inclusive scan
            const uint32_t v = (threadIdx.x >= d) ? s_scan[threadIdx.x - d] : 0u;
            __syncthreads();
            s_scan[threadIdx.x] += v;
            __syncthreads();
        }
        const uint32_t base = s_carry + s_scan[threadIdx.x] - len;
        if (r < R) {
            for (uint32_t i = 0; i < len; ++i) {
                const uint2 en = ent[b + i];
                ec[base + i] = static_cast<unsigned char>(en.x - lo);
                ew[base + i] = __uint_as_float(en.y);
            }
            if (S.direct) {
                const uint32_t f = ridx[r];
                if (f < S.words) lookup[f] = base | ((base + len) << 16);  // an empty row reads as "no row": nothing to add
            } else {
                rps[r] = static_cast<unsigned short>(base);
            }
            if (r + 1u == R) {
                if (!S.direct) rps[R] = static_cast<unsigned short>(base + len);
                hdr[0] = (h.has_bias & 1u) ? (base | ((base + len) << 16)) : 0u;  // the bias row is the chunk's last row
                hdr[3] = base + len;
            }
        }
        __syncthreads();
        if (threadIdx.x == 255u) s_carry += s_scan[255];
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        hdr[1] = hi - lo;
        hdr[2] = R;
    }
}

// one warp per query: candidate position of every beam slot (prefix of the chunk widths) and pairs per chunk
__global__ void __launch_bounds__(128)
xl_cm_count_kernel(const LayerDev L, const QueryDev X, const uint32_t* __restrict__ beam_id,
                   const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, const uint32_t rows, CmWork w,
                   const uint32_t* __restrict__ vc_ptr) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= rows) return;
    const uint32_t cnt = beam_cnt[q];
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
            if (scored) {  // one pair per non-empty column range of the chunk
                const uint32_t v0 = vc_ptr ? vc_ptr[p] : p, nr = vc_ptr ? vc_ptr[p + 1] - v0 : 1u;  // nullptr: no column ranges
                const uint32_t cw = (width + nr - 1u) / max(nr, 1u);
                for (uint32_t hh = 0; hh < nr && hh * cw < width; ++hh) atomicAdd(&w.count[v0 + hh], 1u);
            }
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
                     const uint32_t beam_stride, const uint32_t rows, CmWork w, const uint32_t* __restrict__ vc_ptr) {
    const int lane = threadIdx.x & 31;
    const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= rows) return;
    const uint32_t cnt = beam_cnt[q];
    for (uint32_t j = lane; j < cnt; j += 32) {
        const uint32_t p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
        const uint4 h = *reinterpret_cast<const uint4*>(&L.chunks[p]);
        if ((h.w & kChunkAbsent) || h.y == 0) continue;
        const uint32_t pos = w.slot_pos[static_cast<uint64_t>(q) * beam_stride + j];
        const uint32_t v0 = vc_ptr ? vc_ptr[p] : p, nr = vc_ptr ? vc_ptr[p + 1] - v0 : 1u;
        const uint32_t cw = (h.y + nr - 1u) / max(nr, 1u);
        for (uint32_t hh = 0; hh < nr && hh * cw < h.y; ++hh) {
            const uint32_t at = atomicAdd(&w.count[v0 + hh], 1u);
            w.pair_q[at] = q;
            w.pair_pos[at] = pos + hh * cw;  // first candidate of this column range
        }
    }
}

template <bool DIRECT, int STAGES>
__global__ void __launch_bounds__(kCmMaxWarps * 32)
xl_cm_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, const CmShape S, const unsigned char* __restrict__ images,
                    float* __restrict__ cand, const uint64_t cand_stride_q) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* img = smem_raw;                                   // the staged image of one virtual chunk
    // the image is only ever written by the bulk copy (async proxy), never by this kernel's stores: __restrict__ lets the
    // compiler hoist its loads above the accumulator stores
    const uint32_t* __restrict__ hdr_s = reinterpret_cast<const uint32_t*>(img);
    const uint32_t* __restrict__ look_s = reinterpret_cast<const uint32_t*>(img + S.off_lookup);   // direct table, or feature-map bits
    const unsigned short* __restrict__ pre_s = reinterpret_cast<const unsigned short*>(img + S.off_pre);
    const unsigned short* __restrict__ rp_s = reinterpret_cast<const unsigned short*>(img + S.off_rp);
    const float* __restrict__ ew_s = reinterpret_cast<const float*>(img + S.off_ew);
    const unsigned char* __restrict__ ec_s = img + S.off_ec;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    constexpr int kStride = kCmFeat + 1;
    constexpr int kBuf = 32 * kStride;                               // words per staging array
    unsigned char* mine = smem_raw + S.img_bytes + static_cast<size_t>(warp) * cm_warp_bytes(S.acc_cols, STAGES);
    uint32_t* st_idx = reinterpret_cast<uint32_t*>(mine);            // [STAGES][32][kStride]
    float* st_val = reinterpret_cast<float*>(st_idx + STAGES * kBuf); // [STAGES][32][kStride]
    float* my_acc = st_val + STAGES * kBuf + lane;                   // [acc_cols][32], this lane's column of it

    __shared__ __align__(8) unsigned long long s_mbar;
    const uint32_t mbar = static_cast<uint32_t>(__cvta_generic_to_shared(&s_mbar));
    if (threadIdx.x == 0) {
        cm_mbar_init(mbar, 1u);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();

    // ---- this CTA's contiguous share of the (virtual-)chunk-sorted pair list
    const uint32_t n_vc = S.n_vc;
    const uint64_t P = w.bucket_ptr[n_vc];
    const uint32_t begin = static_cast<uint32_t>(P * blockIdx.x / gridDim.x);
    const uint32_t end = static_cast<uint32_t>(P * (blockIdx.x + 1ull) / gridDim.x);
    if (begin >= end) return;
    uint32_t c;
    {
        uint32_t lo = 0, hi = n_vc;  // largest c with bucket_ptr[c] <= begin (then skip empty buckets forward)
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (w.bucket_ptr[mid] <= begin) lo = mid; else hi = mid;
        }
        c = lo;
    }
    // staging geometry: one warp-wide cp.async instruction copies kCmFeat consecutive features of kPerIter pairs; this lane
    // always serves feature fl of the pairs sub, sub + kPerIter, ...
    constexpr int kPerIter = 32 / kCmFeat;
    constexpr int kIters = 32 / kPerIter;
    const int sub = lane / kCmFeat, fl = lane % kCmFeat;
    uint32_t parity = 0;

    for (uint32_t i = begin; i < end;) {
        while (w.bucket_ptr[c + 1] <= i) ++c;                        // virtual chunk holding pair i
        const uint32_t run_end = min(end, w.bucket_ptr[c + 1]);
        // ---- stage the image: ONE bulk copy (every warp has left the previous image: barrier first)
        __syncthreads();
        if (threadIdx.x == 0) cm_bulk_load(static_cast<uint32_t>(__cvta_generic_to_shared(img)), images + static_cast<uint64_t>(c) * S.img_bytes, S.img_bytes, mbar);
        cm_mbar_wait(mbar, parity);
        parity ^= 1u;
        const uint32_t bias_range = hdr_s[0];
        const uint32_t n_cols = hdr_s[1];

        // ---- warps take 32-pair slices of the run
        for (uint32_t s0 = i + static_cast<uint32_t>(warp) * 32u; s0 < run_end; s0 += static_cast<uint32_t>(nwarps) * 32u) {
            const uint32_t pidx = s0 + lane;
            const bool have = pidx < run_end;
            uint32_t q = 0, pos = 0, qn = 0;
            uint64_t qb = 0;
            if (have) {
                q = w.pair_q[pidx];
                pos = w.pair_pos[pidx];
                qb = X.row_ptr[q] - X.nnz_base;
                qn = static_cast<uint32_t>(X.row_ptr[q + 1] - X.nnz_base - qb);
            }
            const uint32_t qn_max = __reduce_max_sync(kFull, qn);
            // the pairs this lane copies for: source pointers (feature fl of the pair's row) and row lengths, once per slice
            uint32_t src_o[kIters], src_n[kIters];  // offsets fit 32 bits: a tile of queries holds < 2^32 non-zeros
#pragma unroll
            for (int it = 0; it < kIters; ++it) {
                const int pi = it * kPerIter + sub;
                src_o[it] = static_cast<uint32_t>(__shfl_sync(kFull, qb, pi)) + fl;
                src_n[it] = __shfl_sync(kFull, qn, pi);
            }
            // query features travel global -> shared by cp.async through a ring of STAGES buffers: rounds r + 1 .. r + STAGES - 1
            // are in flight while round r is processed.  Row i of a buffer = the next kCmFeat features of the slice's pair i
            // (stride 9 words: the lane-per-row reads are bank-conflict free).  A (possibly empty) group is committed for every
            // round slot, so "all but the newest STAGES - 1 groups are complete" always means "round r has landed".
            auto stage_round = [&](uint32_t t0, int buf) {
                if (t0 < qn_max) {
                    uint32_t* di = st_idx + buf * kBuf + sub * kStride + fl;
                    float* dv = st_val + buf * kBuf + sub * kStride + fl;
#pragma unroll
                    for (int it = 0; it < kIters; ++it) {
                        if (t0 + fl < src_n[it]) {
                            cm_cp_async4(di + it * kPerIter * kStride, X.col_idx + src_o[it] + t0);
                            cm_cp_async4(dv + it * kPerIter * kStride, X.val + src_o[it] + t0);
                        }
                    }
                }
                cm_cp_async_commit();
            };
            __syncwarp();
#pragma unroll
            for (int st = 0; st < STAGES - 1; ++st) stage_round(static_cast<uint32_t>(st) * kCmFeat, st);
            for (uint32_t col = 0; col < n_cols; ++col) my_acc[col * 32] = 0.0f;
            uint32_t prev_f = kCmEmpty;
            int buf = 0;
            for (uint32_t t0 = 0; t0 < qn_max; t0 += kCmFeat, buf = (buf + 1 == STAGES) ? 0 : buf + 1) {
                stage_round(t0 + (STAGES - 1) * kCmFeat, (buf + STAGES - 1) % STAGES);  // refills the buffer consumed last round
                cm_cp_async_wait<STAGES - 1>();
                __syncwarp();
                uint32_t* my_idx = st_idx + buf * kBuf + lane * kStride;
                float* my_val = st_val + buf * kBuf + lane * kStride;
                const uint32_t n_here = (qn > t0) ? min(static_cast<uint32_t>(kCmFeat), qn - t0) : 0u;
                // phase 1: look the features up -- all loads of the round are issued before the first store (eight independent
                // feature / lookup / value loads in flight per lane) -- then compact the hits IN PLACE as {entry range, x}
                uint32_t fq[kCmFeat], rq[kCmFeat];
                float xq[kCmFeat];
#pragma unroll
                for (int k = 0; k < kCmFeat; ++k) {
                    fq[k] = (static_cast<uint32_t>(k) < n_here) ? my_idx[k] : kCmEmpty;
                    xq[k] = my_val[k];
                }
#pragma unroll
                for (int k = 0; k < kCmFeat; ++k) {
                    const uint32_t f = fq[k];
                    const bool dup = (f == prev_f);  // a repeated column index only counts once (the first occurrence)
                    if (static_cast<uint32_t>(k) < n_here) prev_f = f;
                    uint32_t range = 0;
                    if (static_cast<uint32_t>(k) < n_here && !dup && f < L.w_rows) {
                        if (DIRECT) {
                            range = look_s[f];
                        } else {
                            const uint32_t word = look_s[f >> 5];
                            const uint32_t bit = f & 31u;
                            if ((word >> bit) & 1u) {
                                const uint32_t row = static_cast<uint32_t>(pre_s[f >> 5]) + __popc(word & ((1u << bit) - 1u));
                                range = static_cast<uint32_t>(rp_s[row]) | (static_cast<uint32_t>(rp_s[row + 1]) << 16);
                            }
                        }
                    }
                    rq[k] = range;
                }
                uint32_t cnt = 0;
#pragma unroll
                for (int k = 0; k < kCmFeat; ++k) {
                    if (static_cast<int>(rq[k] >> 16) > static_cast<int>(rq[k] & 0xFFFFu)) {
                        my_idx[cnt] = rq[k];
                        my_val[cnt] = xq[k];
                        ++cnt;
                    }
                }
                // phase 2: the hit rows' entries, in feature order, into this lane's accumulators
                if (L.has_dup_cols) {
                    // a row may repeat a column (non-canonical W): strictly one entry at a time
                    for (uint32_t hi = 0; hi < cnt; ++hi) {
                        const uint32_t range = my_idx[hi];
                        const float x = my_val[hi];
                        const uint32_t ee = range >> 16;
                        for (uint32_t e = range & 0xFFFFu; e < ee; ++e) {
                            float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
                            *a = __fadd_rn(*a, __fmul_rn(x, ew_s[e]));
                        }
                    }
                } else {
                    // the columns of ONE row are distinct, so its entries are independent: four at a time -- their column /
                    // weight loads, accumulator loads, multiply-adds and stores overlap instead of forming one dependent chain
                    // per entry.  Rows stay in feature order (program order of the accumulator accesses).
                    for (uint32_t hi = 0; hi < cnt; ++hi) {
                        const uint32_t range = my_idx[hi];
                        const float x = my_val[hi];
                        const uint32_t ee = range >> 16;
                        for (uint32_t e = range & 0xFFFFu; e < ee; e += 4u) {
                            const uint32_t n = ee - e;  // >= 1
                            const uint32_t c0 = ec_s[e];
                            const uint32_t c1 = ec_s[e + (n > 1u ? 1u : 0u)];
                            const uint32_t c2 = ec_s[e + (n > 2u ? 2u : 0u)];
                            const uint32_t c3 = ec_s[e + (n > 3u ? 3u : 0u)];
                            const float w0 = ew_s[e];
                            const float w1 = ew_s[e + (n > 1u ? 1u : 0u)];
                            const float w2 = ew_s[e + (n > 2u ? 2u : 0u)];
                            const float w3 = ew_s[e + (n > 3u ? 3u : 0u)];
                            float* a0 = my_acc + c0 * 32u;
                            float* a1 = my_acc + c1 * 32u;
                            float* a2 = my_acc + c2 * 32u;
                            float* a3 = my_acc + c3 * 32u;
                            const float v0 = *a0, v1 = *a1, v2 = *a2, v3 = *a3;
                            *a0 = __fadd_rn(v0, __fmul_rn(x, w0));
                            if (n > 1u) *a1 = __fadd_rn(v1, __fmul_rn(x, w1));
                            if (n > 2u) *a2 = __fadd_rn(v2, __fmul_rn(x, w2));
                            if (n > 3u) *a3 = __fadd_rn(v3, __fmul_rn(x, w3));
                        }
                    }
                }
                __syncwarp();
            }
            cm_cp_async_wait<0>();
            if (have && bias_range) {  // bias row last (inference.hpp:806-811)
                const uint32_t ee = bias_range >> 16;
                for (uint32_t e = bias_range & 0xFFFFu; e < ee; ++e) {
                    float* a = my_acc + static_cast<uint32_t>(ec_s[e]) * 32u;
                    *a = __fadd_rn(*a, __fmul_rn(L.bias, ew_s[e]));
                }
            }
            if (have) {
                float* dst = cand + static_cast<uint64_t>(q) * cand_stride_q + pos;
                for (uint32_t col = 0; col < n_cols; ++col) dst[col] = my_acc[col * 32];
            }
        }
        i = run_end;
    }
}


// ------------------------------------------------------------------------------------------------------------------
// Lane-per-pair scoring WITHOUT a staged image, for layers whose chunks cannot be staged (large feature spaces: the feature map
// alone is 62 KB of bits per chunk; chunks visited by few pairs: the 3M-label leaf has ~61 pairs per chunk).  Same walk as
// xl_cm_scores_kernel -- pairs in chunk order, a lane marches through ITS OWN pair, private accumulators acc[col][lane], no
// cross-lane compaction, no conflict rounds -- but the lookups (feature-map cell), the matched rows' extents and their entries
// are read from global memory: because the pairs are chunk-sorted, the 32 lanes of a warp and the neighbouring warps work on
// the same chunk, so those reads hit L1 / L2 (the query-major kernel sends every probe of such a layer to DRAM).  Every load
// phase of a round is issued for all of the round's features / hits before the first use (8 independent loads in flight per
// lane).  No CTA-wide state: warps are fully independent (no barrier, no staging of an image).
// ------------------------------------------------------------------------------------------------------------------
struct CmgPlan {
    bool eligible = false;
    uint32_t warps = 0, grid = 0;
    size_t smem = 0;
};

constexpr int kCmgStages = 3;

inline CmgPlan cmg_plan(uint32_t c_max, uint32_t e_max, uint32_t n_chunks, uint64_t pairs, uint32_t n_sm, bool force) {
    CmgPlan p;
    if (c_max == 0 || c_max > 256u || e_max >= 65535u || pairs == 0 || n_chunks == 0) return p;
    if (!force && pairs < kCmMinPairs) return p;
    const size_t per_warp = cm_warp_bytes(c_max, kCmgStages);
    uint32_t warps = static_cast<uint32_t>(std::min<size_t>(kCmMaxWarps, (kCmSmemBudget - 64) / per_warp));
    if (warps < 2) return p;
    p.eligible = true;
    p.warps = warps;
    p.smem = warps * per_warp + 64;
    p.grid = n_sm * (p.smem <= (110u << 10) ? 2u : 1u);  // warps are independent: two CTAs per SM when they fit
    return p;
}

__global__ void __launch_bounds__(kCmMaxWarps * 32)
xl_cmg_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, float* __restrict__ cand, const uint64_t cand_stride_q,
                     const uint32_t acc_cols) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int STAGES = kCmgStages;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    constexpr int kStride = kCmFeat + 1;
    constexpr int kBuf = 32 * kStride;
    unsigned char* mine = smem_raw + static_cast<size_t>(warp) * cm_warp_bytes(acc_cols, STAGES);
    uint32_t* st_idx = reinterpret_cast<uint32_t*>(mine);
    float* st_val = reinterpret_cast<float*>(st_idx + STAGES * kBuf);
    float* my_acc = st_val + STAGES * kBuf + lane;

    // ---- this WARP's contiguous share of the chunk-sorted pair list (whole 32-pair slices, never across a chunk boundary)
    const uint32_t n_chunks = L.n_chunks;
    const uint64_t P = w.bucket_ptr[n_chunks];
    const uint64_t gw = static_cast<uint64_t>(blockIdx.x) * nwarps + warp, tw = static_cast<uint64_t>(gridDim.x) * nwarps;
    const uint32_t begin = static_cast<uint32_t>(P * gw / tw);
    const uint32_t end = static_cast<uint32_t>(P * (gw + 1ull) / tw);
    if (begin >= end) return;
    uint32_t c;
    {
        uint32_t lo = 0, hi = n_chunks;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (w.bucket_ptr[mid] <= begin) lo = mid; else hi = mid;
        }
        c = lo;
    }
    constexpr int kPerIter = 32 / kCmFeat;
    constexpr int kIters = 32 / kPerIter;
    const int sub = lane / kCmFeat, fl = lane % kCmFeat;

    for (uint32_t s0 = begin; s0 < end;) {
        while (w.bucket_ptr[c + 1] <= s0) ++c;
        const uint32_t run_end = min(end, w.bucket_ptr[c + 1]);
        const uint32_t s1 = min(s0 + 32u, run_end);                  // this slice: pairs [s0, s1) of chunk c
        const ChunkHeader h = L.chunks[c];
        const uint2* __restrict__ fm_g = L.featmap + static_cast<uint64_t>(c) * L.fm_words;
        const uint2* __restrict__ ext_g = reinterpret_cast<const uint2*>(L.rowext + h.meta_off);
        const uint2* __restrict__ ent_g = L.entries + h.ent_off;
        const uint32_t n_cols = h.n_cols;
        const uint32_t R = h.nnz_rows;

        const uint32_t pidx = s0 + lane;
        const bool have = pidx < s1;
        uint32_t q = 0, pos = 0, qn = 0;
        uint64_t qb = 0;
        if (have) {
            q = w.pair_q[pidx];
            pos = w.pair_pos[pidx];
            qb = X.row_ptr[q] - X.nnz_base;
            qn = static_cast<uint32_t>(X.row_ptr[q + 1] - X.nnz_base - qb);
        }
        const uint32_t qn_max = __reduce_max_sync(kFull, qn);
        uint32_t src_o[kIters], src_n[kIters];
#pragma unroll
        for (int it = 0; it < kIters; ++it) {
            const int pi = it * kPerIter + sub;
            src_o[it] = static_cast<uint32_t>(__shfl_sync(kFull, qb, pi)) + fl;
            src_n[it] = __shfl_sync(kFull, qn, pi);
        }
        auto stage_round = [&](uint32_t t0, int buf) {
            if (t0 < qn_max) {
                uint32_t* di = st_idx + buf * kBuf + sub * kStride + fl;
                float* dv = st_val + buf * kBuf + sub * kStride + fl;
#pragma unroll
                for (int it = 0; it < kIters; ++it) {
                    if (t0 + fl < src_n[it]) {
                        cm_cp_async4(di + it * kPerIter * kStride, X.col_idx + src_o[it] + t0);
                        cm_cp_async4(dv + it * kPerIter * kStride, X.val + src_o[it] + t0);
                    }
                }
            }
            cm_cp_async_commit();
        };
        __syncwarp();
#pragma unroll
        for (int st = 0; st < STAGES - 1; ++st) stage_round(static_cast<uint32_t>(st) * kCmFeat, st);
        for (uint32_t col = 0; col < n_cols; ++col) my_acc[col * 32] = 0.0f;
        uint32_t prev_f = kCmEmpty;
        int buf = 0;
        for (uint32_t t0 = 0; t0 < qn_max; t0 += kCmFeat, buf = (buf + 1 == STAGES) ? 0 : buf + 1) {
            stage_round(t0 + (STAGES - 1) * kCmFeat, (buf + STAGES - 1) % STAGES);
            cm_cp_async_wait<STAGES - 1>();
            __syncwarp();
            uint32_t* my_idx = st_idx + buf * kBuf + lane * kStride;
            float* my_val = st_val + buf * kBuf + lane * kStride;
            const uint32_t n_here = (qn > t0) ? min(static_cast<uint32_t>(kCmFeat), qn - t0) : 0u;
            // phase 1a: the round's features and their feature-map cells (global, independent loads)
            uint32_t fq[kCmFeat];
            float xq[kCmFeat];
            uint2 cell[kCmFeat];
#pragma unroll
            for (int k = 0; k < kCmFeat; ++k) {
                fq[k] = (static_cast<uint32_t>(k) < n_here) ? my_idx[k] : kCmEmpty;
                xq[k] = my_val[k];
            }
#pragma unroll
            for (int k = 0; k < kCmFeat; ++k) {
                const uint32_t f = fq[k];
                const bool dup = (f == prev_f);  // a repeated column index only counts once (the first occurrence)
                if (static_cast<uint32_t>(k) < n_here) prev_f = f;
                const bool live = static_cast<uint32_t>(k) < n_here && !dup && f < L.w_rows && R > 0;
                cell[k] = live ? __ldg(fm_g + (f >> 5)) : make_uint2(0u, 0u);
                if (!live) fq[k] = 0u;  // bit 0 of an all-zero cell: no hit
            }
            // phase 1b: rows of the hits, compacted in place as {row, x}
            uint32_t cnt = 0;
#pragma unroll
            for (int k = 0; k < kCmFeat; ++k) {
                const uint32_t bit = fq[k] & 31u;
                if ((cell[k].x >> bit) & 1u) {
                    my_idx[cnt] = cell[k].y + __popc(cell[k].x & ((1u << bit) - 1u));
                    my_val[cnt] = xq[k];
                    ++cnt;
                }
            }
            // phase 1c: the hit rows' extents (independent loads), stored back as packed ranges
#pragma unroll
            for (int k0 = 0; k0 < kCmFeat; k0 += 4) {
                uint2 ex[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ex[u] = (static_cast<uint32_t>(k0 + u) < cnt) ? __ldg(ext_g + my_idx[k0 + u]) : make_uint2(0u, 0u);
#pragma unroll
                for (int u = 0; u < 4; ++u) if (static_cast<uint32_t>(k0 + u) < cnt) my_idx[k0 + u] = ex[u].x | (ex[u].y << 16);
            }
            // phase 2: entries of the hit rows, in feature order; four independent entries of a row at a time (a row's columns
            // are distinct unless the model is non-canonical: then strictly one at a time)
            for (uint32_t hi = 0; hi < cnt; ++hi) {
                const uint32_t range = my_idx[hi];
                const float x = my_val[hi];
                const uint32_t ee = range >> 16;
                if (L.has_dup_cols) {
                    for (uint32_t e = range & 0xFFFFu; e < ee; ++e) {
                        const uint2 en = __ldg(ent_g + e);
                        float* a = my_acc + en.x * 32u;
                        *a = __fadd_rn(*a, __fmul_rn(x, __uint_as_float(en.y)));
                    }
                    continue;
                }
                for (uint32_t e = range & 0xFFFFu; e < ee; e += 4u) {
                    const uint32_t n = ee - e;
                    const uint2 e0 = __ldg(ent_g + e);
                    const uint2 e1 = __ldg(ent_g + e + (n > 1u ? 1u : 0u));
                    const uint2 e2 = __ldg(ent_g + e + (n > 2u ? 2u : 0u));
                    const uint2 e3 = __ldg(ent_g + e + (n > 3u ? 3u : 0u));
                    float* a0 = my_acc + e0.x * 32u;
                    float* a1 = my_acc + e1.x * 32u;
                    float* a2 = my_acc + e2.x * 32u;
                    float* a3 = my_acc + e3.x * 32u;
                    const float v0 = *a0, v1 = *a1, v2 = *a2, v3 = *a3;
                    *a0 = __fadd_rn(v0, __fmul_rn(x, __uint_as_float(e0.y)));
                    if (n > 1u) *a1 = __fadd_rn(v1, __fmul_rn(x, __uint_as_float(e1.y)));
                    if (n > 2u) *a2 = __fadd_rn(v2, __fmul_rn(x, __uint_as_float(e2.y)));
                    if (n > 3u) *a3 = __fadd_rn(v3, __fmul_rn(x, __uint_as_float(e3.y)));
                }
            }
            __syncwarp();
        }
        cm_cp_async_wait<0>();
        if (have && (h.has_bias & 1u) && R > 0) {  // bias row last (inference.hpp:806-811)
            const uint2 br = __ldg(ext_g + (R - 1u));
            for (uint32_t e = br.x; e < br.y; ++e) {
                const uint2 en = __ldg(ent_g + e);
                float* a = my_acc + en.x * 32u;
                *a = __fadd_rn(*a, __fmul_rn(L.bias, __uint_as_float(en.y)));
            }
        }
        if (have) {
            float* dst = cand + static_cast<uint64_t>(q) * cand_stride_q + pos;
            for (uint32_t col = 0; col < n_cols; ++col) dst[col] = my_acc[col * 32];
        }
        s0 = s1;
    }
}
