// Pair-sorted feature-map scoring: ONE WARP PER (query, chunk) PAIR, pairs taken from the chunk-sorted pair list that the
// chunk-major bucketing (xl_cm_count / scan / scatter) produces.
//
// Included by xlinear_engine.cu (inside its anonymous namespace, after xl_flush_impl and the bucketing kernels).
//
// Why: the query-major kernel (one CTA per query, one warp per beam slot) sends every probe of a wide layer to DRAM -- on the
// 3M-label leaf layer 256 M scattered 32-byte sector reads per batch, the kernel sits at 27 % of the HBM bandwidth with the
// long-scoreboard stall dominating (profiles/r01_h).  The arithmetic is fine; the ORDER of the work is not: the ~61
// pairs that visit one leaf chunk are spread over the whole grid.  Sorted by chunk, they run back to back on neighbouring
// warps, the first pair of a chunk prefetches the chunk's feature map into L2 with coalesced 128-byte requests (sequential
// DRAM traffic, each map read once), and every later probe / extent / entry access of that chunk is an L2 hit.
//
// The per-pair algorithm is the feature-map path of xl_chunk_scores_kernel, unchanged: 128-feature blocks probe the map,
// matches are compacted in feature order, xl_flush_impl applies the matched rows' entries in the reference's order (bias row
// last) -- same bits (tests/test_chunk_major_gpu.py::test_pair_sorted_kernel_*).
#pragma once

constexpr int kPwWarps = 8;

__host__ __device__ inline size_t pw_warp_bytes(uint32_t q_cap) {
    return static_cast<size_t>(q_cap) * 8 + sizeof(WarpScratch<kMCapLookup>);
}

__global__ void __launch_bounds__(kPwWarps * 32, 4)
xl_pair_scores_kernel(const LayerDev L, const QueryDev X, const CmWork w, const uint32_t* __restrict__ pair_chunk,
                      float* __restrict__ cand, const uint64_t cand_stride_q, const uint32_t q_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned char* mine = smem_raw + static_cast<size_t>(warp) * ((pw_warp_bytes(q_cap) + 15) & ~static_cast<size_t>(15));
    uint32_t* q_idx_s = reinterpret_cast<uint32_t*>(mine);
    float* q_val_s = reinterpret_cast<float*>(mine + static_cast<size_t>(q_cap) * 4);
    WarpScratch<kMCapLookup>& ws = *reinterpret_cast<WarpScratch<kMCapLookup>*>(mine + static_cast<size_t>(q_cap) * 8);

    const uint32_t n_pairs = w.bucket_ptr[L.n_chunks];
    const uint32_t g = blockIdx.x * kPwWarps + warp;
    if (g >= n_pairs) return;
    const uint32_t p = pair_chunk[g];
    const uint32_t q = w.pair_q[g];
    const uint32_t pos = w.pair_pos[g];
    const ChunkHeader h = L.chunks[p];
    const uint2* fm = L.featmap + static_cast<uint64_t>(p) * L.fm_words;
    if (g == w.bucket_ptr[p]) {
        // first pair of the chunk: pull the chunk's feature map into L2 with coalesced requests (one 128-byte line per lane and
        // step) -- the probes of the chunk's other pairs, running on the neighbouring warps, then hit L2
        const char* base = reinterpret_cast<const char*>(fm);
        const uint32_t bytes = L.fm_words * 8u;
        for (uint32_t off = lane * 128u; off < bytes; off += 32u * 128u)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(base + off));
    }
    const uint64_t qb = X.row_ptr[q] - X.nnz_base;
    const int qn = static_cast<int>(X.row_ptr[q + 1] - X.nnz_base - qb);
    const uint32_t* qidx = X.col_idx + qb;
    const float* qval = X.val + qb;
    if (qn <= static_cast<int>(q_cap)) {  // stage the query row (longer rows are read through L1 / L2)
        for (int i = lane; i < qn; i += 32) { q_idx_s[i] = qidx[i]; q_val_s[i] = qval[i]; }
        qidx = q_idx_s;
        qval = q_val_s;
    }
    __syncwarp();
    const bool chunk_bias = (h.has_bias & 1u) != 0u;
    const uint32_t R = h.nnz_rows;
    const uint2* ext = reinterpret_cast<const uint2*>(L.rowext + h.meta_off);
    const uint2* ent = L.entries + h.ent_off;
    float* blk = cand + static_cast<uint64_t>(q) * cand_stride_q + pos;
    const bool in_smem = h.n_cols <= static_cast<uint32_t>(kCSmem);
    float* out = in_smem ? ws.out : blk;
    for (uint32_t c = lane; c < h.n_cols; c += 32) out[c] = 0.0f;
    __syncwarp();

    constexpr int MCAP = kMCapLookup;
    int m = 0;
    unsigned long long e_total = 0;
    int tb0 = 0;
    do {
        if (R > 0 && tb0 < qn) {
            uint2 cell[4];
            uint32_t feat[4];
            bool live[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int t = tb0 + 32 * u + lane;
                live[u] = false;
                feat[u] = 0;
                cell[u] = make_uint2(0u, 0u);
                if (t < qn) {
                    const uint32_t f = qidx[t];
                    const bool dup = (t > 0) && (qidx[t - 1] == f);  // a repeated column index only counts once
                    if (!dup && f < L.w_rows) { live[u] = true; feat[u] = f; cell[u] = __ldg(fm + (f >> 5)); }
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (tb0 + 32 * u >= qn) break;
                const int t = tb0 + 32 * u + lane;
                const uint32_t bit = feat[u] & 31u;
                const bool hit = live[u] && ((cell[u].x >> bit) & 1u);
                const unsigned mask = __ballot_sync(kFull, hit);
                if (mask == 0u) continue;
                if (hit) {
                    const uint32_t at = m + __popc(mask & ((1u << lane) - 1u));
                    ws.ms[at] = cell[u].y + __popc(cell[u].x & ((1u << bit) - 1u));
                    ws.mx[at] = qval[t];
                }
                m += __popc(mask);
            }
        }
        tb0 += 128;
        const bool last_block = tb0 >= qn;
        if (last_block && chunk_bias) {
            __syncwarp();
            if (lane == 0) { ws.ms[m] = R - 1u; ws.mx[m] = L.bias; }
            ++m;
        }
        if (last_block || m > MCAP - 130) {  // room for the next block of <= 128 matches and the bias row
            xl_flush_impl(ws, m, ext, ent, out, lane, e_total);
            m = 0;
        }
    } while (tb0 < qn);
    __syncwarp();
    if (in_smem) {
        for (uint32_t c = lane; c < h.n_cols; c += 32) blk[c] = ws.out[c];
    }
}
