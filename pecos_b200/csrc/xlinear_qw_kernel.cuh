// xl_query_warp_scores_kernel: one WARP walks one query through ALL the chunks of its beam at once.
//
// Included by xlinear_engine.cu (inside its anonymous namespace, after the small device helpers).
//
// Why: scoring a (query, chunk) pair touches only a few hundred {col, val} entries, and the per-pair bookkeeping of the
// warp-per-chunk kernel (flush set-up, scans, conflict rounds inside a 60..90 column block) costs more than the
// arithmetic.  Here the matched entries of the WHOLE beam are processed in feature-major order:
//
//   * entries of one feature hit different chunks => different output columns => no ordering constraint between them;
//   * 32 consecutive staged entries spread over (beam x chunk-width) ~ 600..1800 targets, so two of them rarely collide
//     (collisions are still resolved exactly, in staged order, through __match_any_sync);
//   * every phase is a flat, lane-parallel loop with its independent loads issued back to back.
//
// Arithmetic order per output column is unchanged: matched features ascending, un-fused multiply and add, bias row last
// (pecos/core/xmc/inference.hpp:788-811).  Eligibility (checked on the host, otherwise the warp-per-chunk kernel runs):
// sparse queries, feature maps present, beam <= kQwSlots, candidate row <= kQwNCap floats, query nnz <= kQwQCap.
#pragma once

constexpr int kQwWarps = 4;        // queries per CTA
constexpr int kQwSlots = 32;       // beam slots (one lane per slot in the prologue)
constexpr int kQwPairs = 256;      // (feature, chunk) probes per tile
constexpr int kQwECap = 512;       // staged entries per apply pass
constexpr uint32_t kQwNCap = 2048; // candidate row capacity (floats)
constexpr uint32_t kQwQCap = 512;  // query non-zeros staged per warp

struct QwSlot {
    const uint2* fm;       // feature-map cells of the chunk
    const uint32_t* rp;    // row_ptr (u32, relative to ent)
    const uint2* ent;      // entries of the chunk
    uint32_t base;         // first candidate position of this slot
    uint32_t n_rows;       // R
    uint32_t flags;        // bit0 has_bias, bit1 absent
    uint32_t n_cols;
};

__host__ __device__ inline size_t qw_warp_bytes(uint32_t q_cap, uint32_t n_cap) {
    return static_cast<size_t>(q_cap) * 8                      // qidx + qval
           + sizeof(QwSlot) * kQwSlots                          // slots
           + static_cast<size_t>(kQwPairs) * (4 + 4 + 4)        // ms, mj, mx
           + static_cast<size_t>(kQwPairs + 4) * 4              // off
           + static_cast<size_t>(kQwECap) * 8                   // stage
           + static_cast<size_t>(n_cap) * 4;                    // out
}

template <bool STATS>
__global__ void __launch_bounds__(kQwWarps * 32)
xl_query_warp_scores_kernel(const LayerDev L, const QueryDev X, const uint32_t* __restrict__ beam_id,
                            const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, float* __restrict__ cand,
                            const uint64_t cand_stride_q, unsigned long long* stats, const uint32_t q_cap,
                            const uint32_t n_cap, const uint32_t rows) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kQwWarps + warp;
    if (q >= rows) return;
    const size_t slice = (qw_warp_bytes(q_cap, n_cap) + 15) & ~static_cast<size_t>(15);
    unsigned char* base_ptr = smem_raw + warp * slice;
    uint32_t* qidx = reinterpret_cast<uint32_t*>(base_ptr);
    float* qval = reinterpret_cast<float*>(qidx + q_cap);
    QwSlot* slots = reinterpret_cast<QwSlot*>(qval + q_cap);
    uint32_t* ms = reinterpret_cast<uint32_t*>(slots + kQwSlots);   // matched row slot -> first entry offset
    uint32_t* mj = ms + kQwPairs;                                   // beam slot of the match
    float* mx = reinterpret_cast<float*>(mj + kQwPairs);            // multiplier
    uint32_t* off = reinterpret_cast<uint32_t*>(mx + kQwPairs);     // entry prefix
    uint2* stage = reinterpret_cast<uint2*>(off + kQwPairs + 4);    // {target position, bits of x*w}
    float* out = reinterpret_cast<float*>(stage + kQwECap);         // candidate row of this query

    // ---- prologue: all global loads first, shared-memory stores afterwards (issue is in order)
    const uint32_t cnt = min(beam_cnt[q], static_cast<uint32_t>(kQwSlots));
    const uint64_t qb = X.row_ptr[q] - X.nnz_base;
    const int qn = static_cast<int>(X.row_ptr[q + 1] - X.nnz_base - qb);
    const uint32_t my_p = (static_cast<uint32_t>(lane) < cnt) ? beam_id[static_cast<uint64_t>(q) * beam_stride + lane] : 0u;
    ChunkHeader my_h;
    my_h.n_cols = 0; my_h.nnz_rows = 0; my_h.has_bias = 0; my_h.meta_off = 0; my_h.ent_off = 0; my_h.col_begin = 0;
    if (static_cast<uint32_t>(lane) < cnt) my_h = L.chunks[my_p];
    for (int i0 = lane; i0 < qn; i0 += 128) {
        uint32_t fi[4];
        float fv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; fi[u] = (i < qn) ? X.col_idx[qb + i] : 0u; fv[u] = (i < qn) ? X.val[qb + i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u; if (i < qn) { qidx[i] = fi[u]; qval[i] = fv[u]; } }
    }
    // slot table + exclusive prefix of the chunk widths (candidate positions are global over the beam)
    const uint32_t incl = warp_incl_scan(my_h.n_cols, lane);
    const uint32_t n_total = __shfl_sync(kFull, incl, 31);
    if (static_cast<uint32_t>(lane) < cnt) {
        QwSlot s;
        const uint32_t R4 = (my_h.nnz_rows + 3u) & ~3u;
        s.fm = L.featmap + static_cast<uint64_t>(my_p) * L.fm_words;
        s.rp = L.meta + my_h.meta_off + R4;
        s.ent = L.entries + my_h.ent_off;
        s.base = incl - my_h.n_cols;
        s.n_rows = my_h.nnz_rows;
        s.flags = my_h.has_bias;
        s.n_cols = my_h.n_cols;
        slots[lane] = s;
    }
    for (uint32_t i = lane; i < n_total; i += 32) out[i] = 0.0f;
    __syncwarp();

    unsigned long long st_match = 0, st_ent = 0;

    // Applies the m matches collected in (ms, mj, mx), which are in feature-major order.
    auto flush = [&](int m) {
        if (m == 0) return;
        __syncwarp();
        // row extents (8 matches per lane, all loads in flight), prefix over the tile
        constexpr int PER = kQwPairs / 32;
        uint32_t a[PER], c[PER];
        uint32_t local = 0;
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = lane * PER + u;
            a[u] = 0; c[u] = 0;
            if (i < m) {
                const uint32_t* rp = slots[mj[i]].rp;
                const uint32_t s = ms[i];
                const uint32_t lo = __ldg(rp + s), hi = __ldg(rp + s + 1);
                a[u] = lo; c[u] = hi - lo;
            }
            local += c[u];
        }
        const uint32_t inc = warp_incl_scan(local, lane);
        uint32_t run = inc - local;
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        __syncwarp();
#pragma unroll
        for (int u = 0; u < PER; ++u) {
            const int i = lane * PER + u;
            if (i < m) { ms[i] = a[u]; off[i] = run; run += c[u]; }
        }
        if (lane == 0) off[m] = total;
        __syncwarp();
        st_ent += total;
        st_match += m;

        int i0 = 0;
        uint32_t part = 0;
        while (i0 < m) {
            const uint32_t row_begin = off[i0];
            const uint32_t tbase = row_begin + part;
            int i_next;
            uint32_t part_next = 0, ne;
            if (part == 0 && off[i0 + 1] - row_begin <= static_cast<uint32_t>(kQwECap)) {
                const int i1 = last_le_u32(off, m + 1, tbase + kQwECap);
                for (int i = i0 + lane; i < i1; i += 32) {
                    const QwSlot& sl = slots[mj[i]];
                    const uint2* src = sl.ent + ms[i];
                    const uint32_t n = off[i + 1] - off[i];
                    const uint32_t o = off[i] - tbase;
                    const uint32_t cb = sl.base;
                    const float x = mx[i];
                    uint32_t j = 0;
                    for (; j + 4 <= n; j += 4) {
                        const uint2 e0 = __ldg(src + j), e1 = __ldg(src + j + 1), e2 = __ldg(src + j + 2), e3 = __ldg(src + j + 3);
                        stage[o + j] = make_uint2(cb + e0.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e0.y))));
                        stage[o + j + 1] = make_uint2(cb + e1.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e1.y))));
                        stage[o + j + 2] = make_uint2(cb + e2.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e2.y))));
                        stage[o + j + 3] = make_uint2(cb + e3.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e3.y))));
                    }
                    if (j < n) {
                        const uint2 e0 = __ldg(src + j);
                        const uint2 e1 = (j + 1 < n) ? __ldg(src + j + 1) : e0;
                        const uint2 e2 = (j + 2 < n) ? __ldg(src + j + 2) : e0;
                        stage[o + j] = make_uint2(cb + e0.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e0.y))));
                        if (j + 1 < n) stage[o + j + 1] = make_uint2(cb + e1.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e1.y))));
                        if (j + 2 < n) stage[o + j + 2] = make_uint2(cb + e2.x, __float_as_uint(__fmul_rn(x, __uint_as_float(e2.y))));
                    }
                }
                ne = off[i1] - tbase;
                i_next = i1;
            } else {  // one row longer than the staging area: split it (its entries are applied in stored order)
                const QwSlot& sl = slots[mj[i0]];
                const uint32_t n_left = off[i0 + 1] - tbase;
                ne = min(static_cast<uint32_t>(kQwECap), n_left);
                const uint2* src = sl.ent + ms[i0] + part;
                const float x = mx[i0];
                for (uint32_t g = lane; g < ne; g += 32) {
                    const uint2 en = __ldg(src + g);
                    stage[g] = make_uint2(sl.base + en.x, __float_as_uint(__fmul_rn(x, __uint_as_float(en.y))));
                }
                if (ne == n_left) { i_next = i0 + 1; } else { i_next = i0; part_next = part + ne; }
            }
            __syncwarp();
            for (uint32_t g0 = 0; g0 < ne; g0 += 32) {
                const uint32_t g = g0 + lane;
                const bool valid = g < ne;
                const uint2 s = valid ? stage[g] : make_uint2(0xFFFFFFFFu - lane, 0u);
                const unsigned peers = __match_any_sync(kFull, s.x);
                const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
                const uint32_t rounds = __reduce_max_sync(kFull, valid ? rank : 0u);
                for (uint32_t r = 0; r <= rounds; ++r) {
                    if (valid && rank == r) out[s.x] = __fadd_rn(out[s.x], __uint_as_float(s.y));
                    __syncwarp();
                }
            }
            __syncwarp();
            i0 = i_next;
            part = part_next;
        }
    };

    // ---- feature tiles: T features x cnt chunks probes per tile, at most kQwPairs
    if (cnt > 0 && qn > 0) {
        const int T = max(1, kQwPairs / static_cast<int>(cnt));
        for (int t0 = 0; t0 < qn; t0 += T) {
            const int tn = min(T, qn - t0);
            const int n_pairs = tn * static_cast<int>(cnt);
            int m = 0;
            for (int pb = 0; pb < n_pairs; pb += 128) {
                uint2 cell[4];
                uint32_t feat[4];
                int tt[4], jj[4];
                bool live[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int pi = pb + 32 * u + lane;
                    live[u] = false; feat[u] = 0; tt[u] = 0; jj[u] = 0; cell[u] = make_uint2(0u, 0u);
                    if (pi < n_pairs) {
                        tt[u] = t0 + pi / static_cast<int>(cnt);
                        jj[u] = pi - (tt[u] - t0) * static_cast<int>(cnt);
                        const uint32_t f = qidx[tt[u]];
                        const bool dup = (tt[u] > 0) && (qidx[tt[u] - 1] == f);   // only the first of repeated indices counts
                        const QwSlot& sl = slots[jj[u]];
                        if (!dup && f < L.w_rows && sl.n_rows > 0 && !(sl.flags & kChunkAbsent)) {
                            live[u] = true; feat[u] = f; cell[u] = __ldg(sl.fm + (f >> 5));
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (pb + 32 * u >= n_pairs) break;
                    const uint32_t bit = feat[u] & 31u;
                    const bool hit = live[u] && ((cell[u].x >> bit) & 1u);
                    const unsigned mask = __ballot_sync(kFull, hit);
                    if (mask == 0u) continue;
                    if (hit) {
                        const uint32_t pos = m + __popc(mask & ((1u << lane) - 1u));
                        ms[pos] = cell[u].y + __popc(cell[u].x & ((1u << bit) - 1u));
                        mj[pos] = static_cast<uint32_t>(jj[u]);
                        mx[pos] = qval[tt[u]];
                    }
                    m += __popc(mask);
                }
            }
            flush(m);
        }
    }
    // ---- bias rows last: one pseudo-feature for every chunk with an explicit bias row
    {
        const bool has = static_cast<uint32_t>(lane) < cnt && (slots[lane].flags & 1u) && !(slots[lane].flags & kChunkAbsent);
        const unsigned mask = __ballot_sync(kFull, has);
        if (mask) {
            if (has) {
                const uint32_t pos = __popc(mask & ((1u << lane) - 1u));
                ms[pos] = slots[lane].n_rows - 1u;
                mj[pos] = static_cast<uint32_t>(lane);
                mx[pos] = L.bias;
            }
            flush(__popc(mask));
        }
    }
    __syncwarp();
    float* dst = cand + static_cast<uint64_t>(q) * cand_stride_q;
    for (uint32_t i = lane; i < n_total; i += 32) dst[i] = out[i];
    if (STATS) {
        // per-query counters of SURVEY 8(d); absent chunks (index sharding) are not scored and not counted
        unsigned long long rows_sum = 0, cols_sum = 0, chunks = 0;
        if (static_cast<uint32_t>(lane) < cnt && !(slots[lane].flags & kChunkAbsent)) { rows_sum = slots[lane].n_rows; cols_sum = slots[lane].n_cols; chunks = 1; }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            rows_sum += __shfl_xor_sync(kFull, rows_sum, d);
            cols_sum += __shfl_xor_sync(kFull, cols_sum, d);
            chunks += __shfl_xor_sync(kFull, chunks, d);
        }
        if (lane == 0) {
            atomicAdd(&stats[0], chunks);
            atomicAdd(&stats[1], rows_sum);
            atomicAdd(&stats[2], st_match);
            atomicAdd(&stats[3], st_ent);
            atomicAdd(&stats[4], cols_sum);
            if (cnt > 0) atomicAdd(&stats[5], static_cast<unsigned long long>(qn));
        }
    }
}
