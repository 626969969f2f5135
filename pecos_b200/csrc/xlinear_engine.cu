// XR-Linear beam search on B200 (sm_100a): kernels + engine.  See xlinear_engine.h for the reference map.
//
// Per tree layer two kernels run over a tile of queries:
//
//   xl_chunk_scores_kernel   one CTA per query, one warp per beam slot (= one weight chunk).  The warp streams the
//                            chunk's sorted row-index list from HBM with 128-bit loads, intersects it with the query's
//                            sorted feature list held in shared memory, gathers the matched rows' {col,val} entries and
//                            accumulates them into the chunk's dense output block IN THE REFERENCE'S ORDER:
//                            ascending feature index, separate round-to-nearest multiply and add, bias row last
//                            (inference.hpp:788-811; dense queries: bias first, inference.hpp:823-837).
//                            HBM-bound; algorithmic bytes per (query, chunk) = 32 + 4R + 16m + 8e + 4c  (SURVEY 8d).
//
//   xl_topk_kernel           one CTA per query: post-processor transform in double precision (inference.hpp:208-238),
//                            combine with the parent's path score, then exact top-k on the composite key
//                            (score desc, position-in-prolongated-row asc) == sorted_csr's comparator
//                            (inference.hpp:1265-1273).  Bitonic sort of 64-bit keys in shared memory.
//
// No FMA contraction anywhere on the score path: products use __fmul_rn, sums use __fadd_rn (the reference build has
// no -march flag, so its XR-Linear loops are scalar mulss/addss).
#include "xlinear_engine.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <exception>
#include <thread>

namespace pb200 {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kWarpsMax = 10;    // warps per CTA of the chunk kernel
constexpr int kMCap = 256;       // match list capacity per warp
constexpr int kMFlush = 128;     // flush the match list once it holds this many rows (kMCap - 128 new per pass)
constexpr int kCSmem = 128;      // chunk widths up to this accumulate in shared memory, wider ones in the HBM block
constexpr int kQCap = 1024;      // query non-zeros staged in shared memory (longer queries are read through L1/L2)
constexpr int kSortCap = 2048;   // keys sorted in shared memory per pass of the top-k kernel
constexpr int kTopkThreads = 256;

constexpr int kMCapLookup = 256; // the lookup kernel collects a block of <= 128 matches (+ bias row) between flush checks

template <int MCAP>
struct __align__(16) WarpScratch {
    uint32_t ms[MCAP];         // chunk-row index of each match; becomes the row's first entry offset during flush
    float mx[MCAP];            // multiplier of the row: query value, or the bias
    uint32_t off[MCAP + 4];    // exclusive prefix of the matched rows' entry counts
    float out[kCSmem];         // dense output block of the chunk
};

__device__ __forceinline__ uint4 ld_stream_u4(const uint32_t* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ int lower_bound_u32(const uint32_t* a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// largest i in [0, n) with a[i] <= key, given a[0] <= key
__device__ __forceinline__ int last_le_u32(const uint32_t* a, int n, uint32_t key) {
    int lo = 0, hi = n;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (a[mid] <= key) lo = mid; else hi = mid;
    }
    return lo;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(kFull, v, d);
        if (lane >= d) v += t;
    }
    return v;
}

// Rows of the 32 consecutive entries [G, G + 32) of the concatenated matched rows (off[] = their prefix sums, off[m] =
// total, every row non-empty).  row_first = row holding entry G (warp-uniform; updated to the row holding entry G + 32).
// Lane L looks at candidate row row_first + 1 + L: if it starts inside the group it sets the bit of its first entry; a
// lane's row is then row_first + (number of row starts at or before the lane).  ~12 instructions instead of a
// log2(m)-step binary search per entry.
__device__ __forceinline__ int xl_rows_of_group(const uint32_t* off, int m, uint32_t G, int& row_first, int lane) {
    const int r = row_first + 1 + lane;
    uint32_t bit = 0;
    if (r < m) {
        const uint32_t p = off[r] - G;
        if (p < 32u) bit = 1u << p;
    }
    const uint32_t starts = __reduce_or_sync(kFull, bit);
    const int my_row = row_first + __popc(starts & (0xFFFFFFFFu >> (31 - lane)));
    const int r31 = __shfl_sync(kFull, my_row, 31);
    row_first = (off[r31 + 1] == G + 32u) ? r31 + 1 : r31;
    return my_row;
}

// Apply the matched rows collected in ws (in ascending feature order) to the output block.
//
// Entry-parallel: the m matched rows hold `total` entries; lane L of group g owns entry 32g + L of their concatenation
// (its row found by a binary search in the rows' prefix sums), so ragged rows -- one entry for a rare feature, every column
// for the bias row -- cost the same per entry, consecutive lanes read consecutive entries, and nothing is staged in
// shared memory.  Entries of one 32-group that hit the same column (they come from different rows, or from a row that
// repeats a column) are added in lane order = concatenation order = ascending feature order; __match_any_sync finds the
// collisions, so the common collision-free group costs a single round.  Groups are applied in order.
template <int MCAP>
__device__ __forceinline__ void xl_flush_impl(WarpScratch<MCAP>& ws, int m, const uint2* __restrict__ ext,
                                              const uint2* __restrict__ ent, float* out, int lane,
                                              unsigned long long& e_total) {
    if (m == 0) return;
    __syncwarp();
    constexpr int PER = MCAP / 32;
    uint32_t c[PER];
    uint32_t local = 0;
    bool empty_row = false;
#pragma unroll
    for (int u = 0; u < PER; ++u) {  // PER independent 8-byte loads per lane; ms[i]: chunk row -> its first entry
        const int i = lane * PER + u;
        c[u] = 0;
        if (i < m) {
            const uint2 lh = __ldg(ext + ws.ms[i]);
            ws.ms[i] = lh.x;
            c[u] = lh.y - lh.x;
            empty_row |= (c[u] == 0u);
        }
        local += c[u];
    }
    const bool search = __any_sync(kFull, empty_row);  // never for chunks built from a CSC matrix (rows have >= 1 entry)
    const uint32_t incl = warp_incl_scan(local, lane);
    uint32_t run = incl - local;
    const uint32_t total = __shfl_sync(kFull, incl, 31);
#pragma unroll
    for (int u = 0; u < PER; ++u) {
        const int i = lane * PER + u;
        if (i < m) { ws.off[i] = run; run += c[u]; }
    }
    if (lane == 0) ws.off[m] = total;
    __syncwarp();
    e_total += total;

    int row_first = 0;
    for (uint32_t g0 = 0; g0 < total; g0 += 128u) {
        uint2 e[4];
        float x[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {  // four entry loads in flight per lane
            const uint32_t G = g0 + 32u * u;
            const uint32_t g = G + lane;
            e[u] = make_uint2(0xFFFFFFFFu - lane, 0u);  // idle lanes: distinct pseudo-columns, never applied
            x[u] = 0.0f;
            if (G < total) {
                int i = 0;
                if (!search) i = xl_rows_of_group(ws.off, m, G, row_first, lane);
                if (g < total) {
                    if (search) i = last_le_u32(ws.off, m, g);  // off[i] <= g < off[i + 1]
                    e[u] = __ldg(ent + ws.ms[i] + (g - ws.off[i]));
                    x[u] = ws.mx[i];
                }
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (g0 + 32u * u >= total) break;
            const bool valid = (g0 + 32u * u + lane) < total;
            const float v = __fmul_rn(x[u], __uint_as_float(e[u].y));
            const unsigned peers = __match_any_sync(kFull, e[u].x);
            const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
            const uint32_t rounds = __reduce_max_sync(kFull, valid ? rank : 0u);
            for (uint32_t r = 0; r <= rounds; ++r) {
                if (valid && rank == r) out[e[u].x] = __fadd_rn(out[e[u].x], v);
                __syncwarp();
            }
        }
    }
}

// out-of-line copy for the kernels that flush from several places
template <int MCAP>
__device__ __noinline__ void xl_flush(WarpScratch<MCAP>& ws, int m, const uint2* __restrict__ ext,
                                      const uint2* __restrict__ ent, float* out, int lane, unsigned long long& e_total) {
    xl_flush_impl(ws, m, ext, ent, out, lane, e_total);
}

// DENSE: row-major dense queries.  LOOKUP: sparse queries probe the chunk's feature map (one 8-byte cell per query
// feature) instead of streaming the chunk's row list -- same matches in the same order, far fewer bytes/instructions.
template <bool DENSE, bool STATS, bool LOOKUP>
__global__ void __launch_bounds__(kWarpsMax * 32, LOOKUP ? 4 : 1)  // lookup variant: <= 51 registers => 4 x 10-warp CTAs per SM
xl_chunk_scores_kernel(const LayerDev L, const QueryDev X, const uint32_t* __restrict__ beam_id,
                       const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, float* __restrict__ cand,
                       const uint64_t cand_stride_q, const uint32_t c_stride, unsigned long long* stats,
                       const uint32_t q_cap, const uint32_t sb_cap, const uint32_t hdr_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* q_idx_s = reinterpret_cast<uint32_t*>(smem_raw);
    float* q_val_s = reinterpret_cast<float*>(smem_raw + q_cap * 4);
    uint32_t* slot_base = reinterpret_cast<uint32_t*>(smem_raw + q_cap * 8);  // [cnt + 1] first candidate of each beam slot
    ChunkHeader* hdr_s = reinterpret_cast<ChunkHeader*>(smem_raw + q_cap * 8 + sb_cap * 4);  // [hdr_cap] beam chunk headers
    constexpr int MCAP = LOOKUP ? kMCapLookup : kMCap;
    WarpScratch<MCAP>* scratch =
        reinterpret_cast<WarpScratch<MCAP>*>(smem_raw + q_cap * 8 + sb_cap * 4 + static_cast<size_t>(hdr_cap) * sizeof(ChunkHeader));

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const uint32_t q = blockIdx.x;

    // ---- prologue.  All global loads are issued before the first shared-memory store (issue is in order: a store that
    // waits for its operand would otherwise serialise the independent load chains below).
    const uint32_t cnt = beam_cnt[q];
    uint64_t qb = 0, qe = 0;
    if (!DENSE) { qb = X.row_ptr[q] - X.nnz_base; qe = X.row_ptr[q + 1] - X.nnz_base; }
    const uint32_t my_p = (threadIdx.x < cnt) ? beam_id[static_cast<uint64_t>(q) * beam_stride + threadIdx.x] : 0u;
    const uint32_t* qidx = nullptr;
    const float* qval = nullptr;
    int qn = 0;
    uint32_t first_idx = 0;
    float first_val = 0.0f;
    bool staged = false;
    if (!DENSE) {
        qn = static_cast<int>(qe - qb);
        qidx = X.col_idx + qb;
        qval = X.val + qb;
        staged = qn <= static_cast<int>(q_cap);
        if (staged && static_cast<int>(threadIdx.x) < qn) { first_idx = qidx[threadIdx.x]; first_val = qval[threadIdx.x]; }
    } else {
        qval = X.val + static_cast<uint64_t>(q) * X.cols;
    }
    ChunkHeader my_h;
    my_h.n_cols = 0;
    if (threadIdx.x < cnt) my_h = L.chunks[my_p];
    if (staged) {
        if (static_cast<int>(threadIdx.x) < qn) { q_idx_s[threadIdx.x] = first_idx; q_val_s[threadIdx.x] = first_val; }
        for (int i = threadIdx.x + blockDim.x; i < qn; i += blockDim.x) { q_idx_s[i] = qidx[i]; q_val_s[i] = qval[i]; }
        qidx = q_idx_s;
        qval = q_val_s;
    }
    // candidates of a query are stored compactly in prolongation order: slot j starts at the sum of the widths before it
    if (threadIdx.x < cnt) {
        slot_base[threadIdx.x + 1] = my_h.n_cols;
        my_h.col_begin = my_p;  // this kernel never needs col_begin: the cached copy carries the chunk id instead
        if (threadIdx.x < hdr_cap) hdr_s[threadIdx.x] = my_h;
    }
    for (uint32_t j = threadIdx.x + blockDim.x; j < cnt; j += blockDim.x) {
        const uint32_t pj = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
        ChunkHeader hh = L.chunks[pj];
        slot_base[j + 1] = hh.n_cols;
        hh.col_begin = pj;
        if (j < hdr_cap) hdr_s[j] = hh;
    }
    __syncthreads();
    if (warp == 0) {  // widths -> first candidate position of every slot (warp scan, 32 slots per step)
        uint32_t run = 0;
        for (uint32_t j0 = 0; j0 < cnt; j0 += 32) {
            const uint32_t j = j0 + lane;
            const uint32_t w = (j < cnt) ? slot_base[j + 1] : 0u;
            const uint32_t incl = warp_incl_scan(w, lane);
            if (j < cnt) slot_base[j + 1] = run + incl;
            run += __shfl_sync(kFull, incl, 31);
        }
        if (lane == 0) slot_base[0] = 0;
    }
    __syncthreads();
    WarpScratch<MCAP>& ws = scratch[warp];
    unsigned long long st_chunks = 0, st_rows = 0, st_match = 0, st_ent = 0, st_cols = 0;

    for (uint32_t j = warp; j < cnt; j += nwarps) {
        uint32_t p;
        ChunkHeader h;
        if (j < hdr_cap) {
            h = hdr_s[j];
            p = h.col_begin;
        } else {
            p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
            h = L.chunks[p];
        }
        if (h.has_bias & kChunkAbsent) continue;  // leaf chunk owned by another GPU (index sharding): not scored here
        const bool chunk_bias = (h.has_bias & 1u) != 0u;
        const uint32_t R = h.nnz_rows;
        const uint32_t R4 = (R + 3u) & ~3u;
        const uint32_t* ridx = L.meta + h.meta_off;
        const uint2* ext = reinterpret_cast<const uint2*>(L.rowext + h.meta_off);  // {begin, end} of every chunk row
        const uint2* ent = L.entries + h.ent_off;
        float* blk = cand + static_cast<uint64_t>(q) * cand_stride_q + slot_base[j];
        const bool in_smem = h.n_cols <= static_cast<uint32_t>(kCSmem);
        float* out = in_smem ? ws.out : blk;
        for (uint32_t c = lane; c < h.n_cols; c += 32) out[c] = 0.0f;
        __syncwarp();

        int m = 0;
        unsigned long long e_total = 0, m_total = 0;

        if (DENSE) {
            // chunk_ops<drm, bin_search>: bias row first, then every chunk row (inference.hpp:823-837)
            const uint32_t r_lim = chunk_bias ? R - 1u : R;
            if (chunk_bias) {
                if (lane == 0) { ws.ms[0] = R - 1u; ws.mx[0] = L.bias; }
                m = 1;
                __syncwarp();
            }
            for (uint32_t base = 0; base < r_lim; base += 128u) {
                const uint32_t i = base + lane * 4u;
                const uint32_t n_here = (i < r_lim) ? min(4u, r_lim - i) : 0u;
                uint4 v = make_uint4(0, 0, 0, 0);
                if (n_here) v = ld_stream_u4(ridx + i);
                const uint32_t incl = warp_incl_scan(n_here, lane);
                const uint32_t tot = __shfl_sync(kFull, incl, 31);
                uint32_t pos = m + incl - n_here;
                if (n_here > 0) { ws.ms[pos] = i; ws.mx[pos] = qval[v.x]; }
                if (n_here > 1) { ws.ms[pos + 1] = i + 1; ws.mx[pos + 1] = qval[v.y]; }
                if (n_here > 2) { ws.ms[pos + 2] = i + 2; ws.mx[pos + 2] = qval[v.z]; }
                if (n_here > 3) { ws.ms[pos + 3] = i + 3; ws.mx[pos + 3] = qval[v.w]; }
                m += static_cast<int>(tot);
                if (m >= kMFlush) {
                    m_total += m;
                    xl_flush(ws, m, ext, ent, out, lane, e_total);
                    m = 0;
                }
            }
        } else {
            // chunk_ops<csr, bin_search>: matched rows in ascending feature order, bias row last (inference.hpp:788-811)
            if (LOOKUP) {
                // Blocks of 128 query features: four probe rounds issued back to back (4 independent cell loads in flight
                // per lane), matches compacted in feature order.  The match list takes a whole block (+ the bias row), so
                // the single flush site sits outside the probe registers' live range.
                const uint2* fm = L.featmap + static_cast<uint64_t>(p) * L.fm_words;
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
                                // a repeated column index only counts once: the reference's marching loop consumes the first
                                const bool dup = (t > 0) && (qidx[t - 1] == f);
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
                                const uint32_t pos = m + __popc(mask & ((1u << lane) - 1u));
                                ws.ms[pos] = cell[u].y + __popc(cell[u].x & ((1u << bit) - 1u));
                                ws.mx[pos] = qval[t];
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
                        m_total += m;
                        xl_flush_impl(ws, m, ext, ent, out, lane, e_total);
                        m = 0;
                    }
                } while (tb0 < qn);
            } else if (qn > 0 && R > 0) {
                const uint32_t qmin = qidx[0];
                const uint32_t qmax = qidx[qn - 1];
                const uint4 sentinel = make_uint4(0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu);
                uint32_t i = lane * 4u;
                uint4 v_next = (i < R4) ? ld_stream_u4(ridx + i) : sentinel;
                for (uint32_t base = 0; base < R4; base += 128u) {
                    const uint4 v = v_next;
                    i = base + lane * 4u;
                    const uint32_t i_n = i + 128u;
                    v_next = (i_n < R4) ? ld_stream_u4(ridx + i_n) : sentinel;
                    const uint32_t first = __shfl_sync(kFull, v.x, 0);
                    if (first > qmax) break;  // the remaining (sorted) chunk rows lie beyond the query's last feature
                    int t0 = -1, t1 = -1, t2 = -1, t3 = -1;
                    if (v.x >= qmin && v.x <= qmax) { int t = lower_bound_u32(qidx, qn, v.x); if (t < qn && qidx[t] == v.x) t0 = t; }
                    if (v.y >= qmin && v.y <= qmax) { int t = lower_bound_u32(qidx, qn, v.y); if (t < qn && qidx[t] == v.y) t1 = t; }
                    if (v.z >= qmin && v.z <= qmax) { int t = lower_bound_u32(qidx, qn, v.z); if (t < qn && qidx[t] == v.z) t2 = t; }
                    if (v.w >= qmin && v.w <= qmax) { int t = lower_bound_u32(qidx, qn, v.w); if (t < qn && qidx[t] == v.w) t3 = t; }
                    const uint32_t n_here = (t0 >= 0) + (t1 >= 0) + (t2 >= 0) + (t3 >= 0);
                    if (__ballot_sync(kFull, n_here > 0) == 0u) continue;
                    const uint32_t incl = warp_incl_scan(n_here, lane);
                    const uint32_t tot = __shfl_sync(kFull, incl, 31);
                    uint32_t pos = m + incl - n_here;
                    if (t0 >= 0) { ws.ms[pos] = i; ws.mx[pos] = qval[t0]; ++pos; }
                    if (t1 >= 0) { ws.ms[pos] = i + 1; ws.mx[pos] = qval[t1]; ++pos; }
                    if (t2 >= 0) { ws.ms[pos] = i + 2; ws.mx[pos] = qval[t2]; ++pos; }
                    if (t3 >= 0) { ws.ms[pos] = i + 3; ws.mx[pos] = qval[t3]; ++pos; }
                    m += static_cast<int>(tot);
                    if (m >= kMFlush) {
                        m_total += m;
                        xl_flush(ws, m, ext, ent, out, lane, e_total);
                        m = 0;
                    }
                }
            }
            if (!LOOKUP && chunk_bias) {
                __syncwarp();
                if (lane == 0) { ws.ms[m] = R - 1u; ws.mx[m] = L.bias; }
                ++m;
            }
        }
        if (!LOOKUP || DENSE) {
            m_total += m;
            xl_flush(ws, m, ext, ent, out, lane, e_total);
        }
        __syncwarp();
        if (in_smem) {
            for (uint32_t c = lane; c < h.n_cols; c += 32) blk[c] = ws.out[c];
        }
        __syncwarp();
        if (STATS) { st_chunks += 1; st_rows += R; st_match += m_total; st_ent += e_total; st_cols += h.n_cols; }
    }
    if (STATS) {
        if (lane == 0 && st_chunks) {
            atomicAdd(&stats[0], st_chunks);
            atomicAdd(&stats[1], st_rows);
            atomicAdd(&stats[2], st_match);
            atomicAdd(&stats[3], st_ent);
            atomicAdd(&stats[4], st_cols);
        }
        if (threadIdx.x == 0 && cnt > 0) atomicAdd(&stats[5], static_cast<unsigned long long>(DENSE ? X.cols : qn));
    }
}

// ------------------------------------------------------------------------------------------------------------------
// post-processor (inference.hpp:208-238).  The reference evaluates these through double precision libm calls and
// narrows to float; we do the same arithmetic with CUDA's double routines.  exp/log may differ from glibc in the last
// ulp of the DOUBLE result, which survives the narrowing to float only with probability ~2^-29 per element.
// Integer powers p <= 4 are formed by exact/singly-rounded multiplications (== correctly rounded pow).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double xl_hinge_pow(float v, int p) {
    const double zd = fmax(0.0, 1.0 - static_cast<double>(v));
    const float z = static_cast<float>(zd);
    const double x = static_cast<double>(z);
    switch (p) {
        case 0: return 1.0;
        case 1: return x;
        case 2: return x * x;
        case 3: return (x * x) * x;
        case 4: { const double x2 = x * x; return x2 * x2; }
        default: return pow(x, static_cast<double>(static_cast<size_t>(static_cast<long long>(p))));
    }
}

__device__ __forceinline__ float xl_transform(float v, int kind, int p) {
    switch (kind) {
        case PP_SIGMOID: {
            const float e = static_cast<float>(exp(static_cast<double>(-v)));  // expf(-v), correctly rounded
            return static_cast<float>(1.0 / (1.0 + static_cast<double>(e)));
        }
        case PP_LOG_SIGMOID: {
            const float e = static_cast<float>(exp(static_cast<double>(-v)));
            return static_cast<float>(-log(1.0 + static_cast<double>(e)));
        }
        case PP_LP_HINGE: return static_cast<float>(exp(-xl_hinge_pow(v, p)));
        case PP_LOG_LP_HINGE: return static_cast<float>(-xl_hinge_pow(v, p));
        default: return v;
    }
}

__device__ __forceinline__ float xl_combine(float x, float parent, int kind) {
    switch (kind) {
        case PP_SIGMOID:
        case PP_LP_HINGE: return __fmul_rn(x, parent);
        case PP_LOG_SIGMOID:
        case PP_LOG_LP_HINGE: return __fadd_rn(x, parent);
        default: return x;
    }
}

__device__ __forceinline__ unsigned long long xl_make_key(float v, uint32_t pos) {
    uint32_t u = __float_as_uint(v);
    if ((u & 0x7FFFFFFFu) == 0u) u = 0u;  // -0.0 and +0.0 compare equal in the reference comparator
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return (static_cast<unsigned long long>(u) << 32) | static_cast<unsigned long long>(0xFFFFFFFFu - pos);
}

#include "xlinear_qw_kernel.cuh"
#include "xlinear_cm_kernel.cuh"

// descending bitonic sort of n (power of two) keys; a may live in shared or global memory
__device__ void xl_bitonic_desc(unsigned long long* a, uint32_t n) {
    for (uint32_t k = 2; k <= n; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long x = a[i], y = a[ixj];
                    const bool desc_block = ((i & k) == 0);
                    if (desc_block ? (x < y) : (x > y)) { a[i] = y; a[ixj] = x; }
                }
            }
            __syncthreads();
        }
    }
}

__device__ __forceinline__ uint32_t next_pow2_u32(uint32_t v) {
    if (v <= 2) return 2;
    return 1u << (32 - __clz(v - 1));
}

__global__ void __launch_bounds__(kTopkThreads)
xl_topk_kernel(const LayerDev L, const int pp_kind, const int pp_p, const int combine, const uint32_t k,
               const uint32_t* __restrict__ beam_id, const float* __restrict__ beam_val,
               const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, const float* __restrict__ cand,
               const uint64_t cand_stride_q, const uint32_t c_stride, uint32_t* __restrict__ out_id,
               float* __restrict__ out_val, uint32_t* __restrict__ out_cnt, const uint32_t out_stride,
               unsigned long long* sortbuf, const uint64_t sortbuf_stride, const uint32_t b_prev,
               unsigned long long* stats, unsigned long long* __restrict__ out_key) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);
    uint32_t* s_base = reinterpret_cast<uint32_t*>(smem_raw + kSortCap * 8);  // [b_prev + 1]
    uint32_t* s_colbeg = s_base + (b_prev + 1);                               // [b_prev]
    float* s_pval = reinterpret_cast<float*>(s_colbeg + b_prev);              // [b_prev]

    const uint32_t q = blockIdx.x;
    const uint32_t cnt = beam_cnt[q];
    for (uint32_t j = threadIdx.x; j < cnt; j += blockDim.x) {
        const uint32_t p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
        const ChunkHeader h = L.chunks[p];
        s_base[j + 1] = h.n_cols;
        s_colbeg[j] = (h.has_bias & kChunkAbsent) ? 0xFFFFFFFFu : h.col_begin;  // absent: candidates exist elsewhere
        s_pval[j] = beam_val[static_cast<uint64_t>(q) * beam_stride + j];
    }
    __syncthreads();
    __shared__ uint32_t s_owned;
    if (threadIdx.x == 0) {
        uint32_t run = 0, owned = 0;
        s_base[0] = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            if (s_colbeg[j] != 0xFFFFFFFFu) owned += s_base[j + 1];
            run += s_base[j + 1];
            s_base[j + 1] = run;
        }
        s_owned = owned;
    }
    __syncthreads();
    const uint32_t n_valid = s_base[cnt];       // positions are global (all beam slots), keys only for owned slots
    const uint32_t kk = min(k, s_owned);
    if (threadIdx.x == 0) {
        out_cnt[q] = kk;
        if (stats) atomicAdd(&stats[6], static_cast<unsigned long long>(kk));
    }
    if (n_valid == 0) return;

    const float* cq = cand + static_cast<uint64_t>(q) * cand_stride_q;
    auto score_at = [&](uint32_t cpos, uint32_t& label) -> float {
        const uint32_t j = static_cast<uint32_t>(last_le_u32(s_base, static_cast<int>(cnt), cpos));
        const uint32_t off = cpos - s_base[j];
        float v = xl_transform(cq[cpos], pp_kind, pp_p);
        if (combine) v = xl_combine(v, s_pval[j], pp_kind);
        label = s_colbeg[j] + off;
        return v;
    };
    auto key_at = [&](uint32_t cpos) -> unsigned long long {
        const uint32_t j = static_cast<uint32_t>(last_le_u32(s_base, static_cast<int>(cnt), cpos));
        if (s_colbeg[j] == 0xFFFFFFFFu) return 0ull;  // scored on another GPU
        uint32_t lab;
        return xl_make_key(score_at(cpos, lab), cpos);
    };

    unsigned long long* sorted = keys;
    if (n_valid <= static_cast<uint32_t>(kSortCap)) {
        const uint32_t P = next_pow2_u32(n_valid);
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
            unsigned long long key = 0ull;
            if (i < n_valid) key = key_at(i);
            keys[i] = key;
        }
        __syncthreads();
        xl_bitonic_desc(keys, P);
    } else if (k <= static_cast<uint32_t>(kSortCap / 2)) {
        // streaming: keys[0, KP) hold the best so far, keys[KP, kSortCap) the next tile of candidates
        const uint32_t KP = next_pow2_u32(k);
        const uint32_t tile = kSortCap - KP;
        for (uint32_t i = threadIdx.x; i < KP; i += blockDim.x) keys[i] = 0ull;
        for (uint32_t t0 = 0; t0 < n_valid; t0 += tile) {
            for (uint32_t i = threadIdx.x; i < tile; i += blockDim.x) {
                const uint32_t cpos = t0 + i;
                unsigned long long key = 0ull;
                if (cpos < n_valid) key = key_at(cpos);
                keys[KP + i] = key;
            }
            __syncthreads();
            xl_bitonic_desc(keys, kSortCap);
        }
    } else {
        // wide beam AND large k: sort the whole candidate list in the HBM scratch
        sorted = sortbuf + static_cast<uint64_t>(q) * sortbuf_stride;
        const uint32_t P = next_pow2_u32(n_valid);
        for (uint32_t i = threadIdx.x; i < P; i += blockDim.x) {
            unsigned long long key = 0ull;
            if (i < n_valid) key = key_at(i);
            sorted[i] = key;
        }
        __syncthreads();
        xl_bitonic_desc(sorted, P);
    }

    for (uint32_t r = threadIdx.x; r < kk; r += blockDim.x) {
        const uint32_t cpos = 0xFFFFFFFFu - static_cast<uint32_t>(sorted[r] & 0xFFFFFFFFull);
        uint32_t label;
        const float v = score_at(cpos, label);  // recomputed so that the stored value keeps its exact bits (-0.0)
        if (L.label_of_col) label = L.label_of_col[label];
        out_id[static_cast<uint64_t>(q) * out_stride + r] = label;
        out_val[static_cast<uint64_t>(q) * out_stride + r] = v;
        if (out_key) out_key[static_cast<uint64_t>(q) * out_stride + r] = sorted[r];
    }
}

// Narrow-beam variant: one WARP per query (kSelWarps queries per CTA).  The keys of the <= kSelKeys candidates sit in the
// warp's shared-memory slice; the top-k is extracted by k rounds of warp arg-max on the 64-bit composite key (keys are
// unique because they embed the candidate position), the winning lane rescanning only its own stride-32 subset.
// Same keys, same order, same values as xl_topk_kernel.
constexpr int kSelWarps = 4;
constexpr int kSelKeys = 1024;   // merge kernel capacity (static shared memory)
constexpr int kSelKeysMax = 4096; // warp top-k capacity (dynamic shared memory, sized per launch)
constexpr int kSelSlots = 64;
constexpr int kSelK = 64;

struct SelScratch {  // view into the warp's slice of dynamic shared memory
    unsigned long long* keys;
    uint32_t* base;    // [kSelSlots + 1]
    uint32_t* colbeg;  // [kSelSlots]
    float* pval;       // [kSelSlots]
};

__host__ __device__ inline size_t sel_warp_bytes(uint32_t key_cap) {
    return static_cast<size_t>(key_cap) * 8 + (kSelSlots + 1 + kSelSlots + kSelSlots) * 4 + 12;  // padded to 16 below
}

__global__ void __launch_bounds__(kSelWarps * 32)
xl_topk_warp_kernel(const LayerDev L, const int pp_kind, const int pp_p, const int combine, const uint32_t k,
                    const uint32_t* __restrict__ beam_id, const float* __restrict__ beam_val,
                    const uint32_t* __restrict__ beam_cnt, const uint32_t beam_stride, const float* __restrict__ cand,
                    const uint64_t cand_stride_q, const uint32_t c_stride, uint32_t* __restrict__ out_id,
                    float* __restrict__ out_val, uint32_t* __restrict__ out_cnt, const uint32_t out_stride,
                    const uint32_t rows, unsigned long long* stats, unsigned long long* __restrict__ out_key,
                    const uint32_t key_cap) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kSelWarps + warp;
    if (q >= rows) return;
    const size_t slice = (sel_warp_bytes(key_cap) + 15) & ~static_cast<size_t>(15);
    SelScratch S;
    S.keys = reinterpret_cast<unsigned long long*>(smem_raw + warp * slice);
    S.base = reinterpret_cast<uint32_t*>(S.keys + key_cap);
    S.colbeg = S.base + (kSelSlots + 1);
    S.pval = reinterpret_cast<float*>(S.colbeg + kSelSlots);
    const uint32_t cnt = beam_cnt[q];
    for (uint32_t j = lane; j < cnt; j += 32) {
        const uint32_t p = beam_id[static_cast<uint64_t>(q) * beam_stride + j];
        const ChunkHeader h = L.chunks[p];
        S.base[j + 1] = h.n_cols;
        S.colbeg[j] = (h.has_bias & kChunkAbsent) ? 0xFFFFFFFFu : h.col_begin;
        S.pval[j] = beam_val[static_cast<uint64_t>(q) * beam_stride + j];
    }
    __syncwarp();
    uint32_t owned = 0;
    if (lane == 0) {
        uint32_t run = 0;
        S.base[0] = 0;
        for (uint32_t j = 0; j < cnt; ++j) {
            if (S.colbeg[j] != 0xFFFFFFFFu) owned += S.base[j + 1];
            run += S.base[j + 1];
            S.base[j + 1] = run;
        }
    }
    owned = __shfl_sync(kFull, owned, 0);
    __syncwarp();
    const uint32_t n_valid = S.base[cnt];
    const uint32_t kk = min(k, owned);
    if (lane == 0) {
        out_cnt[q] = kk;
        if (stats) atomicAdd(&stats[6], static_cast<unsigned long long>(kk));
    }
    if (n_valid == 0) return;
    const float* cq = cand + static_cast<uint64_t>(q) * cand_stride_q;
    auto score_at = [&](uint32_t cpos, uint32_t& label) -> float {
        const uint32_t j = static_cast<uint32_t>(last_le_u32(S.base, static_cast<int>(cnt), cpos));
        const uint32_t off = cpos - S.base[j];
        float v = xl_transform(cq[cpos], pp_kind, pp_p);
        if (combine) v = xl_combine(v, S.pval[j], pp_kind);
        label = S.colbeg[j] + off;
        return v;
    };
    // pass 1: raw scores -> shared memory; the candidates of a query are contiguous, so this is a flat coalesced copy
    // with eight independent loads in flight per lane
    for (uint32_t i0 = lane; i0 < n_valid; i0 += 256) {
        float r[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + 32u * u; r[u] = (i < n_valid) ? cq[i] : 0.0f; }
#pragma unroll
        for (int u = 0; u < 8; ++u) { const uint32_t i = i0 + 32u * u; if (i < n_valid) S.keys[i] = static_cast<unsigned long long>(__float_as_uint(r[u])); }
    }
    __syncwarp();
    if (owned != n_valid) {  // index sharding: the ranges of leaf chunks scored on other GPUs hold no data
        for (uint32_t j = 0; j < cnt; ++j) {
            if (S.colbeg[j] != 0xFFFFFFFFu) continue;
            for (uint32_t i = S.base[j] + lane; i < S.base[j + 1]; i += 32) S.keys[i] = 0xFFFFFFFFFFFFFFFFull;
        }
        __syncwarp();
    }
    // pass 2: post-processor (double-precision libm chains), combine, composite key.  Four independent candidates per lane
    // and step, so the long dependent exp/log chains of different candidates overlap.
    for (uint32_t i0 = lane; i0 < n_valid; i0 += 128) {
        unsigned long long raw[4];
        float pv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + 32u * u;
            raw[u] = 0xFFFFFFFFFFFFFFFFull;
            pv[u] = 0.0f;
            if (i < n_valid) {
                raw[u] = S.keys[i];
                pv[u] = S.pval[last_le_u32(S.base, static_cast<int>(cnt), i)];
            }
        }
        float v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = xl_transform(__uint_as_float(static_cast<uint32_t>(raw[u])), pp_kind, pp_p);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + 32u * u;
            if (i < n_valid) {
                float w = v[u];
                if (combine) w = xl_combine(w, pv[u], pp_kind);
                S.keys[i] = (raw[u] == 0xFFFFFFFFFFFFFFFFull) ? 0ull : xl_make_key(w, i);
            }
        }
    }
    __syncwarp();
    unsigned long long best = 0ull;  // lane-local maximum over positions lane, lane+32, ...
    for (uint32_t i = lane; i < n_valid; i += 32) { const unsigned long long key = S.keys[i]; best = key > best ? key : best; }
    __syncwarp();
    for (uint32_t r = 0; r < kk; ++r) {
        unsigned long long top = best;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, top, d);
            top = o > top ? o : top;
        }
        const uint32_t cpos = 0xFFFFFFFFu - static_cast<uint32_t>(top & 0xFFFFFFFFull);
        if ((cpos & 31u) == static_cast<uint32_t>(lane)) {
            uint32_t label;
            const float v = score_at(cpos, label);  // recomputed: keeps the exact bits (-0.0) of the stored value
            if (L.label_of_col) label = L.label_of_col[label];
            out_id[static_cast<uint64_t>(q) * out_stride + r] = label;
            out_val[static_cast<uint64_t>(q) * out_stride + r] = v;
            if (out_key) out_key[static_cast<uint64_t>(q) * out_stride + r] = top;
            S.keys[cpos] = 0ull;
            best = 0ull;
            for (uint32_t i = lane; i < n_valid; i += 32) { const unsigned long long key = S.keys[i]; best = key > best ? key : best; }
        }
        __syncwarp();
    }
}

#include "xlinear_topk_filter.cuh"

#define PB200_SELECTED_KERNELS
#include "xlinear_selected.cuh"
#undef PB200_SELECTED_KERNELS

// K3 (index sharding): merge the per-GPU top-k lists gathered by ONE all-gather into the global top-k.
// gathered layout: [world][rows][stride] for keys / ids / vals and [world][rows] for counts.  Keys are globally unique
// (they embed the candidate's position in the full prolongated row), so the merge is an exact arg-max selection.
__global__ void __launch_bounds__(kSelWarps * 32)
xl_merge_topk_kernel(const unsigned long long* __restrict__ g_keys, const uint32_t* __restrict__ g_ids,
                     const float* __restrict__ g_vals, const uint32_t* __restrict__ g_cnt, const uint32_t world,
                     const uint32_t rows, const uint32_t stride, const uint32_t k, uint32_t* __restrict__ out_id,
                     float* __restrict__ out_val, uint32_t* __restrict__ out_cnt) {
    __shared__ unsigned long long s_keys[kSelWarps][kSelKeys];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kSelWarps + warp;
    if (q >= rows) return;
    unsigned long long* keys = s_keys[warp];
    const uint32_t n = world * stride;
    unsigned long long best = 0ull;
    uint32_t total = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t g = i / stride, r = i - g * stride;
        const uint32_t c = g_cnt[static_cast<uint64_t>(g) * rows + q];
        unsigned long long key = 0ull;
        if (r < c) { key = g_keys[(static_cast<uint64_t>(g) * rows + q) * stride + r]; ++total; }
        keys[i] = key;
        best = key > best ? key : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(kFull, total, d);
    __syncwarp();
    const uint32_t kk = min(k, total);
    if (lane == 0) out_cnt[q] = kk;
    for (uint32_t rnk = 0; rnk < kk; ++rnk) {
        unsigned long long top = best;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, top, d);
            top = o > top ? o : top;
        }
        // the owner lane finds the slot of `top` among its stride-32 subset
        uint32_t slot = 0xFFFFFFFFu;
        if (best == top) {
            for (uint32_t i = lane; i < n; i += 32) if (keys[i] == top) { slot = i; break; }
        }
        if (slot != 0xFFFFFFFFu) {
            const uint32_t g = slot / stride, r = slot - g * stride;
            const uint64_t src = (static_cast<uint64_t>(g) * rows + q) * stride + r;
            out_id[static_cast<uint64_t>(q) * k + rnk] = g_ids[src];
            out_val[static_cast<uint64_t>(q) * k + rnk] = g_vals[src];
            keys[slot] = 0ull;
            best = 0ull;
            for (uint32_t i = lane; i < n; i += 32) { const unsigned long long key = keys[i]; best = key > best ? key : best; }
        }
        __syncwarp();
    }
}

// Packed exchange records for index sharding: ONE 16-byte {key, id, value} record per (query, rank) slot, key == 0 marks an
// empty slot (a valid key is never 0: its low word is ~position), so the per-query counts need not travel: the whole exchange
// is a single all-gather of one buffer.
struct __align__(16) ShardRecord {
    unsigned long long key;
    uint32_t id;
    float val;
};
static_assert(sizeof(ShardRecord) == 16, "shard record must stay 16 bytes");

__global__ void xl_shard_pack_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ ids,
                                     const float* __restrict__ vals, const uint32_t* __restrict__ cnt, const uint32_t rows,
                                     const uint32_t stride, ShardRecord* __restrict__ rec) {
    const uint64_t i = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= static_cast<uint64_t>(rows) * stride) return;
    const uint32_t q = static_cast<uint32_t>(i / stride), r = static_cast<uint32_t>(i - static_cast<uint64_t>(q) * stride);
    ShardRecord out{0ull, 0u, 0.0f};
    if (r < cnt[q]) { out.key = keys[i]; out.id = ids[i]; out.val = vals[i]; }
    rec[i] = out;
}

// merge of the gathered packed records [world][rows][stride]: same selection as xl_merge_topk_kernel
__global__ void __launch_bounds__(kSelWarps * 32)
xl_merge_packed_kernel(const ShardRecord* __restrict__ g_rec, const uint32_t world, const uint32_t rows, const uint32_t stride,
                       const uint32_t k, uint32_t* __restrict__ out_id, float* __restrict__ out_val, uint32_t* __restrict__ out_cnt) {
    __shared__ unsigned long long s_keys[kSelWarps][kSelKeys];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t q = blockIdx.x * kSelWarps + warp;
    if (q >= rows) return;
    unsigned long long* keys = s_keys[warp];
    const uint32_t n = world * stride;
    unsigned long long best = 0ull;
    uint32_t total = 0;
    for (uint32_t i = lane; i < n; i += 32) {
        const uint32_t g = i / stride, r = i - g * stride;
        const unsigned long long key = g_rec[(static_cast<uint64_t>(g) * rows + q) * stride + r].key;
        if (key) ++total;
        keys[i] = key;
        best = key > best ? key : best;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) total += __shfl_xor_sync(kFull, total, d);
    __syncwarp();
    const uint32_t kk = min(k, total);
    if (lane == 0) out_cnt[q] = kk;
    for (uint32_t rnk = 0; rnk < kk; ++rnk) {
        unsigned long long top = best;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const unsigned long long o = __shfl_xor_sync(kFull, top, d);
            top = o > top ? o : top;
        }
        uint32_t slot = 0xFFFFFFFFu;
        if (best == top) {
            for (uint32_t i = lane; i < n; i += 32) if (keys[i] == top) { slot = i; break; }
        }
        if (slot != 0xFFFFFFFFu) {
            const uint32_t g = slot / stride, r = slot - g * stride;
            const ShardRecord rc = g_rec[(static_cast<uint64_t>(g) * rows + q) * stride + r];
            out_id[static_cast<uint64_t>(q) * k + rnk] = rc.id;
            out_val[static_cast<uint64_t>(q) * k + rnk] = rc.val;
            keys[slot] = 0ull;
            best = 0ull;
            for (uint32_t i = lane; i < n; i += 32) { const unsigned long long key = keys[i]; best = key > best ? key : best; }
        }
        __syncwarp();
    }
}

// Row extents as one 8-byte record per chunk row, derived on the device from the row_ptr array of the chunk (the score
// kernels then need ONE load per matched row).  The chunk's region of meta[] holds R4 + roundup4(R + 1) >= 2R words, so
// rowext[] simply mirrors meta[]'s indexing.
__global__ void xl_build_rowext_kernel(const ChunkHeader* __restrict__ chunks, const uint32_t* __restrict__ meta,
                                       uint32_t* __restrict__ rowext, const uint32_t n_chunks) {
    const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
    const int lane = threadIdx.x & 31;
    for (uint32_t c = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; c < n_chunks; c += warps) {
        const ChunkHeader h = chunks[c];
        if (h.has_bias & kChunkAbsent) continue;
        const uint32_t R = h.nnz_rows;
        const uint32_t* rp = meta + h.meta_off + ((R + 3u) & ~3u);
        uint2* dst = reinterpret_cast<uint2*>(rowext + h.meta_off);
        for (uint32_t r = lane; r < R; r += 32) dst[r] = make_uint2(rp[r], rp[r + 1]);
    }
}

__global__ void xl_init_beam_kernel(uint32_t* beam_id, float* beam_val, uint32_t* beam_cnt, uint32_t beam_stride,
                                    uint32_t rows) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q < rows) {
        beam_id[static_cast<uint64_t>(q) * beam_stride] = 0u;   // prev_layer_pred = ones(Q x 1) (inference.hpp:2462-2463)
        beam_val[static_cast<uint64_t>(q) * beam_stride] = 1.0f;
        beam_cnt[q] = 1u;
    }
}

uint32_t max_row_nnz(const uint64_t* row_ptr, uint32_t rows) {
    uint64_t m = 0;
    for (uint32_t r = 0; r < rows; ++r) m = std::max<uint64_t>(m, row_ptr[r + 1] - row_ptr[r]);
    return static_cast<uint32_t>(std::min<uint64_t>(m, 0xFFFFFFFFull));
}

uint32_t next_pow2_host(uint64_t v) {
    uint64_t p = 2;
    while (p < v) p <<= 1;
    return static_cast<uint32_t>(p);
}

size_t chunk_kernel_smem(int warps, bool lookup, uint32_t q_cap, uint32_t sb_cap, uint32_t hdr_cap) {
    return static_cast<size_t>(q_cap) * 8 + static_cast<size_t>(sb_cap) * 4 + static_cast<size_t>(hdr_cap) * sizeof(ChunkHeader) +
           static_cast<size_t>(warps) * (lookup ? sizeof(WarpScratch<kMCapLookup>) : sizeof(WarpScratch<kMCap>));
}
size_t topk_kernel_smem(uint32_t b_prev) { return static_cast<size_t>(kSortCap) * 8 + (static_cast<size_t>(b_prev) * 3 + 1) * 4; }

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
// engine
// ------------------------------------------------------------------------------------------------------------------
XLinearEngine::XLinearEngine(std::unique_ptr<XLinearHostModel> host, int device) : host_(std::move(host)), device_(device) {
    PB200_CUDA(cudaSetDevice(device_));
    PB200_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& e : ev_) PB200_CUDA(cudaEventCreate(&e));
    PB200_CUDA(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking));
    for (auto& e : up_ev_) PB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    for (auto& e : use_ev_) PB200_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    if (const char* env = std::getenv("PB200_XL_PIPELINE")) {  // 0 = off, n >= 2 = number of sub-tiles (default 4)
        const int v = std::atoi(env);
        pipeline_uploads_ = v != 0;
        if (v >= 2) pipeline_parts_ = static_cast<uint32_t>(std::min(v, 64));
    }
    layers_.resize(host_->layers.size());
    uint64_t cmimg_budget = 8ull << 30;     // bytes of HBM the chunk images of the chunk-major kernel may take in total
    if (const char* env = std::getenv("PB200_CMIMG_MB")) cmimg_budget = std::strtoull(env, nullptr, 10) << 20;
    {
        int sms = 0;
        PB200_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_));
        n_sm_ = static_cast<uint32_t>(std::max(sms, 1));
    }
    uint64_t featmap_budget = 32ull << 30;  // bytes of HBM the feature maps may take in total
    if (const char* env = std::getenv("PB200_FEATMAP_MB")) featmap_budget = std::strtoull(env, nullptr, 10) << 20;
    for (size_t d = 0; d < layers_.size(); ++d) {
        auto& src = host_->layers[d];
        auto& dst = layers_[d];
        dst.chunks.upload(src.chunks.data(), src.chunks.size(), stream_);
        dst.meta.upload(src.meta.data(), src.meta.size(), stream_);
        dst.entries.upload(reinterpret_cast<const uint2*>(src.entries.data()), src.entries.size(), stream_);
        if (src.reordered) dst.label_of_col.upload(src.label_of_col.data(), src.label_of_col.size(), stream_);
        // query-driven lookup structure, unless it would blow the memory budget (then the kernel streams the row lists)
        dst.view.featmap = nullptr;
        dst.view.fm_words = 0;
        if (featmap_budget >= feature_map_bytes(src)) {
            featmap_budget -= feature_map_bytes(src);
            build_feature_map(src);
            dst.featmap.upload(reinterpret_cast<const uint2*>(src.featmap.data()), src.featmap.size(), stream_);
            PB200_CUDA(cudaStreamSynchronize(stream_));
            dst.view.featmap = dst.featmap.get();
            dst.view.fm_words = src.fm_words;
            model_bytes_ += src.featmap.size() * 8;
            std::vector<uint2_host>().swap(src.featmap);
        }
        dst.e_max = 0;
        for (const ChunkHeader& ch : src.chunks) {
            if ((ch.has_bias & kChunkAbsent) || ch.nnz_rows == 0) continue;
            const uint32_t* rp = src.meta.data() + ch.meta_off + round_up4(ch.nnz_rows);
            dst.e_max = std::max(dst.e_max, rp[ch.nnz_rows]);
        }
        dst.rowext.reserve(std::max<uint64_t>(src.meta.size(), 4));
        xl_build_rowext_kernel<<<148 * 8, 256, 0, stream_>>>(dst.chunks.get(), dst.meta.get(), dst.rowext.get(), src.n_chunks);
        PB200_CUDA(cudaGetLastError());
        model_bytes_ += src.meta.size() * 4;
        dst.view.chunks = dst.chunks.get();
        dst.view.meta = dst.meta.get();
        dst.view.rowext = dst.rowext.get();
        dst.view.entries = dst.entries.get();
        dst.view.label_of_col = src.reordered ? dst.label_of_col.get() : nullptr;
        dst.view.n_cols = src.n_cols;
        dst.view.n_chunks = src.n_chunks;
        dst.view.c_max = src.c_max;
        dst.view.w_rows = src.w_rows;
        dst.view.bias = src.bias;
        dst.view.has_dup_cols = src.has_dup_cols ? 1 : 0;
        // chunk images for the chunk-major score kernel (xlinear_cm_kernel.cuh), where the layer's shape allows it.  A column
        // cap cuts the widest chunks of the layer into column ranges ("virtual chunks") when the layer does not fit uncut.
        dst.cm_shape = CmShape{};
        if (dst.view.featmap) {
            auto layout_for_cap = [&](uint32_t cap, std::vector<uint32_t>* vc_ptr_out, uint32_t* e_max_out) {
                uint32_t n_vc = 0, best = 0;
                std::vector<uint32_t> cnt;
                if (vc_ptr_out) vc_ptr_out->assign(static_cast<size_t>(src.n_chunks) + 1, 0u);
                for (uint32_t p = 0; p < src.n_chunks; ++p) {
                    const ChunkHeader& ch = src.chunks[p];
                    if (vc_ptr_out) (*vc_ptr_out)[p] = n_vc;
                    if ((ch.has_bias & kChunkAbsent) || ch.n_cols == 0) continue;
                    const uint32_t nr = (ch.n_cols + cap - 1) / cap;
                    n_vc += nr;
                    if (!e_max_out || ch.nnz_rows == 0) continue;
                    const uint32_t* rp = src.meta.data() + ch.meta_off + round_up4(ch.nnz_rows);
                    const uint32_t width = (ch.n_cols + nr - 1) / nr;
                    cnt.assign(nr, 0u);
                    const ChunkEntry* en = src.entries.data() + ch.ent_off;
                    for (uint32_t i = 0; i < rp[ch.nnz_rows]; ++i) ++cnt[std::min(en[i].col_offset / width, nr - 1)];
                    for (uint32_t v : cnt) best = std::max(best, v);
                }
                if (vc_ptr_out) (*vc_ptr_out)[src.n_chunks] = n_vc;
                if (e_max_out) *e_max_out = best;
                return n_vc;
            };
            // Policy (measured, eurlex-4k leaf: chunks of 62 +- 8 columns, widest 85): uncut 6 warps 0.78 ms; cap 64 (a third of
            // the chunks cut) 10 warps 0.94 ms; every chunk in two ranges 14 warps 1.04 ms -- more warps raise the issue rate
            // (32 -> 43 %) but the repeated lookups (+21 % instructions) and the uneven cost of cut / uncut pairs inside the
            // equal-count CTA shares cost more.  So: the LARGEST cap that fits at all (no cut whenever the layer fits uncut).
            CmShape shape;
            const uint32_t n_real = layout_for_cap(std::max<uint32_t>(src.c_max, 1u), nullptr, nullptr);
            for (uint32_t cap = std::max<uint32_t>(src.c_max, 1u); n_real > 0; cap = cap * 7 / 8) {
                const uint32_t n_vc = layout_for_cap(cap, nullptr, nullptr);
                if (static_cast<uint64_t>(n_vc) * 100u > static_cast<uint64_t>(n_real) * kCmMaxDup) break;
                uint32_t e_cap = dst.e_max;
                if (cap < src.c_max) layout_for_cap(cap, nullptr, &e_cap);
                const CmShape cand = cm_shape(src.fm_words, src.w_rows, src.r_max, e_cap, cap, src.n_chunks, n_vc);
                if (cand.ok) { shape = cand; break; }
                if (cap < 8u) break;
            }
            const uint64_t bytes = static_cast<uint64_t>(shape.img_bytes) * shape.n_vc;
            if (shape.ok && bytes <= cmimg_budget) {
                cmimg_budget -= bytes;
                std::vector<uint32_t> vc_ptr;
                layout_for_cap(shape.col_cap, &vc_ptr, nullptr);
                dst.cm_vc_ptr.upload(vc_ptr.data(), vc_ptr.size(), stream_);
                PB200_CUDA(cudaStreamSynchronize(stream_));
                shape.vc_ptr = dst.cm_vc_ptr.get();
                dst.cm_shape = shape;
                dst.cm_images.reserve(std::max<uint64_t>(bytes, 1));
                xl_cm_build_images_kernel<<<shape.n_vc, 256, 0, stream_>>>(dst.view, shape, dst.cm_images.get());
                PB200_CUDA(cudaGetLastError());
                model_bytes_ += bytes;
            }
        }
        model_bytes_ += src.chunks.size() * sizeof(ChunkHeader) + src.meta.size() * 4 + src.entries.size() * 8 +
                        src.label_of_col.size() * 4;
    }
    PB200_CUDA(cudaStreamSynchronize(stream_));
    // host copies of the big arrays are no longer needed
    for (auto& l : host_->layers) {
        std::vector<uint32_t>().swap(l.meta);
        std::vector<ChunkEntry>().swap(l.entries);
    }
    const int max_smem = 200 * 1024;
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_chunk_scores_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_topk_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_topk_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_cm_scores_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCmSmemBudget)));
    PB200_CUDA(cudaFuncSetAttribute(xl_cm_scores_kernel<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCmSmemBudget)));
    PB200_CUDA(cudaFuncSetAttribute(xl_cm_scores_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCmSmemBudget)));
    PB200_CUDA(cudaFuncSetAttribute(xl_cm_scores_kernel<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCmSmemBudget)));
    PB200_CUDA(cudaFuncSetAttribute(xl_cmg_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kCmSmemBudget)));
    PB200_CUDA(cudaFuncSetAttribute(xl_query_warp_scores_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(xl_query_warp_scores_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    stats_dev_.reserve(8 * layers_.size());
    if (const char* env = std::getenv("PB200_XL_KERNEL_MODE")) set_kernel_mode(std::atoi(env));  // A/B runs of bench.py
    layer_profile_.assign(layers_.size(), XLinearLayerProfile{});
    layer_stats_.assign(layers_.size(), XLinearStats{});
}

XLinearEngine::~XLinearEngine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
    if (copy_stream_) cudaStreamSynchronize(copy_stream_);
    for (auto& e : up_ev_) if (e) cudaEventDestroy(e);
    for (auto& e : use_ev_) if (e) cudaEventDestroy(e);
    if (copy_stream_) cudaStreamDestroy(copy_stream_);
    if (stream_) cudaStreamDestroy(stream_);
}

void XLinearEngine::set_kernel_mode(int mode) {
    // 0: first generation (row-list streaming + block-wide sort); 1: default (query-warp / feature-map kernels + warp top-k);
    // 2: feature-map lookups with one warp per chunk (no query-warp kernel); 3: query-warp kernel wherever eligible;
    // 4: as 1 but the warp top-k evaluates the post-processor for every candidate (no estimate filter);
    // 5: as 1, and the chunk-major score kernel wherever it FITS (its reuse / occupancy heuristics ignored; tests);
    // 6: as 1 WITHOUT the chunk-major score kernel (query-major kernels only, for A/B tests)
    const bool on = mode != 0;
    for (auto& l : layers_) l.view.featmap = (on && l.featmap.capacity()) ? l.featmap.get() : nullptr;
    force_block_topk_ = !on;
    no_query_warp_ = (mode == 2);
    force_query_warp_ = (mode == 3);
    no_topk_filter_ = (mode == 4);
    chunk_major_ = on && (mode != 6);
    cm_force_ = (mode == 5);
    cmg_ = (mode == 8 || mode == 9 || mode == 10);  // 9: as 1 plus the image-less lane-per-pair kernel on lookup layers without an image
    cmg_all_ = (mode == 8 || mode == 10);  // 8: image-less lane-per-pair kernel also in place of the query-warp kernel
    cm_image_ = (mode != 10);              // 10: as 8, and in place of the staged-image kernel too (tests: every layer on it)
}

bool XLinearEngine::has_feature_maps() const {
    for (auto& l : layers_) if (!l.featmap.capacity()) return false;
    return true;
}

void XLinearEngine::reset_profile() {
    layer_profile_.assign(layers_.size(), XLinearLayerProfile{});
    launches_ = 0;
}

std::vector<XLinearEngine::LayerPlan> XLinearEngine::make_plan_(uint32_t beam_size, const char* post_processor,
                                                                 uint32_t only_topk) const {
    const size_t depth = layers_.size();
    std::vector<LayerPlan> plan(depth);
    uint32_t b_prev = 1;
    for (size_t d = 0; d < depth; ++d) {
        const auto& L = host_->layers[d];
        // local_only_topk (inference.hpp:2471) then only_topk_to_use (inference.hpp:2055)
        const uint32_t local = (d + 1 == depth) ? only_topk : beam_size;
        const uint32_t k = local > 0 ? local : static_cast<uint32_t>(L.only_topk);
        plan[d].k = k;
        plan[d].pp = post_processor ? parse_post_processor(post_processor) : L.post_processor;
        plan[d].b_prev = b_prev;
        const uint64_t cand_max = static_cast<uint64_t>(b_prev) * std::max<uint32_t>(L.c_max, 1u);
        plan[d].k_cap = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(k, cand_max)));
        if (topk_kernel_smem(b_prev) > 200u * 1024u)
            throw std::runtime_error("pecos_b200: beam of " + std::to_string(b_prev) + " nodes exceeds the supported maximum");
        b_prev = plan[d].k_cap;
    }
    return plan;
}

uint32_t XLinearEngine::pick_tile_rows_(const std::vector<LayerPlan>& plan, uint32_t rows) const {
    uint64_t budget = 8ull << 30;
    if (const char* env = std::getenv("PB200_WORKSPACE_MB")) budget = std::max<uint64_t>(64, std::strtoull(env, nullptr, 10)) << 20;
    uint64_t per_query = 16;
    uint64_t beam_stride = 1, cand_max = 1, sort_max = 0;
    for (size_t d = 0; d < plan.size(); ++d) {
        const uint64_t c = static_cast<uint64_t>(plan[d].b_prev) * std::max<uint32_t>(host_->layers[d].c_max, 1u);
        cand_max = std::max(cand_max, c);
        beam_stride = std::max<uint64_t>(beam_stride, plan[d].k_cap);
        if (c > static_cast<uint64_t>(kSortCap) && plan[d].k > static_cast<uint32_t>(kSortCap / 2)) sort_max = std::max<uint64_t>(sort_max, next_pow2_host(c));
    }
    per_query += 2 * beam_stride * 8 + cand_max * 4 + sort_max * 8;
    uint64_t tile = std::max<uint64_t>(1, budget / per_query);
    return static_cast<uint32_t>(std::min<uint64_t>(tile, std::max<uint32_t>(rows, 1u)));
}

void XLinearEngine::ensure_workspace_(const std::vector<LayerPlan>& plan, uint32_t tile_rows) {
    uint64_t beam_stride = 1, cand_max = 1, sort_max = 0;
    for (size_t d = 0; d < plan.size(); ++d) {
        const uint64_t c = static_cast<uint64_t>(plan[d].b_prev) * std::max<uint32_t>(host_->layers[d].c_max, 1u);
        cand_max = std::max(cand_max, c);
        beam_stride = std::max<uint64_t>(beam_stride, std::max<uint32_t>(plan[d].k_cap, plan[d].b_prev));
        if (c > static_cast<uint64_t>(kSortCap) && plan[d].k > static_cast<uint32_t>(kSortCap / 2)) sort_max = std::max<uint64_t>(sort_max, next_pow2_host(c));
    }
    beam_stride_ = static_cast<uint32_t>(beam_stride);
    for (int b = 0; b < 2; ++b) {
        beam_id_[b].reserve(static_cast<uint64_t>(tile_rows) * beam_stride);
        beam_val_[b].reserve(static_cast<uint64_t>(tile_rows) * beam_stride);
        beam_cnt_[b].reserve(tile_rows);
    }
    if (chunk_major_ && has_feature_maps()) {
        uint64_t b_max = 1, chunks_max = 1;
        for (size_t d = 0; d < plan.size(); ++d) {
            b_max = std::max<uint64_t>(b_max, plan[d].b_prev);
            chunks_max = std::max<uint64_t>(chunks_max, host_->layers[d].n_chunks);
        }
        cm_slot_pos_.reserve(static_cast<uint64_t>(tile_rows) * beam_stride);
        cm_pair_q_.reserve(static_cast<uint64_t>(tile_rows) * b_max * 4);   // up to 4 column ranges per chunk
        cm_pair_pos_.reserve(static_cast<uint64_t>(tile_rows) * b_max * 4);
        cm_count_.reserve(chunks_max * 4 + 1);
        cm_bucket_ptr_.reserve(chunks_max * 4 + 1);
        cm_item_ptr_.reserve(chunks_max * 4 + 1);
    }
    cand_.reserve(static_cast<uint64_t>(tile_rows) * cand_max);
    if (sort_max) sortbuf_.reserve(static_cast<uint64_t>(tile_rows) * sort_max);
}

// Launches the score kernel of layer d for the beam held in beam_*_[cur] (capacity b_prev slots per query): fills
// cand_[q * b_prev * c_max + slot_base(j) + c] with the raw scores of every child of every beam node, in prolongation order.
// Chooses between the chunk-major, query-warp, feature-map, row-list and dense kernels.  Returns the kernel id.
int XLinearEngine::score_layer_(size_t d, const QueryDev& q, uint32_t b_prev, int cur, bool collect_stats) {
    const LayerDev& L = layers_[d].view;
    const uint32_t rows = q.rows;
    const bool dense = (q.row_ptr == nullptr);
    const uint32_t c_stride = std::max<uint32_t>(L.c_max, 1u);
    const uint64_t cand_stride_q = static_cast<uint64_t>(b_prev) * c_stride;
    unsigned long long* stats = collect_stats ? stats_dev_.get() + 8 * d : nullptr;
    // spread the beam slots evenly: b = 10 -> 10 warps x 1 slot, b = 20 -> 10 warps x 2 slots
    const uint32_t rounds = (b_prev + kWarpsMax - 1) / kWarpsMax;
    const int warps = static_cast<int>(std::max<uint32_t>(1, (b_prev + rounds - 1) / std::max<uint32_t>(rounds, 1)));
    const dim3 grid(rows), block(warps * 32);
    const bool lookup = !dense && L.featmap != nullptr;
    // query staging area: as small as the batch's longest row allows (occupancy), at most kQCap non-zeros
    const uint32_t q_cap = dense ? 32u : std::min<uint32_t>(kQCap, std::max<uint32_t>(32u, (q.max_row_nnz + 31u) & ~31u));
    const uint32_t sb_cap = (b_prev + 1u + 3u) & ~3u;
    const uint32_t hdr_cap = b_prev <= 128u ? b_prev : 0u;  // beam chunk headers cached in shared memory
    const size_t smem1 = chunk_kernel_smem(warps, lookup, q_cap, sb_cap, hdr_cap);
    auto launch = [&](auto kernel) {
        kernel<<<grid, block, smem1, stream_>>>(L, q, bid_(cur), bcnt_(cur), beam_stride_, cand_at_(cand_stride_q),
                                               cand_stride_q, c_stride, stats, q_cap, sb_cap, hdr_cap);
    };
    // One warp per query over the whole beam (feature-major).  Measured on B200: it wins when the beam consists of
    // MANY NARROW chunks (per-chunk bookkeeping dominates: S layers 1-4, 20 x 8 columns: 2.1-2.9 ms vs 3.1 ms), and
    // loses on wide chunks where one warp per chunk keeps more loads in flight (E leaf 2.2 vs 1.45 ms, S leaf 12.7 vs
    // 8.4 ms).  force_query_warp_ (kernel mode 3) selects it whenever it is eligible, for tests.
    const bool qw_eligible = lookup && b_prev <= static_cast<uint32_t>(kQwSlots) &&
                             cand_stride_q <= static_cast<uint64_t>(kQwNCap) && q.max_row_nnz <= kQwQCap;
    const bool query_warp = qw_eligible && !no_query_warp_ &&
                            (force_query_warp_ || (b_prev >= 16u && cand_stride_q <= 256u));
    // Chunk-major scoring (xlinear_cm_kernel.cuh) wherever the layer's feature map + largest chunk fit in shared memory
    // and the chunks are visited by enough pairs to amortise the staging; otherwise the query-major kernels below.
    // (the statistics pass always runs the query-major kernels: their counters are the canonical ones)
    const bool cm_offsets_fit = static_cast<uint64_t>(rows) * std::max<uint32_t>(q.max_row_nnz, 1u) < (1ull << 32);  // 32-bit feature offsets
    const CmPlan cm = (chunk_major_ && cm_image_ && lookup && !collect_stats && cm_offsets_fit && cm_slot_pos_.capacity() && layers_[d].cm_images.capacity())
                          ? cm_plan(layers_[d].cm_shape, L.n_chunks, static_cast<uint64_t>(rows) * b_prev, n_sm_, cm_force_)
                          : CmPlan{};
    const bool chunk_major = cm.eligible;
    // the same lane-per-pair walk without a staged image (xl_cmg_scores_kernel) where the image variant does not apply: layers
    // of large feature spaces / chunks visited by few pairs.  OPT-IN (kernel modes 8-10): bit-exact but measured slower than the
    // query-major kernels on the 3M-label model (leaf 16.4 vs 5.4 ms: 92 registers + 224 KB keep 12 warps per SM, every lookup a
    // dependent global load; profiles/r02_l_ncu_cmg.txt).  Mode 9: in place of the feature-map chunk kernel; 8: also of the
    // query-warp kernel.
    const CmgPlan cmg = (chunk_major_ && cmg_ && lookup && !collect_stats && !chunk_major && cm_offsets_fit && cm_slot_pos_.capacity() &&
                         (cmg_all_ || !query_warp))
                            ? cmg_plan(L.c_max, layers_[d].e_max, L.n_chunks, static_cast<uint64_t>(rows) * b_prev, n_sm_, cm_force_ || cmg_all_)
                            : CmgPlan{};
    const bool chunk_major_global = cmg.eligible;
    if (chunk_major_global) {
        CmWork w{cm_slot_pos_.get(), cm_count_.get(), cm_bucket_ptr_.get(), cm_item_ptr_.get(), cm_pair_q_.get(), cm_pair_pos_.get(),
                 cmg.warps * 32u};
        PB200_CUDA(cudaMemsetAsync(w.count, 0, (static_cast<uint64_t>(L.n_chunks) + 1) * 4, stream_));
        const uint32_t warp_grid = (rows * 32u + 127u) / 128u;
        xl_cm_count_kernel<<<warp_grid, 128, 0, stream_>>>(L, q, bid_(cur), bcnt_(cur), beam_stride_, rows, w, nullptr);
        xl_cm_scan_kernel<<<1, 1024, 0, stream_>>>(L.n_chunks, w);
        xl_cm_scatter_kernel<<<warp_grid, 128, 0, stream_>>>(L, bid_(cur), bcnt_(cur), beam_stride_, rows, w, nullptr);
        xl_cmg_scores_kernel<<<cmg.grid, cmg.warps * 32, cmg.smem, stream_>>>(L, q, w, cand_at_(cand_stride_q), cand_stride_q, L.c_max);
        launches_ += 3;
    } else if (chunk_major) {
        const CmShape& shape = layers_[d].cm_shape;
        const uint32_t n_vc = shape.n_vc;
        CmWork w{cm_slot_pos_.get(), cm_count_.get(), cm_bucket_ptr_.get(), cm_item_ptr_.get(), cm_pair_q_.get(), cm_pair_pos_.get(),
                 cm.warps * 32u};
        PB200_CUDA(cudaMemsetAsync(w.count, 0, (static_cast<uint64_t>(n_vc) + 1) * 4, stream_));
        const uint32_t warp_grid = (rows * 32u + 127u) / 128u;
        xl_cm_count_kernel<<<warp_grid, 128, 0, stream_>>>(L, q, bid_(cur), bcnt_(cur), beam_stride_, rows, w, shape.vc_ptr);
        xl_cm_scan_kernel<<<1, 1024, 0, stream_>>>(n_vc, w);
        xl_cm_scatter_kernel<<<warp_grid, 128, 0, stream_>>>(L, bid_(cur), bcnt_(cur), beam_stride_, rows, w, shape.vc_ptr);
        auto launch_cm = [&](auto kernel) {
            kernel<<<cm.grid, cm.warps * 32, cm.smem, stream_>>>(L, q, w, shape, layers_[d].cm_images.get(), cand_at_(cand_stride_q), cand_stride_q);
        };
        if (shape.direct) { if (shape.stages == 4) launch_cm(xl_cm_scores_kernel<true, 4>); else launch_cm(xl_cm_scores_kernel<true, 2>); }
        else { if (shape.stages == 4) launch_cm(xl_cm_scores_kernel<false, 4>); else launch_cm(xl_cm_scores_kernel<false, 2>); }
        launches_ += 3;  // + the score kernel counted below
    } else if (query_warp) {
        const uint32_t qw_qcap = std::max<uint32_t>(32u, (q.max_row_nnz + 31u) & ~31u);
        const uint32_t qw_ncap = static_cast<uint32_t>((cand_stride_q + 31) & ~static_cast<uint64_t>(31));
        const size_t qw_smem = kQwWarps * ((qw_warp_bytes(qw_qcap, qw_ncap) + 15) & ~static_cast<size_t>(15));
        const dim3 qw_grid((rows + kQwWarps - 1) / kQwWarps);
        if (collect_stats)
            xl_query_warp_scores_kernel<true><<<qw_grid, kQwWarps * 32, qw_smem, stream_>>>(
                L, q, bid_(cur), bcnt_(cur), beam_stride_, cand_at_(cand_stride_q), cand_stride_q, stats, qw_qcap, qw_ncap, rows);
        else
            xl_query_warp_scores_kernel<false><<<qw_grid, kQwWarps * 32, qw_smem, stream_>>>(
                L, q, bid_(cur), bcnt_(cur), beam_stride_, cand_at_(cand_stride_q), cand_stride_q, stats, qw_qcap, qw_ncap, rows);
    } else if (dense) {
        if (collect_stats) launch(xl_chunk_scores_kernel<true, true, false>);
        else launch(xl_chunk_scores_kernel<true, false, false>);
    } else if (lookup) {
        if (collect_stats) launch(xl_chunk_scores_kernel<false, true, true>);
        else launch(xl_chunk_scores_kernel<false, false, true>);
    } else {
        if (collect_stats) launch(xl_chunk_scores_kernel<false, true, false>);
        else launch(xl_chunk_scores_kernel<false, false, false>);
    }
    PB200_CUDA(cudaGetLastError());
    ++launches_;
    layer_profile_[d].scores_kernel = chunk_major_global ? 5 : chunk_major ? 4 : query_warp ? 3 : dense ? 2 : lookup ? 1 : 0;
    return layer_profile_[d].scores_kernel;
}

// Runs every layer over one tile of queries.  The last layer writes into res_*_dev_ at row offset res_row0_.
// ext_beam: beam_*_[0] already hold the beam entering the first layer (single-layer entry point); combine_first: that
// layer combines its scores with the beam values (a previous prediction was given).
void XLinearEngine::run_tile_(const QueryDev& q, const std::vector<LayerPlan>& plan, bool collect_stats, bool ext_beam,
                              int combine_first, size_t d_begin, size_t d_end) {
    const uint32_t rows = q.rows;
    if (rows == 0) return;
    const bool dense = (q.row_ptr == nullptr);
    int cur = static_cast<int>(d_begin & 1);  // every layer flips the ping-pong beam buffers once
    if (!ext_beam && d_begin == 0) {
        xl_init_beam_kernel<<<(rows + 255) / 256, 256, 0, stream_>>>(bid_(cur), bval_(cur),
                                                                     bcnt_(cur), beam_stride_, rows);
        ++launches_;
    }
    const size_t depth = plan.size();
    d_end = std::min(d_end, depth);
    for (size_t d = d_begin; d < d_end; ++d) {
        const LayerDev& L = layers_[d].view;
        const LayerPlan& lp = plan[d];
        const int combine = (d == 0) ? combine_first : 1;
        const uint32_t c_stride = std::max<uint32_t>(L.c_max, 1u);
        const uint64_t cand_stride_q = static_cast<uint64_t>(lp.b_prev) * c_stride;
        unsigned long long* stats = collect_stats ? stats_dev_.get() + 8 * d : nullptr;
        if (profile_) PB200_CUDA(cudaEventRecord(ev_[0], stream_));
        const dim3 grid(rows);
        const bool chunk_major = score_layer_(d, q, lp.b_prev, cur, collect_stats) == 4;
        (void)chunk_major;
        if (profile_) PB200_CUDA(cudaEventRecord(ev_[1], stream_));

        const bool last = (d + 1 == depth);
        uint32_t* o_id; float* o_val; uint32_t* o_cnt; uint32_t o_stride;
        unsigned long long* o_key = nullptr;
        if (last && ext_ids_) {  // index-sharded run: local top-k goes straight into the caller's (NCCL send) buffers
            o_id = ext_ids_ + static_cast<uint64_t>(res_rows_) * res_stride_;
            o_val = ext_vals_ + static_cast<uint64_t>(res_rows_) * res_stride_;
            o_cnt = ext_cnt_ + res_rows_;
            o_key = ext_keys_ + static_cast<uint64_t>(res_rows_) * res_stride_;
            o_stride = res_stride_;
        } else if (last) {
            o_id = res_ids_dev_.get() + static_cast<uint64_t>(res_rows_) * res_stride_;
            o_val = res_vals_dev_.get() + static_cast<uint64_t>(res_rows_) * res_stride_;
            o_cnt = res_cnt_dev_.get() + res_rows_;
            o_stride = res_stride_;
        } else {
            o_id = bid_(cur ^ 1); o_val = bval_(cur ^ 1); o_cnt = bcnt_(cur ^ 1);
            o_stride = beam_stride_;
        }
        const uint64_t sort_stride = next_pow2_host(cand_stride_q);
        const bool warp_select = !force_block_topk_ && lp.b_prev <= static_cast<uint32_t>(kSelSlots) &&
                                 cand_stride_q <= static_cast<uint64_t>(kSelKeysMax) && lp.k <= static_cast<uint32_t>(kSelK);
        const bool hinge = lp.pp.kind == PP_LP_HINGE || lp.pp.kind == PP_LOG_LP_HINGE;
        const bool filter_select = !force_block_topk_ && !no_topk_filter_ && lp.k <= 32u &&
                                   lp.b_prev <= static_cast<uint32_t>(kFltSlots) &&
                                   cand_stride_q <= static_cast<uint64_t>(kFltKeysMax) &&
                                   (!hinge || (lp.pp.p >= 0 && lp.pp.p <= 4));
        if (filter_select) {
            const uint32_t key_cap = static_cast<uint32_t>((cand_stride_q + 127) & ~static_cast<uint64_t>(127));
            const size_t flt_smem = kFltWarps * flt_warp_bytes(key_cap);
            xl_topk_filter_kernel<<<(rows + kFltWarps - 1) / kFltWarps, kFltWarps * 32, flt_smem, stream_>>>(
                L, lp.pp.kind, lp.pp.p, combine, lp.k, bid_(cur), bval_(cur), bcnt_(cur),
                beam_stride_, cand_at_(cand_stride_q), cand_stride_q, o_id, o_val, o_cnt, o_stride, rows, stats, o_key, key_cap);
        } else if (warp_select) {
            const uint32_t key_cap = static_cast<uint32_t>((cand_stride_q + 31) & ~static_cast<uint64_t>(31));
            const size_t sel_smem = kSelWarps * ((sel_warp_bytes(key_cap) + 15) & ~static_cast<size_t>(15));
            xl_topk_warp_kernel<<<(rows + kSelWarps - 1) / kSelWarps, kSelWarps * 32, sel_smem, stream_>>>(
                L, lp.pp.kind, lp.pp.p, combine, lp.k, bid_(cur), bval_(cur), bcnt_(cur),
                beam_stride_, cand_at_(cand_stride_q), cand_stride_q, c_stride, o_id, o_val, o_cnt, o_stride, rows, stats, o_key, key_cap);
        } else {
            xl_topk_kernel<<<grid, kTopkThreads, topk_kernel_smem(lp.b_prev), stream_>>>(
                L, lp.pp.kind, lp.pp.p, combine, lp.k, bid_(cur), bval_(cur), bcnt_(cur),
                beam_stride_, cand_at_(cand_stride_q), cand_stride_q, c_stride, o_id, o_val, o_cnt, o_stride, sortbuf_.get(), sort_stride,
                lp.b_prev, stats, o_key);
        }
        PB200_CUDA(cudaGetLastError());
        ++launches_;
        layer_profile_[d].topk_kernel = filter_select ? 2 : warp_select ? 1 : 0;
        if (profile_) {
            PB200_CUDA(cudaEventRecord(ev_[2], stream_));
            PB200_CUDA(cudaEventSynchronize(ev_[2]));
            float a = 0.f, b = 0.f;
            PB200_CUDA(cudaEventElapsedTime(&a, ev_[0], ev_[1]));
            PB200_CUDA(cudaEventElapsedTime(&b, ev_[1], ev_[2]));
            layer_profile_[d].scores_ms += a;
            layer_profile_[d].topk_ms += b;
            layer_profile_[d].launches += 2;
        }
        cur ^= 1;
    }
}

XLinearEngine::Result XLinearEngine::finish_result_(uint32_t rows, uint32_t stride) {
    out_ids_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    out_vals_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    out_cnt_.reserve(static_cast<uint64_t>(rows) + 1);
    if (rows) {
        PB200_CUDA(cudaMemcpyAsync(out_ids_.get(), res_ids_dev_.get(), static_cast<uint64_t>(rows) * stride * 4, cudaMemcpyDeviceToHost, stream_));
        PB200_CUDA(cudaMemcpyAsync(out_vals_.get(), res_vals_dev_.get(), static_cast<uint64_t>(rows) * stride * 4, cudaMemcpyDeviceToHost, stream_));
        PB200_CUDA(cudaMemcpyAsync(out_cnt_.get(), res_cnt_dev_.get(), static_cast<uint64_t>(rows) * 4, cudaMemcpyDeviceToHost, stream_));
    }
    PB200_CUDA(cudaStreamSynchronize(stream_));
    Result r;
    r.rows = rows;
    r.stride = stride;
    r.out_cols = host_->layers.back().out_cols;
    r.ids = out_ids_.get();
    r.vals = out_vals_.get();
    r.cnt = out_cnt_.get();
    return r;
}

XLinearEngine::Result XLinearEngine::predict_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val,
                                                 uint32_t rows, uint32_t cols, uint32_t beam_size,
                                                 const char* post_processor, uint32_t only_topk) {
    PB200_CUDA(cudaSetDevice(device_));
    const auto plan = make_plan_(beam_size, post_processor, only_topk);
    const uint32_t tile = pick_tile_rows_(plan, rows);
    ensure_workspace_(plan, tile);
    const uint32_t stride = plan.back().k_cap;
    res_stride_ = stride;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    // Whole batch fits the workspace (the common case): the rows are uploaded in `parts` chunks on the copy stream into ONE
    // staging set; the UPPER layers (cheap) run per chunk as it lands, overlapping the rest of the upload, and the LAST layer
    // -- where the time goes, and where the chunk-major kernel wants as many pairs per launch as it can get -- runs once over
    // the whole batch.
    if (pipeline_uploads_ && rows >= 4096u && tile >= rows) {
        const size_t depth = plan.size();
        const uint64_t nnz0 = row_ptr[0], nnz = row_ptr[rows] - nnz0;
        x_row_ptr_.reserve(static_cast<uint64_t>(rows) + 1);
        x_col_idx_.reserve(nnz);
        x_val_.reserve(nnz);
        const uint32_t part = ((rows + pipeline_parts_ - 1u) / pipeline_parts_ + 31u) & ~31u;
        std::vector<cudaEvent_t> evs;
        for (uint32_t r0 = 0; r0 < rows; r0 += part) {
            const uint32_t tr = std::min(part, rows - r0);
            const uint64_t b = row_ptr[r0] - nnz0, e = row_ptr[r0 + tr] - nnz0;
            PB200_CUDA(cudaMemcpyAsync(x_row_ptr_.get() + r0, row_ptr + r0, (static_cast<uint64_t>(tr) + 1) * 8, cudaMemcpyHostToDevice, copy_stream_));
            PB200_CUDA(cudaMemcpyAsync(x_col_idx_.get() + b, col_idx + nnz0 + b, (e - b) * 4, cudaMemcpyHostToDevice, copy_stream_));
            PB200_CUDA(cudaMemcpyAsync(x_val_.get() + b, val + nnz0 + b, (e - b) * 4, cudaMemcpyHostToDevice, copy_stream_));
            cudaEvent_t ev;
            PB200_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            PB200_CUDA(cudaEventRecord(ev, copy_stream_));
            evs.push_back(ev);
        }
        size_t t = 0;
        for (uint32_t r0 = 0; r0 < rows; r0 += part, ++t) {
            const uint32_t tr = std::min(part, rows - r0);
            PB200_CUDA(cudaStreamWaitEvent(stream_, evs[t], 0));
            if (depth > 1) {
                QueryDev q{x_row_ptr_.get() + r0, x_col_idx_.get(), x_val_.get(), nnz0, tr, cols, max_row_nnz(row_ptr + r0, tr)};
                row_off_ = r0;
                res_rows_ = r0;
                run_tile_(q, plan, false, false, 0, 0, depth - 1);
            }
        }
        row_off_ = 0;
        res_rows_ = 0;
        QueryDev q{x_row_ptr_.get(), x_col_idx_.get(), x_val_.get(), nnz0, rows, cols, max_row_nnz(row_ptr, rows)};
        run_tile_(q, plan, false, false, 0, depth - 1, depth);
        Result r = finish_result_(rows, stride);
        for (auto ev : evs) cudaEventDestroy(ev);
        return r;
    }
    // Host buffers: the batch is cut into sub-tiles whose uploads (copy stream, two staging sets) overlap the scoring of
    // the previous sub-tile; results stay on the device until the last sub-tile is done.
    uint32_t sub = tile;
    if (pipeline_uploads_ && rows >= 4096u) {
        const uint32_t part = ((rows + pipeline_parts_ - 1u) / pipeline_parts_ + 31u) & ~31u;
        sub = std::min(tile, std::max<uint32_t>(1024u, part));
    }
    uint64_t max_nnz = 0;
    for (uint32_t r0 = 0; r0 < rows; r0 += sub) max_nnz = std::max(max_nnz, row_ptr[r0 + std::min(sub, rows - r0)] - row_ptr[r0]);
    DeviceBuffer<uint64_t>* s_rp[2] = {&x_row_ptr_, &x2_row_ptr_};
    DeviceBuffer<uint32_t>* s_ci[2] = {&x_col_idx_, &x2_col_idx_};
    DeviceBuffer<float>* s_va[2] = {&x_val_, &x2_val_};
    const bool two_sets = rows > sub;
    for (int b = 0; b < (two_sets ? 2 : 1); ++b) {  // no (synchronising) reallocation inside the pipeline
        s_rp[b]->reserve(static_cast<uint64_t>(sub) + 1);
        s_ci[b]->reserve(max_nnz);
        s_va[b]->reserve(max_nnz);
    }
    uint32_t t = 0;
    for (uint32_t r0 = 0; r0 < rows; r0 += sub, ++t) {
        const uint32_t tr = std::min(sub, rows - r0);
        const uint64_t base = row_ptr[r0], end = row_ptr[r0 + tr];
        const int b = two_sets ? static_cast<int>(t & 1u) : 0;
        cudaStream_t up = two_sets ? copy_stream_ : stream_;
        if (two_sets && t >= 2) PB200_CUDA(cudaStreamWaitEvent(copy_stream_, use_ev_[b], 0));
        s_rp[b]->upload(row_ptr + r0, static_cast<uint64_t>(tr) + 1, up);
        s_ci[b]->upload(col_idx + base, end - base, up);
        s_va[b]->upload(val + base, end - base, up);
        if (two_sets) {
            PB200_CUDA(cudaEventRecord(up_ev_[b], copy_stream_));
            PB200_CUDA(cudaStreamWaitEvent(stream_, up_ev_[b], 0));
        }
        QueryDev q{s_rp[b]->get(), s_ci[b]->get(), s_va[b]->get(), base, tr, cols, max_row_nnz(row_ptr + r0, tr)};
        res_rows_ = r0;
        run_tile_(q, plan, false);
        if (two_sets) PB200_CUDA(cudaEventRecord(use_ev_[b], stream_));
    }
    return finish_result_(rows, stride);
}

XLinearEngine::Result XLinearEngine::predict_drm(const float* dense, uint32_t rows, uint32_t cols, uint32_t beam_size,
                                                 const char* post_processor, uint32_t only_topk) {
    PB200_CUDA(cudaSetDevice(device_));
    const auto plan = make_plan_(beam_size, post_processor, only_topk);
    uint32_t tile = pick_tile_rows_(plan, rows);
    const uint64_t max_dense_rows = std::max<uint64_t>(1, (4ull << 30) / (static_cast<uint64_t>(std::max<uint32_t>(cols, 1u)) * 4));
    tile = static_cast<uint32_t>(std::min<uint64_t>(tile, max_dense_rows));
    ensure_workspace_(plan, tile);
    const uint32_t stride = plan.back().k_cap;
    res_stride_ = stride;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    for (uint32_t r0 = 0; r0 < rows; r0 += tile) {
        const uint32_t tr = std::min(tile, rows - r0);
        x_val_.upload(dense + static_cast<uint64_t>(r0) * cols, static_cast<uint64_t>(tr) * cols, stream_);
        QueryDev q{nullptr, nullptr, x_val_.get(), 0, tr, cols, cols};
        res_rows_ = r0;
        run_tile_(q, plan, false);
        if (r0 + tr < rows) PB200_CUDA(cudaStreamSynchronize(stream_));
    }
    return finish_result_(rows, stride);
}

XLinearEngine::Result XLinearEngine::predict_single_layer(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val,
                                                          const float* dense, uint32_t rows, uint32_t cols,
                                                          const uint64_t* codes_row_ptr, const uint32_t* codes_col_idx,
                                                          const float* codes_val, const char* post_processor,
                                                          uint32_t only_topk) {
    PB200_CUDA(cudaSetDevice(device_));
    if (host_->layers.size() != 1) throw std::runtime_error("pecos_b200: predict_single_layer needs a one-layer model");
    const auto& HL = host_->layers[0];
    const bool have_codes = codes_row_ptr != nullptr;
    // beam entering the layer: the given previous prediction, or every parent with value 1 (libpecos.cpp:209-219)
    uint32_t b_prev = have_codes ? 1u : std::max<uint32_t>(HL.n_chunks, 1u);
    if (have_codes) {
        for (uint32_t r = 0; r < rows; ++r)
            b_prev = std::max<uint32_t>(b_prev, static_cast<uint32_t>(std::min<uint64_t>(codes_row_ptr[r + 1] - codes_row_ptr[r], 0xFFFFFFFFull)));
        const uint64_t n = codes_row_ptr[rows] - codes_row_ptr[0];
        for (uint64_t i = 0; i < n; ++i)
            if (codes_col_idx[codes_row_ptr[0] + i] >= HL.n_chunks)
                throw std::runtime_error("pecos_b200: csr_codes column index >= C.cols");
    }
    std::vector<LayerPlan> plan(1);
    plan[0].k = only_topk;  // only_topk_to_use = overridden > 0 ? overridden : metadata.only_topk, both = this argument
    plan[0].pp = parse_post_processor(post_processor ? post_processor : HL.post_processor_name.c_str());
    plan[0].b_prev = b_prev;
    const uint64_t cand_max = static_cast<uint64_t>(b_prev) * std::max<uint32_t>(HL.c_max, 1u);
    plan[0].k_cap = static_cast<uint32_t>(std::max<uint64_t>(1, std::min<uint64_t>(only_topk, cand_max)));
    if (topk_kernel_smem(b_prev) > 200u * 1024u || b_prev > 32768u)
        throw std::runtime_error("pecos_b200: beam of " + std::to_string(b_prev) + " nodes exceeds the supported maximum");
    const uint32_t stride = plan[0].k_cap;
    res_stride_ = stride;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    if (only_topk == 0) {  // sorted_csr keeps min(nnz, 0) entries per row (inference.hpp:1237)
        if (rows) PB200_CUDA(cudaMemsetAsync(res_cnt_dev_.get(), 0, static_cast<uint64_t>(rows) * 4, stream_));
        return finish_result_(rows, stride);
    }
    uint32_t tile = pick_tile_rows_(plan, rows);
    if (dense) {
        const uint64_t max_dense_rows = std::max<uint64_t>(1, (4ull << 30) / (static_cast<uint64_t>(std::max<uint32_t>(cols, 1u)) * 4));
        tile = static_cast<uint32_t>(std::min<uint64_t>(tile, max_dense_rows));
    }
    ensure_workspace_(plan, tile);
    beam_id_host_.reserve(static_cast<uint64_t>(tile) * beam_stride_ + 1);
    beam_val_host_.reserve(static_cast<uint64_t>(tile) * beam_stride_ + 1);
    beam_cnt_host_.reserve(static_cast<uint64_t>(tile) + 1);
    for (uint32_t r0 = 0; r0 < rows; r0 += tile) {
        const uint32_t tr = std::min(tile, rows - r0);
        for (uint32_t r = 0; r < tr; ++r) {
            uint32_t* ids = beam_id_host_.get() + static_cast<uint64_t>(r) * beam_stride_;
            float* vals = beam_val_host_.get() + static_cast<uint64_t>(r) * beam_stride_;
            if (have_codes) {
                const uint64_t b = codes_row_ptr[r0 + r], e = codes_row_ptr[r0 + r + 1];
                const uint32_t n = static_cast<uint32_t>(e - b);
                std::memcpy(ids, codes_col_idx + b, static_cast<size_t>(n) * 4);
                std::memcpy(vals, codes_val + b, static_cast<size_t>(n) * 4);
                beam_cnt_host_.get()[r] = n;
            } else {
                for (uint32_t j = 0; j < HL.n_chunks; ++j) { ids[j] = j; vals[j] = 1.0f; }
                beam_cnt_host_.get()[r] = HL.n_chunks;
            }
        }
        PB200_CUDA(cudaMemcpyAsync(beam_id_[0].get(), beam_id_host_.get(), static_cast<uint64_t>(tr) * beam_stride_ * 4, cudaMemcpyHostToDevice, stream_));
        PB200_CUDA(cudaMemcpyAsync(beam_val_[0].get(), beam_val_host_.get(), static_cast<uint64_t>(tr) * beam_stride_ * 4, cudaMemcpyHostToDevice, stream_));
        PB200_CUDA(cudaMemcpyAsync(beam_cnt_[0].get(), beam_cnt_host_.get(), static_cast<uint64_t>(tr) * 4, cudaMemcpyHostToDevice, stream_));
        QueryDev q{};
        if (dense) {
            x_val_.upload(dense + static_cast<uint64_t>(r0) * cols, static_cast<uint64_t>(tr) * cols, stream_);
            q = QueryDev{nullptr, nullptr, x_val_.get(), 0, tr, cols, cols};
        } else {
            const uint64_t base = row_ptr[r0], end = row_ptr[r0 + tr];
            x_row_ptr_.upload(row_ptr + r0, static_cast<uint64_t>(tr) + 1, stream_);
            x_col_idx_.upload(col_idx + base, end - base, stream_);
            x_val_.upload(val + base, end - base, stream_);
            q = QueryDev{x_row_ptr_.get(), x_col_idx_.get(), x_val_.get(), base, tr, cols, max_row_nnz(row_ptr + r0, tr)};
        }
        res_rows_ = r0;
        run_tile_(q, plan, false, /*ext_beam=*/true, /*combine_first=*/have_codes ? 1 : 0);
        PB200_CUDA(cudaStreamSynchronize(stream_));  // the pinned beam staging area is refilled by the next tile
    }
    return finish_result_(rows, stride);
}

void XLinearEngine::resident_upload_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows,
                                        uint32_t cols) {
    PB200_CUDA(cudaSetDevice(device_));
    const uint64_t nnz = row_ptr[rows];
    x_row_ptr_.upload(row_ptr, static_cast<uint64_t>(rows) + 1, stream_);
    x_col_idx_.upload(col_idx, nnz, stream_);
    x_val_.upload(val, nnz, stream_);
    PB200_CUDA(cudaStreamSynchronize(stream_));
    resident_ = QueryDev{x_row_ptr_.get(), x_col_idx_.get(), x_val_.get(), 0, rows, cols, max_row_nnz(row_ptr, rows)};
    has_resident_ = true;
}

double XLinearEngine::resident_predict(uint32_t beam_size, const char* post_processor, uint32_t only_topk, bool collect_stats) {
    if (!has_resident_) throw std::runtime_error("pecos_b200: no resident query batch uploaded");
    PB200_CUDA(cudaSetDevice(device_));
    const auto plan = make_plan_(beam_size, post_processor, only_topk);
    const uint32_t rows = resident_.rows;
    const uint32_t tile = pick_tile_rows_(plan, rows);
    ensure_workspace_(plan, tile);
    const uint32_t stride = plan.back().k_cap;
    res_stride_ = stride;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * stride + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    if (collect_stats) PB200_CUDA(cudaMemsetAsync(stats_dev_.get(), 0, stats_dev_.bytes(), stream_));
    PB200_CUDA(cudaEventRecord(ev_[3], stream_));
    cudaEvent_t stop;
    PB200_CUDA(cudaEventCreate(&stop));
    for (uint32_t r0 = 0; r0 < rows; r0 += tile) {
        const uint32_t tr = std::min(tile, rows - r0);
        QueryDev q = resident_;
        q.row_ptr = resident_.row_ptr + r0;
        q.rows = tr;
        res_rows_ = r0;
        run_tile_(q, plan, collect_stats);
    }
    PB200_CUDA(cudaEventRecord(stop, stream_));
    PB200_CUDA(cudaEventSynchronize(stop));
    float ms = 0.f;
    PB200_CUDA(cudaEventElapsedTime(&ms, ev_[3], stop));
    cudaEventDestroy(stop);
    res_rows_ = rows;
    if (collect_stats) {
        std::vector<unsigned long long> h(8 * layers_.size());
        PB200_CUDA(cudaMemcpy(h.data(), stats_dev_.get(), h.size() * 8, cudaMemcpyDeviceToHost));
        for (size_t d = 0; d < layers_.size(); ++d) {
            XLinearStats s{};
            s.chunks = h[8 * d + 0]; s.chunk_rows = h[8 * d + 1]; s.matched = h[8 * d + 2]; s.entries = h[8 * d + 3];
            s.out_cols = h[8 * d + 4]; s.query_nnz = h[8 * d + 5]; s.beam_out = h[8 * d + 6];
            layer_stats_[d] = s;
        }
    }
    return static_cast<double>(ms);
}

uint32_t XLinearEngine::sharded_local_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows,
                                          uint32_t cols, uint32_t beam_size, const char* post_processor, uint32_t only_topk,
                                          uint32_t stride_capacity, unsigned long long* keys_dev, uint32_t* ids_dev,
                                          float* vals_dev, uint32_t* cnt_dev) {
    PB200_CUDA(cudaSetDevice(device_));
    const auto plan = make_plan_(beam_size, post_processor, only_topk);
    const uint32_t stride = plan.back().k_cap;
    if (stride > stride_capacity) throw std::runtime_error("pecos_b200: sharded output buffers are too narrow for this top-k");
    const uint32_t tile = pick_tile_rows_(plan, rows);
    ensure_workspace_(plan, tile);
    res_stride_ = stride;
    ext_keys_ = keys_dev; ext_ids_ = ids_dev; ext_vals_ = vals_dev; ext_cnt_ = cnt_dev;
    try {
        for (uint32_t r0 = 0; r0 < rows; r0 += tile) {
            const uint32_t tr = std::min(tile, rows - r0);
            const uint64_t base = row_ptr[r0], end = row_ptr[r0 + tr];
            x_row_ptr_.upload(row_ptr + r0, static_cast<uint64_t>(tr) + 1, stream_);
            x_col_idx_.upload(col_idx + base, end - base, stream_);
            x_val_.upload(val + base, end - base, stream_);
            QueryDev q{x_row_ptr_.get(), x_col_idx_.get(), x_val_.get(), base, tr, cols, max_row_nnz(row_ptr + r0, tr)};
            res_rows_ = r0;
            run_tile_(q, plan, false);
            PB200_CUDA(cudaStreamSynchronize(stream_));
        }
    } catch (...) {
        ext_keys_ = nullptr; ext_ids_ = nullptr; ext_vals_ = nullptr; ext_cnt_ = nullptr;
        throw;
    }
    ext_keys_ = nullptr; ext_ids_ = nullptr; ext_vals_ = nullptr; ext_cnt_ = nullptr;
    return stride;
}

XLinearEngine::Result XLinearEngine::sharded_merge(uint32_t world, uint32_t rows, uint32_t stride, uint32_t only_topk,
                                                   const unsigned long long* g_keys, const uint32_t* g_ids,
                                                   const float* g_vals, const uint32_t* g_cnt) {
    PB200_CUDA(cudaSetDevice(device_));
    if (static_cast<uint64_t>(world) * stride > static_cast<uint64_t>(kSelKeys))
        throw std::runtime_error("pecos_b200: world * top-k exceeds the merge kernel's capacity");
    const uint32_t k = only_topk ? only_topk : static_cast<uint32_t>(host_->layers.back().only_topk);
    const uint32_t k_out = std::min<uint32_t>(k, world * stride);
    res_stride_ = k_out;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * k_out + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * k_out + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    if (rows) {
        xl_merge_topk_kernel<<<(rows + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream_>>>(
            g_keys, g_ids, g_vals, g_cnt, world, rows, stride, k_out, res_ids_dev_.get(), res_vals_dev_.get(), res_cnt_dev_.get());
        PB200_CUDA(cudaGetLastError());
        ++launches_;
    }
    return finish_result_(rows, k_out);
}

uint32_t XLinearEngine::sharded_local_csr_packed(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows,
                                                 uint32_t cols, uint32_t beam_size, const char* post_processor, uint32_t only_topk,
                                                 uint32_t stride_capacity, void* rec_dev) {
    PB200_CUDA(cudaSetDevice(device_));
    const uint64_t n = static_cast<uint64_t>(rows) * std::max<uint32_t>(stride_capacity, 1u);
    shard_keys_.reserve(n + 1);
    shard_ids_.reserve(n + 1);
    shard_vals_.reserve(n + 1);
    shard_cnt_.reserve(static_cast<uint64_t>(rows) + 1);
    const uint32_t stride = sharded_local_csr(row_ptr, col_idx, val, rows, cols, beam_size, post_processor, only_topk, stride_capacity,
                                              shard_keys_.get(), shard_ids_.get(), shard_vals_.get(), shard_cnt_.get());
    const uint64_t total = static_cast<uint64_t>(rows) * stride;
    if (total) {
        xl_shard_pack_kernel<<<static_cast<uint32_t>((total + 255) / 256), 256, 0, stream_>>>(
            shard_keys_.get(), shard_ids_.get(), shard_vals_.get(), shard_cnt_.get(), rows, stride, static_cast<ShardRecord*>(rec_dev));
        PB200_CUDA(cudaGetLastError());
        ++launches_;
    }
    PB200_CUDA(cudaStreamSynchronize(stream_));
    return stride;
}

XLinearEngine::Result XLinearEngine::sharded_merge_packed(uint32_t world, uint32_t rows, uint32_t stride, uint32_t only_topk,
                                                          const void* g_rec) {
    PB200_CUDA(cudaSetDevice(device_));
    if (static_cast<uint64_t>(world) * stride > static_cast<uint64_t>(kSelKeys))
        throw std::runtime_error("pecos_b200: world * top-k exceeds the merge kernel's capacity");
    const uint32_t k = only_topk ? only_topk : static_cast<uint32_t>(host_->layers.back().only_topk);
    const uint32_t k_out = std::min<uint32_t>(k, world * stride);
    res_stride_ = k_out;
    res_ids_dev_.reserve(static_cast<uint64_t>(rows) * k_out + 1);
    res_vals_dev_.reserve(static_cast<uint64_t>(rows) * k_out + 1);
    res_cnt_dev_.reserve(static_cast<uint64_t>(rows) + 1);
    if (rows) {
        xl_merge_packed_kernel<<<(rows + kSelWarps - 1) / kSelWarps, kSelWarps * 32, 0, stream_>>>(
            static_cast<const ShardRecord*>(g_rec), world, rows, stride, k_out, res_ids_dev_.get(), res_vals_dev_.get(), res_cnt_dev_.get());
        PB200_CUDA(cudaGetLastError());
        ++launches_;
    }
    return finish_result_(rows, k_out);
}

XLinearEngine::Result XLinearEngine::resident_fetch() {
    PB200_CUDA(cudaSetDevice(device_));
    return finish_result_(resident_.rows, res_stride_);
}

#define PB200_SELECTED_ENGINE
#include "xlinear_selected.cuh"
#undef PB200_SELECTED_ENGINE

}  // namespace pb200
