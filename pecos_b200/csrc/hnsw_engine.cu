// HNSW dense search on B200 (sm_100a): one warp walks one query (see hnsw_engine.h for the reference map).
//
// Per expansion of the best-first search the warp
//   A. reads the node's neighbour list, test-and-sets the per-warp visited bitmap, compacts the not-yet-visited ids in
//      list order (ballot + popc),
//   B. evaluates ALL their distances, two at a time (one per half-warp).  A half-warp's 16 lanes read one float4 each
//      per 64 components from the permuted row (256 contiguous bytes) and lane j runs exactly the accumulation chain of
//      SIMD lane j of the reference's avx512 kernel: un-fused multiply/add, fold 16 -> 4 -> 1 in the same association,
//      4-wide tail un-fused, scalar tail fused (as GCC compiles that clone) -- distances are bit-identical,
//   C. lane 0 replays the reference's queue updates sequentially in neighbour order with restated libstdc++
//      push_heap / pop_heap, so ties and the evolving upper bound behave exactly like the CPU code.
// The reference evaluates the distance of every unvisited neighbour before touching the queues (hnsw.hpp:897-914), which
// is what makes step B independent of step C.
//
// Sparse (csr) indices (FeatVecSparse{IP,L2}Simd, feat_vectors.hpp:186-210) use the same walk; only step B differs: a half-warp
// streams the neighbour's {index, value} entries (16 per step, 128 contiguous bytes), every lane looks its entry up in the query
// row staged in shared memory (a 8,192-bit filter first, binary search on a filter hit) and the matched products are added in
// ascending index order -- the order of the reference's block intersection (distance_impl/common.hpp:15-86) for rows with strictly
// ascending indices.  The reference's sparse "l2" is -2<x,y> (its squared norms are do_l2_distance_simd(x, x) = 0): restated as is.
//
// HBM traffic per query: n_dist * 4d + n_expand * 4(1+maxM0) + hops * 4(1+maxM) + 4d + 8k.
#include "hnsw_engine.h"

#include <algorithm>
#include <atomic>
#include <cstring>

namespace pb200 {

namespace {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kEfSmemMax = 512;  // result heaps up to this many entries live in shared memory

__device__ __forceinline__ float4 ld_stream_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

template <int METRIC>
__device__ __forceinline__ float chain_step(float acc, float x, float y) {
    if (METRIC == HNSW_IP) return __fadd_rn(acc, __fmul_rn(x, y));
    const float d = __fsub_rn(x, y);
    return __fadd_rn(acc, __fmul_rn(d, d));
}

// Distance between the staged (permuted) query and base vector `node`, computed by one half-warp.
// Every lane of the warp must call it; the result is valid on lanes 0 and 16 (hl == 0).
template <int METRIC, bool SMEM>
__device__ __forceinline__ float half_warp_distance(const HnswDev& ix, const float* qs, const float* row, int hl) {
    const float4* v4 = reinterpret_cast<const float4*>(row);
    const float4* q4 = reinterpret_cast<const float4*>(qs);
    const uint32_t nK = ix.main_pad >> 6;
    float acc = 0.0f;
    uint32_t K = 0;
    for (; K + 4 <= nK; K += 4) {
        const float4 y0 = SMEM ? v4[(K + 0) * 16 + hl] : ld_stream_f4(v4 + (K + 0) * 16 + hl);
        const float4 y1 = SMEM ? v4[(K + 1) * 16 + hl] : ld_stream_f4(v4 + (K + 1) * 16 + hl);
        const float4 y2 = SMEM ? v4[(K + 2) * 16 + hl] : ld_stream_f4(v4 + (K + 2) * 16 + hl);
        const float4 y3 = SMEM ? v4[(K + 3) * 16 + hl] : ld_stream_f4(v4 + (K + 3) * 16 + hl);
        const float4 x0 = q4[(K + 0) * 16 + hl], x1 = q4[(K + 1) * 16 + hl], x2 = q4[(K + 2) * 16 + hl], x3 = q4[(K + 3) * 16 + hl];
        acc = chain_step<METRIC>(acc, x0.x, y0.x); acc = chain_step<METRIC>(acc, x0.y, y0.y);
        acc = chain_step<METRIC>(acc, x0.z, y0.z); acc = chain_step<METRIC>(acc, x0.w, y0.w);
        acc = chain_step<METRIC>(acc, x1.x, y1.x); acc = chain_step<METRIC>(acc, x1.y, y1.y);
        acc = chain_step<METRIC>(acc, x1.z, y1.z); acc = chain_step<METRIC>(acc, x1.w, y1.w);
        acc = chain_step<METRIC>(acc, x2.x, y2.x); acc = chain_step<METRIC>(acc, x2.y, y2.y);
        acc = chain_step<METRIC>(acc, x2.z, y2.z); acc = chain_step<METRIC>(acc, x2.w, y2.w);
        acc = chain_step<METRIC>(acc, x3.x, y3.x); acc = chain_step<METRIC>(acc, x3.y, y3.y);
        acc = chain_step<METRIC>(acc, x3.z, y3.z); acc = chain_step<METRIC>(acc, x3.w, y3.w);
    }
    for (; K < nK; ++K) {
        const float4 y = SMEM ? v4[K * 16 + hl] : ld_stream_f4(v4 + K * 16 + hl);
        const float4 x = q4[K * 16 + hl];
        acc = chain_step<METRIC>(acc, x.x, y.x); acc = chain_step<METRIC>(acc, x.y, y.y);
        acc = chain_step<METRIC>(acc, x.z, y.z); acc = chain_step<METRIC>(acc, x.w, y.w);
    }
    // fold 16 partial sums -> 4:  (a[j] + a[4+j]) + (a[8+j] + a[12+j])   (x86.hpp:138-141)
    const float u = __fadd_rn(acc, __shfl_down_sync(kFull, acc, 4, 16));
    float s = __fadd_rn(u, __shfl_down_sync(kFull, u, 8, 16));
    const uint32_t tl = ix.tail_len;
    const float* yt = row + ix.main_pad;
    const float* xt = qs + ix.main_pad;
    const uint32_t g4 = tl >> 2;
    for (uint32_t g = 0; g < g4; ++g) {  // 4-wide remainder loop (x86.hpp:143-147), lanes 0..3 of the half-warp
        if (hl < 4) s = chain_step<METRIC>(s, xt[g * 4 + hl], yt[g * 4 + hl]);
    }
    const float s1 = __shfl_down_sync(kFull, s, 1, 16);
    const float s2 = __shfl_down_sync(kFull, s, 2, 16);
    const float s3 = __shfl_down_sync(kFull, s, 3, 16);
    float sum = __fadd_rn(__fadd_rn(__fadd_rn(s, s1), s2), s3);  // tmp_sum[0] + tmp_sum[1] + tmp_sum[2] + tmp_sum[3]
    for (uint32_t i = g4 * 4; i < tl; ++i) {  // scalar tail: fused multiply-add in the avx512f clone
        if (METRIC == HNSW_IP) sum = __fmaf_rn(xt[i], yt[i], sum);
        else { const float d = __fsub_rn(xt[i], yt[i]); sum = __fmaf_rn(d, d, sum); }
    }
    if (METRIC == HNSW_IP) return static_cast<float>(1.0 - static_cast<double>(sum));  // feat_vectors.hpp:138-141
    return sum;
}

// ---- bulk asynchronous copies (TMA engine, non-tensor form: SASS UBLKCP) + mbarrier completion -----------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// one elected lane: expect `bytes` on the barrier, then start the copy global -> shared
__device__ __forceinline__ void bulk_load_row(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    uint32_t done = 0;
    do {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done) : "r"(mbar), "r"(parity) : "memory");
    } while (!done);
}

// distances of ids[0..n) -> dist[0..n), two per step (one per half-warp).
// STAGES == 0: each lane loads its float4s straight from HBM.  STAGES > 0: base vectors are brought into a per-warp ring
// of STAGES shared-memory slots by bulk asynchronous copies (one 16-byte-aligned contiguous row each), STAGES rows in
// flight per warp, so the HBM latency of the next rows hides behind the arithmetic of the current pair.
template <int METRIC, int STAGES>
__device__ __forceinline__ void batch_distances(const HnswDev& ix, const float* qs, const uint32_t* ids, float* dist,
                                                uint32_t n, int lane, float* ring, uint32_t mbar0, uint32_t& phase_bits) {
    const int half = lane >> 4, hl = lane & 15;
    if (STAGES == 0) {
        for (uint32_t b = 0; b < n; b += 2) {
            const uint32_t slot = b + half;
            const uint32_t node = ids[min(slot, n - 1)];
            const float d = half_warp_distance<METRIC, false>(ix, qs, ix.vec + static_cast<uint64_t>(node) * ix.vstride, hl);
            if (hl == 0 && slot < n) dist[slot] = d;
        }
        __syncwarp();
        return;
    }
    const uint32_t bytes = ix.vstride * 4u;
    const uint32_t ring0 = smem_addr(ring);
    if (lane == 0) {
        const uint32_t first = min(static_cast<uint32_t>(STAGES), n);
        for (uint32_t s = 0; s < first; ++s)
            bulk_load_row(ring0 + s * bytes, ix.vec + static_cast<uint64_t>(ids[s]) * ix.vstride, bytes, mbar0 + 8u * s);
    }
    for (uint32_t b = 0; b < n; b += 2) {
        constexpr uint32_t kRing = STAGES > 0 ? STAGES : 1;  // (STAGES == 0 never reaches this path)
        const uint32_t slot0 = b % kRing, slot1 = (b + 1) % kRing;
        const bool second = (b + 1) < n;
        const uint32_t my = (half && second) ? slot1 : slot0;
        mbar_wait(mbar0 + 8u * my, (phase_bits >> my) & 1u);
        const float d = half_warp_distance<METRIC, true>(ix, qs, ring + static_cast<size_t>(my) * ix.vstride, hl);
        if (hl == 0 && (b + half) < n) dist[b + half] = d;
        phase_bits ^= (1u << slot0) | (second ? (1u << slot1) : 0u);
        __syncwarp();  // both slots fully consumed before they are refilled
        if (lane == 0) {
            const uint32_t nxt = b + STAGES;
            if (nxt < n) bulk_load_row(ring0 + slot0 * bytes, ix.vec + static_cast<uint64_t>(ids[nxt]) * ix.vstride, bytes, mbar0 + 8u * slot0);
            if (nxt + 1 < n) bulk_load_row(ring0 + slot1 * bytes, ix.vec + static_cast<uint64_t>(ids[nxt + 1]) * ix.vstride, bytes, mbar0 + 8u * slot1);
        }
    }
    __syncwarp();
}

// ---- sparse rows: ordered intersection ---------------------------------------------------------------------------------
constexpr uint32_t kSpFilterWords = 256;  // 8,192-bit membership filter of the query row's indices, per warp
constexpr uint32_t kSpQcapMax = 4096;     // query entries staged per warp at most (longer rows are searched in global memory)

__device__ __forceinline__ uint32_t sp_hash(uint32_t idx) { return (idx * 2654435761u) >> 19; }  // 13 bits

struct SparseQuery {  // one query row: generic pointers (shared-memory copy, or the global arrays for very long rows)
    const uint32_t* idx;
    const float* val;
    uint32_t n;
    const uint32_t* filter;
};

__device__ __forceinline__ uint2 ld_stream_u2(const uint2* p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// one 16-entry step of a half-warp: look the lane's entry up, then add the matched products of this step in entry order
__device__ __forceinline__ float sparse_step(const SparseQuery& q, bool has, uint2 ent, float ret, int lane) {
    bool hit = false;
    float prod = 0.0f;
    if (has) {
        const uint32_t h = sp_hash(ent.x);
        if ((q.filter[h >> 5] >> (h & 31u)) & 1u) {
            uint32_t lo = 0, hi = q.n;  // std::lower_bound
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (q.idx[mid] < ent.x) lo = mid + 1; else hi = mid;
            }
            if (lo < q.n && q.idx[lo] == ent.x) { hit = true; prod = __fmul_rn(q.val[lo], __uint_as_float(ent.y)); }
        }
    }
    const unsigned m = __ballot_sync(kFull, hit);
    if (m == 0u) return ret;
    const int half = lane >> 4;
    unsigned mh = (m >> (16 * half)) & 0xFFFFu;
    const int n_it = max(__popc(m & 0xFFFFu), __popc(m >> 16));
    for (int it = 0; it < n_it; ++it) {
        const int src = mh ? (__ffs(mh) - 1 + 16 * half) : lane;
        const float pv = __shfl_sync(kFull, prod, src);
        if (mh) { ret = __fadd_rn(ret, pv); mh &= mh - 1u; }
    }
    return ret;
}

// distances of ids[0..n) -> dist[0..n) for a sparse index, two rows at a time (one per half-warp)
template <int METRIC>
__device__ __forceinline__ void batch_distances_sparse(const HnswDev& ix, const SparseQuery& q, const uint32_t* ids, float* dist,
                                                       uint32_t n, int lane, unsigned long long& n_entries) {
    const int half = lane >> 4, hl = lane & 15;
    for (uint32_t b = 0; b < n; b += 2) {
        const uint32_t slot = b + half;
        const bool valid = slot < n;
        unsigned long long r0 = 0, r1 = 0;
        if (valid) {
            const uint32_t node = ids[slot];
            r0 = ix.sp_ptr[node];
            r1 = ix.sp_ptr[node + 1];
        }
        const uint32_t len = static_cast<uint32_t>(r1 - r0);
        const uint32_t len_max = max(len, __shfl_xor_sync(kFull, len, 16));
        if (hl == 0) n_entries += len;
        const uint2* row = ix.sp_ent + r0;
        float ret = 0.0f;
        for (uint32_t j0 = 0; j0 < len_max; j0 += 64) {  // four independent 128-byte loads per half-warp in flight
            uint2 e[4];
            bool has[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t j = j0 + 16u * u + hl;
                has[u] = (j < len) && q.n != 0u;
                e[u] = has[u] ? ld_stream_u2(row + j) : make_uint2(0u, 0u);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (j0 + 16u * u < len_max) ret = sparse_step(q, has[u], e[u], ret, lane);
            }
        }
        if (hl == 0 && valid) {
            // FeatVecSparseIPSimd: 1.0 - dot ; FeatVecSparseL2Simd: x_sq + y_sq - 2.0 * dot with x_sq = y_sq = 0 (see the header)
            dist[slot] = (METRIC == HNSW_IP) ? static_cast<float>(1.0 - static_cast<double>(ret))
                                             : static_cast<float>(static_cast<double>(0.0f) - 2.0 * static_cast<double>(ret));
        }
    }
    __syncwarp();
}

// ---- libstdc++ heap algorithms (std::push_heap / std::pop_heap), entries {dist bits, node}; MAXH: std::less ---------
template <bool MAXH>
__device__ __forceinline__ bool heap_comp(uint2 a, float value_dist) {
    const float ad = __uint_as_float(a.x);
    return MAXH ? (ad < value_dist) : (ad > value_dist);
}

template <bool MAXH>
__device__ __forceinline__ void heap_sift_up(uint2* h, int hole, int top, uint2 value) {
    const float vd = __uint_as_float(value.x);
    int parent = (hole - 1) / 2;
    while (hole > top && heap_comp<MAXH>(h[parent], vd)) {
        h[hole] = h[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    h[hole] = value;
}

template <bool MAXH>
__device__ __forceinline__ void heap_adjust(uint2* h, int hole, int len, uint2 value) {
    const int top = hole;
    int child = hole;
    while (child < (len - 1) / 2) {
        child = 2 * (child + 1);
        if (heap_comp<MAXH>(h[child], __uint_as_float(h[child - 1].x))) child--;
        h[hole] = h[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2) {
        child = 2 * (child + 1);
        h[hole] = h[child - 1];
        hole = child - 1;
    }
    heap_sift_up<MAXH>(h, hole, top, value);
}

template <bool MAXH>
__device__ __forceinline__ void heap_push(uint2* h, int& n, float dist, uint32_t node) {
    const uint2 v = make_uint2(__float_as_uint(dist), node);
    h[n] = v;
    ++n;
    heap_sift_up<MAXH>(h, n - 1, 0, v);
}

template <bool MAXH>
__device__ __forceinline__ void heap_pop(uint2* h, int& n) {
    if (n > 1) {
        const uint2 value = h[n - 1];
        h[n - 1] = h[0];
        heap_adjust<MAXH>(h, 0, n - 1, value);
    }
    --n;
}

__device__ __forceinline__ uint32_t permuted_pos_dev(const HnswDev& ix, uint32_t i) {
    const uint32_t m = (ix.feat_dim >> 4) << 4;
    if (i < m) {
        const uint32_t k = i >> 4, j = i & 15u;
        return 64u * (k >> 2) + 4u * j + (k & 3u);
    }
    return ix.main_pad + (i - m);
}

template <int METRIC, int STAGES, bool SPARSE>
__global__ void __launch_bounds__(256)
hnsw_search_kernel(const HnswDev ix, const float* __restrict__ Q, const HnswSparseQueries SQ, const uint32_t nq, const uint32_t efS, const uint32_t topk,
                   const uint32_t ef, uint32_t* __restrict__ out_idx, float* __restrict__ out_val, uint32_t* bitmap_all,
                   const uint32_t bitmap_words, uint32_t* vlist_all, uint2* cand_all, const uint32_t vcap, uint2* topk_all,
                   const uint32_t nbmax, const uint32_t per_warp_bytes, unsigned long long* ctrl) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t gw = blockIdx.x * (blockDim.x >> 5) + warp;
    unsigned char* base = smem_raw + static_cast<size_t>(warp) * per_warp_bytes;
    // per-warp slice: [query | STAGES ring slots | STAGES mbarriers | neighbour ids | distances | result heap]
    // (sparse: [query indices qcap | query values qcap | filter | neighbour ids | distances | result heap])
    float* qs = reinterpret_cast<float*>(base);
    float* ring = qs + ix.vstride;
    unsigned long long* mbars = reinterpret_cast<unsigned long long*>(ring + static_cast<size_t>(STAGES) * ix.vstride);
    uint32_t* sq_idx = reinterpret_cast<uint32_t*>(base);
    float* sq_val = reinterpret_cast<float*>(sq_idx + SQ.qcap);
    uint32_t* sq_filter = reinterpret_cast<uint32_t*>(sq_val + SQ.qcap);
    uint32_t* nb_ids = SPARSE ? sq_filter + kSpFilterWords : reinterpret_cast<uint32_t*>(mbars + STAGES);
    float* nb_dist = reinterpret_cast<float*>(nb_ids + nbmax);
    uint2* topq = topk_all ? topk_all + static_cast<uint64_t>(gw) * (ef + 1) : reinterpret_cast<uint2*>(nb_dist + nbmax);
    const uint32_t mbar0 = smem_addr(mbars);
    uint32_t phase_bits = 0;
    if (STAGES > 0 && !SPARSE) {
        if (lane == 0) {
            for (int s = 0; s < STAGES; ++s) mbar_init(mbar0 + 8u * s, 1u);
            fence_proxy_async_smem();
        }
        __syncwarp();
    }
    uint32_t* bitmap = bitmap_all + static_cast<uint64_t>(gw) * bitmap_words;
    uint32_t* vlist = vlist_all + static_cast<uint64_t>(gw) * vcap;
    uint2* cand = cand_all + static_cast<uint64_t>(gw) * vcap;
    const uint32_t d = ix.feat_dim;

    for (;;) {
        unsigned long long qq = 0;
        if (lane == 0) qq = atomicAdd(&ctrl[0], 1ull);
        qq = __shfl_sync(kFull, qq, 0);
        if (qq >= nq) break;
        const uint32_t q = static_cast<uint32_t>(qq);
        unsigned long long n_dist = 0, n_expand = 0, n_hops = 0, n_entries = 0;

        SparseQuery sq{nullptr, nullptr, 0u, nullptr};
        if (SPARSE) {
            // stage the query row (indices, values) and the membership filter of its indices
            const unsigned long long q0 = SQ.ptr[q];
            sq.n = static_cast<uint32_t>(SQ.ptr[q + 1] - q0);
            for (uint32_t w = lane; w < kSpFilterWords; w += 32) sq_filter[w] = 0u;
            __syncwarp();
            const bool staged = sq.n <= SQ.qcap;
            for (uint32_t i = lane; i < sq.n; i += 32) {
                const uint32_t c = SQ.idx[q0 + i];
                if (staged) { sq_idx[i] = c; sq_val[i] = SQ.val[q0 + i]; }
                const uint32_t h = sp_hash(c);
                atomicOr(&sq_filter[h >> 5], 1u << (h & 31u));
            }
            sq.idx = staged ? sq_idx : SQ.idx + q0;
            sq.val = staged ? sq_val : SQ.val + q0;
            sq.filter = sq_filter;
            __syncwarp();
        } else {
            // stage the query in the permuted layout (padding = 0)
            for (uint32_t i = lane; i < ix.vstride; i += 32) qs[i] = 0.0f;
            __syncwarp();
            const float* qrow = Q + static_cast<uint64_t>(q) * d;
            for (uint32_t i = lane; i < d; i += 32) qs[permuted_pos_dev(ix, i)] = qrow[i];
            __syncwarp();
        }
        auto distances = [&](uint32_t n) {
            if (SPARSE) batch_distances_sparse<METRIC>(ix, sq, nb_ids, nb_dist, n, lane, n_entries);
            else batch_distances<METRIC, STAGES>(ix, qs, nb_ids, nb_dist, n, lane, ring, mbar0, phase_bits);
        };

        // ---- entry point + greedy descent on levels max_level..1 (hnsw.hpp:928-959)
        uint32_t curr = ix.init_node;
        if (lane == 0) nb_ids[0] = curr;
        __syncwarp();
        distances(1);
        float curr_dist = nb_dist[0];
        n_dist += 1;
        __syncwarp();
        for (uint32_t level = ix.max_level; level >= 1; --level) {
            int changed = 1;
            while (changed) {
                changed = 0;
                const uint32_t* nb = ix.l1 + static_cast<uint64_t>(curr) * ix.l1_node_mem + static_cast<uint64_t>(level - 1) * ix.l1_level_mem;
                const uint32_t deg = min(nb[0], ix.l1_max_degree);
                n_hops += 1;
                for (uint32_t j = lane; j < deg; j += 32) nb_ids[j] = nb[1 + j];
                __syncwarp();
                if (deg) distances(deg);
                n_dist += deg;
                if (lane == 0) {
                    for (uint32_t j = 0; j < deg; ++j) {
                        const float nd = nb_dist[j];
                        if (nd < curr_dist) { curr_dist = nd; curr = nb_ids[j]; changed = 1; }
                    }
                }
                curr = __shfl_sync(kFull, curr, 0);
                curr_dist = __shfl_sync(kFull, curr_dist, 0);
                changed = __shfl_sync(kFull, changed, 0);
                __syncwarp();
            }
        }

        // ---- best-first search on level 0 (hnsw.hpp:849-924) with ef = max(efS, topk)
        int ntop = 0, ncand = 0;      // meaningful on lane 0
        uint32_t nvis = 0;            // warp-uniform
        float ub = curr_dist;         // distance(query, entry) recomputed by the reference: same value
        n_dist += 1;
        if (lane == 0) {
            heap_push<true>(topq, ntop, ub, curr);
            heap_push<false>(cand, ncand, ub, curr);
            atomicOr(&bitmap[curr >> 5], 1u << (curr & 31u));
            vlist[0] = curr;
        }
        nvis = 1;
        __syncwarp();
        int overflow = 0;
        for (;;) {
            int done = 0;
            uint32_t node = 0;
            if (lane == 0) {
                if (ncand == 0 || __uint_as_float(cand[0].x) > ub) done = 1;
                else { node = cand[0].y; heap_pop<false>(cand, ncand); }
            }
            done = __shfl_sync(kFull, done, 0);
            if (done) break;
            node = __shfl_sync(kFull, node, 0);
            const uint32_t* nb = ix.nbr0 + static_cast<uint64_t>(node) * ix.n0stride;
            const uint32_t deg = min(nb[0], ix.l0_max_degree);
            n_expand += 1;
            // A. unvisited neighbours, in list order
            uint32_t nu = 0;
            for (uint32_t b = 0; b < deg; b += 32) {
                const uint32_t j = b + lane;
                uint32_t id = 0;
                bool fresh = false;
                if (j < deg) {
                    id = nb[1 + j];
                    const uint32_t bit = 1u << (id & 31u);
                    const uint32_t old = atomicOr(&bitmap[id >> 5], bit);
                    fresh = (old & bit) == 0u;
                }
                const unsigned mask = __ballot_sync(kFull, fresh);
                if (fresh) nb_ids[nu + __popc(mask & ((1u << lane) - 1u))] = id;
                nu += __popc(mask);
            }
            __syncwarp();
            for (uint32_t i = lane; i < nu; i += 32) {
                if (nvis + i < vcap) vlist[nvis + i] = nb_ids[i];
            }
            nvis += nu;
            // B. all distances
            if (nu) distances(nu);
            n_dist += nu;
            // C. sequential replay of the queue updates (hnsw.hpp:904-914)
            if (lane == 0) {
                for (uint32_t i = 0; i < nu; ++i) {
                    const float nd = nb_dist[i];
                    if (static_cast<uint32_t>(ntop) < ef || nd < ub) {
                        if (static_cast<uint32_t>(ncand) >= vcap) { overflow = 1; break; }
                        const uint32_t id = nb_ids[i];
                        heap_push<false>(cand, ncand, nd, id);
                        heap_push<true>(topq, ntop, nd, id);
                        if (static_cast<uint32_t>(ntop) > ef) heap_pop<true>(topq, ntop);
                        if (ntop > 0) ub = __uint_as_float(topq[0].x);
                    }
                }
            }
            overflow = __shfl_sync(kFull, overflow, 0);
            if (overflow) break;
            __syncwarp();
        }
        if (overflow) {
            if (lane == 0) atomicExch(&ctrl[1], 1ull);
        }

        // ---- trim to topk and sort ascending (hnsw.hpp:961-970)
        if (lane == 0) {
            if (topk < efS) while (static_cast<uint32_t>(ntop) > topk) heap_pop<true>(topq, ntop);
            int n = ntop;
            while (n > 1) {  // std::sort_heap
                const uint2 value = topq[n - 1];
                topq[n - 1] = topq[0];
                heap_adjust<true>(topq, 0, n - 1, value);
                --n;
            }
        }
        ntop = __shfl_sync(kFull, ntop, 0);
        __syncwarp();
        for (int k = lane; k < ntop && k < static_cast<int>(topk); k += 32) {
            const uint2 e = topq[k];
            out_idx[static_cast<uint64_t>(q) * topk + k] = e.y;
            out_val[static_cast<uint64_t>(q) * topk + k] = __uint_as_float(e.x);
        }
        // ---- reset the visited bitmap for the next query of this warp
        if (nvis <= vcap) {
            for (uint32_t i = lane; i < nvis; i += 32) bitmap[vlist[i] >> 5] = 0u;
        } else {
            for (uint32_t w = lane; w < bitmap_words; w += 32) bitmap[w] = 0u;
        }
        __syncwarp();
        if (lane == 0) {
            atomicAdd(&ctrl[2], n_dist);
            atomicAdd(&ctrl[3], n_expand);
            atomicAdd(&ctrl[4], n_hops);
            atomicAdd(&ctrl[5], 1ull);
        }
        if (SPARSE) {  // per half-warp partial sums of the stored entries read
            n_entries += __shfl_xor_sync(kFull, n_entries, 16);
            if (lane == 0) atomicAdd(&ctrl[6], n_entries);
        }
    }
}

}  // namespace

// ------------------------------------------------------------------------------------------------------------------
HnswEngine::HnswEngine(std::unique_ptr<HnswHostIndex> host, int device) : host_(std::move(host)), device_(device) {
    PB200_CUDA(cudaSetDevice(device_));
    PB200_CUDA(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
    for (auto& e : ev_) PB200_CUDA(cudaEventCreate(&e));
    const HnswHostIndex& H = *host_;
    const uint64_t N = H.num_node;
    const uint32_t vs = H.sparse ? 0u : H.vstride(), n0 = H.n0stride(), d = H.feat_dim;
    uint64_t sparse_bytes = 0;
    if (H.sparse) {
        // entry offsets + interleaved {index, value} entries + neighbour lists, re-laid-out on the host
        std::vector<unsigned long long> ptr(N + 1, 0ull);
        for (uint64_t i = 0; i < N; ++i) {
            const float* v; const uint32_t* c;
            ptr[i + 1] = ptr[i] + H.l0_sparse_row(static_cast<uint32_t>(i), &v, &c);
        }
        const uint64_t nnz = ptr[N];
        sp_ptr_.upload(ptr.data(), N + 1, stream_);
        sp_ent_.reserve(std::max<uint64_t>(nnz, 1));
        nbr0_.reserve(N * n0);
        const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(N, 1u << 16));
        PinnedBuffer<uint2> se;
        PinnedBuffer<uint32_t> sn;
        sn.reserve(chunk * n0);
        for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
            const uint64_t cn = std::min(chunk, N - c0);
            const uint64_t e0 = ptr[c0], en = ptr[c0 + cn] - e0;
            se.reserve(std::max<uint64_t>(en, 1));
            std::memset(sn.get(), 0, cn * n0 * 4);
            parallel_for_chunks(cn, [&](uint64_t r) {
                const uint32_t node = static_cast<uint32_t>(c0 + r);
                const float* v; const uint32_t* c;
                const uint32_t len = H.l0_sparse_row(node, &v, &c);
                uint2* dst = se.get() + (ptr[node] - e0);
                for (uint32_t j = 0; j < len; ++j) {
                    uint32_t bits;
                    std::memcpy(&bits, v + j, 4);
                    dst[j] = make_uint2(c[j], bits);
                }
                const uint32_t* nb = H.l0_neighborhood(node);
                uint32_t* nd = sn.get() + r * n0;
                const uint32_t deg = std::min(nb[0], H.l0_max_degree);
                nd[0] = deg;
                for (uint32_t j = 0; j < deg; ++j) nd[1 + j] = nb[1 + j];
            });
            if (en) PB200_CUDA(cudaMemcpyAsync(sp_ent_.get() + e0, se.get(), en * sizeof(uint2), cudaMemcpyHostToDevice, stream_));
            PB200_CUDA(cudaMemcpyAsync(nbr0_.get() + c0 * n0, sn.get(), cn * n0 * 4, cudaMemcpyHostToDevice, stream_));
            PB200_CUDA(cudaStreamSynchronize(stream_));
        }
        sparse_bytes = (N + 1) * 8 + nnz * 8;
    } else {
    vec_.reserve(N * vs);
    nbr0_.reserve(N * n0);
    // re-layout in chunks through a pinned staging buffer
    const uint64_t chunk = std::max<uint64_t>(1, std::min<uint64_t>(N, (256ull << 20) / (static_cast<uint64_t>(vs) * 4 + n0 * 4)));
    PinnedBuffer<float> sv;
    PinnedBuffer<uint32_t> sn;
    sv.reserve(chunk * vs);
    sn.reserve(chunk * n0);
    std::vector<uint32_t> pos(d);
    for (uint32_t i = 0; i < d; ++i) pos[i] = H.permuted_pos(i);
    for (uint64_t c0 = 0; c0 < N; c0 += chunk) {
        const uint64_t cn = std::min(chunk, N - c0);
        std::memset(sv.get(), 0, cn * vs * 4);
        std::memset(sn.get(), 0, cn * n0 * 4);
        parallel_for_chunks(cn, [&](uint64_t r) {
            const uint32_t node = static_cast<uint32_t>(c0 + r);
            const float* src = H.l0_vector(node);
            float* dst = sv.get() + r * vs;
            for (uint32_t i = 0; i < d; ++i) dst[pos[i]] = src[i];
            const uint32_t* nb = H.l0_neighborhood(node);
            uint32_t* nd = sn.get() + r * n0;
            const uint32_t deg = std::min(nb[0], H.l0_max_degree);
            nd[0] = deg;
            for (uint32_t j = 0; j < deg; ++j) nd[1 + j] = nb[1 + j];
        });
        PB200_CUDA(cudaMemcpyAsync(vec_.get() + c0 * vs, sv.get(), cn * vs * 4, cudaMemcpyHostToDevice, stream_));
        PB200_CUDA(cudaMemcpyAsync(nbr0_.get() + c0 * n0, sn.get(), cn * n0 * 4, cudaMemcpyHostToDevice, stream_));
        PB200_CUDA(cudaStreamSynchronize(stream_));
    }
    }
    uint64_t l1_len = 0;
    if (H.max_level > 0) {
        l1_len = static_cast<uint64_t>(N) * H.l1_node_mem_size;
        l1_.upload(H.l1_buffer, l1_len, stream_);
        PB200_CUDA(cudaStreamSynchronize(stream_));
    }
    index_bytes_ = N * vs * 4 + N * n0 * 4 + l1_len * 4 + sparse_bytes;
    view_.sp_ptr = sp_ptr_.get();
    view_.sp_ent = sp_ent_.get();
    view_.vec = vec_.get();
    view_.nbr0 = nbr0_.get();
    view_.l1 = l1_.get();
    view_.num_node = H.num_node;
    view_.max_level = H.max_level;
    view_.init_node = H.init_node;
    view_.feat_dim = d;
    view_.vstride = vs;
    view_.main_pad = H.sparse ? 0u : H.main_pad();
    view_.tail_len = H.sparse ? 0u : H.tail_len();
    view_.n0stride = n0;
    view_.l0_max_degree = H.l0_max_degree;
    view_.l1_node_mem = H.l1_node_mem_size;
    view_.l1_level_mem = H.l1_level_mem_size;
    view_.l1_max_degree = H.l1_max_degree;
    view_.metric = H.metric;
    ctrl_.reserve(8);
    PB200_CUDA(cudaMemsetAsync(ctrl_.get(), 0, 8 * sizeof(unsigned long long), stream_));
    PB200_CUDA(cudaStreamSynchronize(stream_));
    const int max_smem = 200 * 1024;
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_IP, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_L2, 0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_IP, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_L2, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_IP, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_L2, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_IP, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    PB200_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<HNSW_L2, 0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
    stages_ = 4;  // rows in flight per warp through the bulk-copy ring; 0 = direct loads (first-generation kernel)
    if (const char* env = std::getenv("PB200_HNSW_STAGES")) {
        const int v = std::atoi(env);
        stages_ = (v <= 0) ? 0 : (v <= 4 ? 4 : 8);
    }
    // the mapped file is no longer needed once the arrays live in HBM
    host_->l0_buffer = nullptr;
    host_->l0_mem_start = nullptr;
    host_->l1_buffer = nullptr;
    host_->store.reset();
}

HnswEngine::~HnswEngine() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    for (auto& e : ev_) if (e) cudaEventDestroy(e);
    if (stream_) cudaStreamDestroy(stream_);
}

uint32_t HnswEngine::per_warp_smem_(uint32_t ef, uint32_t* nbmax_out) const {
    const HnswHostIndex& H = *host_;
    const uint32_t nbmax = ((std::max(H.l0_max_degree, H.l1_max_degree) + 31u) / 32u) * 32u;
    const bool top_in_smem = ef <= kEfSmemMax;
    if (nbmax_out) *nbmax_out = nbmax;
    if (H.sparse)  // [query indices | query values | filter | ids | distances | result heap]
        return (qcap_ * 8u + kSpFilterWords * 4u + nbmax * 8 + (top_in_smem ? (ef + 1) * 8 : 0) + 15u) & ~15u;
    // [query | stages_ ring slots | stages_ mbarriers | ids | distances | result heap]
    return (H.vstride() * 4 * (1u + static_cast<uint32_t>(stages_)) + static_cast<uint32_t>(stages_) * 8u + nbmax * 8 +
            (top_in_smem ? (ef + 1) * 8 : 0) + 15u) & ~15u;
}

void HnswEngine::set_stages(int stages) {
    stages_ = (stages <= 0) ? 0 : (stages <= 4 ? 4 : 8);
    n_warps_ = 0;  // forces the scratch / launch geometry to be recomputed
}

void HnswEngine::ensure_scratch_(uint32_t ef) {
    const HnswHostIndex& H = *host_;
    const bool top_in_smem = ef <= kEfSmemMax;
    const uint32_t per_warp = per_warp_smem_(ef, nullptr);
    uint32_t warps = 8;
    while (warps > 1 && static_cast<uint64_t>(warps) * per_warp > 96u * 1024u) warps >>= 1;
    if (static_cast<uint64_t>(warps) * per_warp > 200u * 1024u)
        throw std::runtime_error("pecos_b200: HNSW query dimension too large for the shared-memory staging area");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_);
    const uint32_t ctas_per_sm = std::max<uint32_t>(1, std::min<uint32_t>((H.sparse ? 32u : 16u) / warps, static_cast<uint32_t>((220u * 1024u) / (static_cast<uint64_t>(warps) * per_warp))));
    uint32_t n_ctas = static_cast<uint32_t>(sms) * ctas_per_sm;
    // bound the scratch footprint (bitmap N/8 bytes per warp)
    const uint64_t words = (static_cast<uint64_t>(H.num_node) + 31) / 32;
    uint32_t vcap = 32768;
    if (const char* env = std::getenv("PB200_HNSW_VCAP")) vcap = static_cast<uint32_t>(std::max<uint64_t>(64, std::strtoull(env, nullptr, 10)));
    vcap = std::max(vcap, vcap_floor_);  // raised by launch_ after a candidate-queue overflow
    vcap = static_cast<uint32_t>(std::min<uint64_t>(vcap, static_cast<uint64_t>(H.num_node) + 1));
    const uint64_t per_warp_scratch = words * 4 + static_cast<uint64_t>(vcap) * 12 + (top_in_smem ? 0 : static_cast<uint64_t>(ef + 1) * 8);
    while (n_ctas > static_cast<uint32_t>(sms) && static_cast<uint64_t>(n_ctas) * warps * per_warp_scratch > (24ull << 30)) n_ctas -= sms;
    const uint32_t n_warps = n_ctas * warps;
    if (n_warps != n_warps_ || vcap != vcap_ || (!top_in_smem && ef > scratch_ef_) || warps != warps_per_cta_) {
        bitmap_.reserve(static_cast<uint64_t>(n_warps) * words);
        PB200_CUDA(cudaMemsetAsync(bitmap_.get(), 0, static_cast<uint64_t>(n_warps) * words * 4, stream_));
        vlist_.reserve(static_cast<uint64_t>(n_warps) * vcap);
        cand_.reserve(static_cast<uint64_t>(n_warps) * vcap);
        if (!top_in_smem) { topk_heap_.reserve(static_cast<uint64_t>(n_warps) * (ef + 1)); scratch_ef_ = ef; }
        n_warps_ = n_warps; warps_per_cta_ = warps; n_ctas_ = n_ctas; vcap_ = vcap;
    }
}

double HnswEngine::launch_once_(const float* q_dev, uint32_t nq, uint32_t efS, uint32_t topk, bool* overflow) {
    const HnswHostIndex& H = *host_;
    const uint32_t ef = std::max(efS, topk);
    if (ef == 0) throw std::runtime_error("pecos_b200: efS and topk are both zero");
    ensure_scratch_(ef);
    uint32_t nbmax = 0;
    const bool top_in_smem = ef <= kEfSmemMax;
    const uint32_t per_warp = per_warp_smem_(ef, &nbmax);
    const uint32_t words = static_cast<uint32_t>((static_cast<uint64_t>(H.num_node) + 31) / 32);
    PB200_CUDA(cudaMemsetAsync(ctrl_.get(), 0, 8 * sizeof(unsigned long long), stream_));
    PB200_CUDA(cudaMemsetAsync(out_idx_.get(), 0, static_cast<uint64_t>(nq) * topk * 4, stream_));
    PB200_CUDA(cudaMemsetAsync(out_val_.get(), 0, static_cast<uint64_t>(nq) * topk * 4, stream_));
    const uint32_t ctas = std::max<uint32_t>(1, std::min<uint32_t>(n_ctas_, (nq + warps_per_cta_ - 1) / warps_per_cta_));
    const size_t smem = static_cast<size_t>(warps_per_cta_) * per_warp;
    const HnswSparseQueries sq{q_ptr_.get(), q_idx_.get(), q_dev, qcap_};  // csr batch (sparse indices; q_dev = its values)
    PB200_CUDA(cudaEventRecord(ev_[0], stream_));
    auto launch = [&](auto kernel) {
        kernel<<<ctas, warps_per_cta_ * 32, smem, stream_>>>(view_, q_dev, sq, nq, efS, topk, ef, out_idx_.get(), out_val_.get(),
                                                             bitmap_.get(), words, vlist_.get(), cand_.get(), vcap_,
                                                             top_in_smem ? nullptr : topk_heap_.get(), nbmax, per_warp, ctrl_.get());
    };
    const bool ip = H.metric == HNSW_IP;
    if (H.sparse) { if (ip) launch(hnsw_search_kernel<HNSW_IP, 0, true>); else launch(hnsw_search_kernel<HNSW_L2, 0, true>); }
    else if (stages_ == 0) { if (ip) launch(hnsw_search_kernel<HNSW_IP, 0, false>); else launch(hnsw_search_kernel<HNSW_L2, 0, false>); }
    else if (stages_ == 4) { if (ip) launch(hnsw_search_kernel<HNSW_IP, 4, false>); else launch(hnsw_search_kernel<HNSW_L2, 4, false>); }
    else { if (ip) launch(hnsw_search_kernel<HNSW_IP, 8, false>); else launch(hnsw_search_kernel<HNSW_L2, 8, false>); }
    PB200_CUDA(cudaGetLastError());
    PB200_CUDA(cudaEventRecord(ev_[1], stream_));
    ++launches_;
    PB200_CUDA(cudaEventSynchronize(ev_[1]));
    float ms = 0.f;
    PB200_CUDA(cudaEventElapsedTime(&ms, ev_[0], ev_[1]));
    unsigned long long flag[2] = {0, 0};
    PB200_CUDA(cudaMemcpy(flag, ctrl_.get(), sizeof(flag), cudaMemcpyDeviceToHost));
    *overflow = flag[1] != 0;
    last_ms_ = ms;
    return ms;
}

// A query whose candidate queue outgrows the per-warp scratch (vcap entries) flags an overflow; the batch is then re-run with
// twice the capacity (at most num_node + 1 entries, which can never overflow: a node enters the queue at most once), instead
// of aborting the host process.
double HnswEngine::launch_(const float* q_dev, uint32_t nq, uint32_t efS, uint32_t topk) {
    for (;;) {
        bool overflow = false;
        const double ms = launch_once_(q_dev, nq, efS, topk, &overflow);
        if (!overflow) return ms;
        const uint64_t cap_max = static_cast<uint64_t>(host_->num_node) + 1;
        if (vcap_ >= cap_max) throw std::runtime_error("pecos_b200: HNSW candidate queue overflow at full capacity (internal error)");
        vcap_floor_ = static_cast<uint32_t>(std::min<uint64_t>(cap_max, static_cast<uint64_t>(vcap_) * 2));
        ++vcap_retries_;
    }
}


void HnswEngine::predict(const float* X, uint32_t nq, uint32_t d, uint32_t efS, uint32_t topk, uint32_t* ret_idx, float* ret_val) {
    PB200_CUDA(cudaSetDevice(device_));
    if (host_->sparse) throw std::runtime_error("pecos_b200: dense queries against a sparse (csr) HNSW index");
    if (d != host_->feat_dim) throw std::runtime_error("pecos_b200: query dimension != index dimension");
    if (nq == 0 || topk == 0) return;
    q_dev_.upload(X, static_cast<uint64_t>(nq) * d, stream_);
    out_idx_.reserve(static_cast<uint64_t>(nq) * topk);
    out_val_.reserve(static_cast<uint64_t>(nq) * topk);
    launch_(q_dev_.get(), nq, efS, topk);
    // rows with fewer than topk results keep the caller's zeros (libpecos.cpp:554-558): our buffers were zeroed too
    PB200_CUDA(cudaMemcpyAsync(ret_idx, out_idx_.get(), static_cast<uint64_t>(nq) * topk * 4, cudaMemcpyDeviceToHost, stream_));
    PB200_CUDA(cudaMemcpyAsync(ret_val, out_val_.get(), static_cast<uint64_t>(nq) * topk * 4, cudaMemcpyDeviceToHost, stream_));
    PB200_CUDA(cudaStreamSynchronize(stream_));
}

// csr query batch -> device (row offsets rebased to the batch), and the per-warp staging capacity for its longest row
void HnswEngine::upload_csr_(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq) {
    const uint64_t e0 = row_ptr[0], nnz = row_ptr[nq] - e0;
    std::vector<unsigned long long> ptr(static_cast<size_t>(nq) + 1);
    uint64_t longest = 0;
    for (uint32_t i = 0; i <= nq; ++i) {
        ptr[i] = row_ptr[i] - e0;
        if (i) longest = std::max<uint64_t>(longest, row_ptr[i] - row_ptr[i - 1]);
    }
    q_ptr_.upload(ptr.data(), static_cast<uint64_t>(nq) + 1, stream_);
    q_idx_.reserve(std::max<uint64_t>(nnz, 1));
    q_dev_.reserve(std::max<uint64_t>(nnz, 1));
    if (nnz) {
        PB200_CUDA(cudaMemcpyAsync(q_idx_.get(), col_idx + e0, nnz * 4, cudaMemcpyHostToDevice, stream_));
        PB200_CUDA(cudaMemcpyAsync(q_dev_.get(), val + e0, nnz * 4, cudaMemcpyHostToDevice, stream_));
    }
    PB200_CUDA(cudaStreamSynchronize(stream_));  // `ptr` is a local
    const uint32_t qcap = static_cast<uint32_t>(std::min<uint64_t>(kSpQcapMax, (std::max<uint64_t>(longest, 1) + 31) / 32 * 32));
    if (qcap != qcap_) { qcap_ = qcap; n_warps_ = 0; }  // launch geometry depends on the staging capacity
}

void HnswEngine::predict_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq, uint32_t cols,
                             uint32_t efS, uint32_t topk, uint32_t* ret_idx, float* ret_val) {
    PB200_CUDA(cudaSetDevice(device_));
    if (!host_->sparse) throw std::runtime_error("pecos_b200: csr queries against a dense HNSW index");
    if (cols != host_->feat_dim) throw std::runtime_error("pecos_b200: query dimension != index dimension");
    if (nq == 0 || topk == 0) return;
    upload_csr_(row_ptr, col_idx, val, nq);
    out_idx_.reserve(static_cast<uint64_t>(nq) * topk);
    out_val_.reserve(static_cast<uint64_t>(nq) * topk);
    launch_(q_dev_.get(), nq, efS, topk);
    PB200_CUDA(cudaMemcpyAsync(ret_idx, out_idx_.get(), static_cast<uint64_t>(nq) * topk * 4, cudaMemcpyDeviceToHost, stream_));
    PB200_CUDA(cudaMemcpyAsync(ret_val, out_val_.get(), static_cast<uint64_t>(nq) * topk * 4, cudaMemcpyDeviceToHost, stream_));
    PB200_CUDA(cudaStreamSynchronize(stream_));
}

void HnswEngine::resident_upload_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq, uint32_t cols) {
    PB200_CUDA(cudaSetDevice(device_));
    if (!host_->sparse) throw std::runtime_error("pecos_b200: csr queries against a dense HNSW index");
    if (cols != host_->feat_dim) throw std::runtime_error("pecos_b200: query dimension != index dimension");
    upload_csr_(row_ptr, col_idx, val, nq);
    res_nq_ = nq;
    res_d_ = cols;
}

void HnswEngine::resident_upload(const float* X, uint32_t nq, uint32_t d) {
    PB200_CUDA(cudaSetDevice(device_));
    if (host_->sparse) throw std::runtime_error("pecos_b200: dense queries against a sparse (csr) HNSW index");
    if (d != host_->feat_dim) throw std::runtime_error("pecos_b200: query dimension != index dimension");
    q_dev_.upload(X, static_cast<uint64_t>(nq) * d, stream_);
    PB200_CUDA(cudaStreamSynchronize(stream_));
    res_nq_ = nq;
    res_d_ = d;
}

double HnswEngine::resident_predict(uint32_t efS, uint32_t topk) {
    PB200_CUDA(cudaSetDevice(device_));
    if (!res_nq_) throw std::runtime_error("pecos_b200: no resident query batch uploaded");
    out_idx_.reserve(static_cast<uint64_t>(res_nq_) * topk);
    out_val_.reserve(static_cast<uint64_t>(res_nq_) * topk);
    res_topk_ = topk;
    return launch_(q_dev_.get(), res_nq_, efS, topk);
}

void HnswEngine::resident_fetch(uint32_t* ret_idx, float* ret_val) {
    PB200_CUDA(cudaSetDevice(device_));
    PB200_CUDA(cudaMemcpy(ret_idx, out_idx_.get(), static_cast<uint64_t>(res_nq_) * res_topk_ * 4, cudaMemcpyDeviceToHost));
    PB200_CUDA(cudaMemcpy(ret_val, out_val_.get(), static_cast<uint64_t>(res_nq_) * res_topk_ * 4, cudaMemcpyDeviceToHost));
}

HnswCounters HnswEngine::counters() {
    PB200_CUDA(cudaSetDevice(device_));
    unsigned long long h[8];
    PB200_CUDA(cudaMemcpy(h, ctrl_.get(), sizeof(h), cudaMemcpyDeviceToHost));
    HnswCounters c;
    c.n_dist = h[2]; c.n_expand = h[3]; c.n_hops = h[4]; c.n_queries = h[5]; c.n_entries = h[6];
    return c;
}

}  // namespace pb200
