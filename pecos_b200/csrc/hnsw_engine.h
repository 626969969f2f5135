// HNSW dense search engine on one B200.
//
// Replaces (reference, CPU/OpenMP, one Searcher per thread):
//   c_ann_hnsw_predict_* ............. pecos/core/libpecos.cpp:527-564
//   HNSW::predict_single ............. pecos/core/ann/hnsw.hpp:927-971
//   HNSW::search_level ............... pecos/core/ann/hnsw.hpp:849-924
//   Searcher / SetOfVistedNodes ...... pecos/core/ann/hnsw.hpp:341-446   (here: per-warp bitmap + heaps in HBM/shared memory)
//   FeatVecDense{IP,L2}Simd::distance  pecos/core/ann/feat_vectors.hpp:134-162 + distance_impl/x86.hpp:121-157, :256-296
#pragma once

#include <memory>
#include <vector>

#include "cuda_util.h"
#include "hnsw_host.h"

namespace pb200 {

struct HnswDev {
    const float* vec;
    const uint32_t* nbr0;
    const uint32_t* l1;
    uint32_t num_node, max_level, init_node, feat_dim;
    uint32_t vstride, main_pad, tail_len, n0stride, l0_max_degree;
    uint32_t l1_node_mem, l1_level_mem, l1_max_degree;
    int metric;
    // sparse (csr) indices: row r = entries [sp_ptr[r], sp_ptr[r+1]) of sp_ent, {index, value bits}, indices ascending
    const unsigned long long* sp_ptr;
    const uint2* sp_ent;
};

struct HnswSparseQueries {  // device CSR of the query batch (sparse indices)
    const unsigned long long* ptr;
    const uint32_t* idx;
    const float* val;
    uint32_t qcap;  // query entries staged per warp in shared memory (longer rows are searched in global memory)
};

struct HnswCounters {  // algorithmic-byte counters of SURVEY.md 8(d), totals over the last search call
    unsigned long long n_dist = 0;    // distance evaluations (one base vector read each)
    unsigned long long n_expand = 0;  // level-0 expansions (one neighbour-list read each)
    unsigned long long n_hops = 0;    // upper-level neighbourhood reads
    unsigned long long n_queries = 0;
    unsigned long long n_entries = 0;  // sparse indices: stored entries of the evaluated base rows (8 bytes each)
};

class HnswEngine {
public:
    HnswEngine(std::unique_ptr<HnswHostIndex> host, int device);
    ~HnswEngine();

    const HnswHostIndex& host() const { return *host_; }
    int metric() const { return host_->metric; }

    // Host-buffer entry point: X row-major nq x d; ret arrays nq x topk (caller-zeroed, like the reference).
    void predict(const float* X, uint32_t nq, uint32_t d, uint32_t efS, uint32_t topk, uint32_t* ret_idx, float* ret_val);

    // Sparse index, csr queries (column indices ascending within a row): the same walk, distances by ordered sparse intersection.
    void predict_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq, uint32_t cols, uint32_t efS,
                     uint32_t topk, uint32_t* ret_idx, float* ret_val);
    void resident_upload_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq, uint32_t cols);
    bool sparse() const { return host_->sparse; }

    // Device-resident queries (bench "value" leg).
    void resident_upload(const float* X, uint32_t nq, uint32_t d);
    double resident_predict(uint32_t efS, uint32_t topk);  // returns device ms of the search kernel
    void resident_fetch(uint32_t* ret_idx, float* ret_val);

    HnswCounters counters();
    uint64_t launches() const { return launches_; }
    uint32_t vcap_retries() const { return vcap_retries_; }  // batches re-run after a candidate-queue overflow
    uint64_t index_bytes() const { return index_bytes_; }
    double last_kernel_ms() const { return last_ms_; }
    // rows kept in flight per warp by the bulk-copy ring: 0 (direct loads), 4 or 8
    void set_stages(int stages);
    int stages() const { return stages_; }

private:
    void ensure_scratch_(uint32_t ef);
    uint32_t per_warp_smem_(uint32_t ef, uint32_t* nbmax_out) const;
    double launch_(const float* q_dev, uint32_t nq, uint32_t efS, uint32_t topk);
    double launch_once_(const float* q_dev, uint32_t nq, uint32_t efS, uint32_t topk, bool* overflow);

    std::unique_ptr<HnswHostIndex> host_;
    int device_ = 0;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[2] = {nullptr, nullptr};
    DeviceBuffer<float> vec_;
    DeviceBuffer<uint32_t> nbr0_;
    DeviceBuffer<uint32_t> l1_;
    HnswDev view_{};
    uint64_t index_bytes_ = 0;

    // per-warp scratch
    uint32_t n_warps_ = 0, warps_per_cta_ = 0, n_ctas_ = 0;
    uint32_t vcap_ = 0, scratch_ef_ = 0;
    uint32_t vcap_floor_ = 0;    // minimum candidate-queue capacity (doubled after an overflow)
    uint32_t vcap_retries_ = 0;
    DeviceBuffer<uint32_t> bitmap_;
    DeviceBuffer<uint32_t> vlist_;
    DeviceBuffer<uint2> cand_;
    DeviceBuffer<uint2> topk_heap_;
    DeviceBuffer<unsigned long long> ctrl_;  // [0] query counter, [1] error flag, [2..5] counters

    DeviceBuffer<unsigned long long> sp_ptr_;
    DeviceBuffer<uint2> sp_ent_;
    DeviceBuffer<unsigned long long> q_ptr_;  // csr query batch
    DeviceBuffer<uint32_t> q_idx_;
    uint32_t qcap_ = 0;
    void upload_csr_(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t nq);
    DeviceBuffer<float> q_dev_;
    DeviceBuffer<uint32_t> out_idx_;
    DeviceBuffer<float> out_val_;
    uint32_t res_nq_ = 0, res_d_ = 0, res_topk_ = 0;
    PinnedBuffer<uint32_t> stage_idx_;
    PinnedBuffer<float> stage_val_;

    uint64_t launches_ = 0;
    double last_ms_ = 0.0;
    int stages_ = 4;
};

}  // namespace pb200
