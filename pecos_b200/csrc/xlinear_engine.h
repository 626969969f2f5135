// XR-Linear beam-search engine on one B200: device-resident model + the two kernels per tree layer.
//
// Replaces (reference, CPU/OpenMP):
//   HierarchicalMLModel::predict ........ pecos/core/xmc/inference.hpp:2446-2488
//   MLModel::predict_internal ........... pecos/core/xmc/inference.hpp:2029-2080
//     prolongate_predictions ............ :1155-1219   (implicit here: beam slot j -> chunk j, no materialised pattern)
//     w_ops::compute_sparse_predictions . :925-1007    -> xl_chunk_scores_kernel
//       chunk_ops<csr|drm, bin_search> .. :769-839
//     transform / combine / sorted_csr .. :1360-1384, :1223-1298 -> xl_topk_kernel
//     reorder_prediction ................ :1919-1923   (label_of_col lookup in xl_topk_kernel)
#pragma once

#include <memory>
#include <string>
#include <vector>

#include "cuda_util.h"
#include "xlinear_host.h"

namespace pb200 {

struct LayerDev {
    const ChunkHeader* chunks;
    const uint32_t* meta;
    const uint32_t* rowext;        // same indexing as meta: chunk region [meta_off, meta_off + 2R) = {begin, end} per row
    const uint2* entries;
    const uint32_t* label_of_col;  // nullptr when the layer is contiguously ordered
    const uint2* featmap;          // per-chunk {bits, prefix} cells for query-driven lookups; nullptr = stream row lists
    uint32_t fm_words;
    uint32_t n_cols;
    uint32_t n_chunks;
    uint32_t c_max;
    uint32_t w_rows;
    float bias;
    int has_dup_cols;
};

// Chunk-major score kernel (xlinear_cm_kernel.cuh)
struct CmShape {  // load-time, per layer: geometry of the chunk images
    bool ok = false;
    bool direct = false;
    uint32_t words = 0;      // direct: w_rows (table entries); else fm_words (feature-map cells)
    uint32_t r_cap = 0, e_cap = 0, acc_cols = 0;
    uint32_t stages = 2;     // depth of the per-warp cp.async ring of query-feature rounds
    uint32_t col_cap = 0;    // widest column range: wider chunks are cut into ranges ("virtual chunks"), each with its own image
    uint32_t n_vc = 0;       // virtual chunks of the layer
    const uint32_t* vc_ptr = nullptr;  // DEVICE [n_chunks + 1]: first virtual chunk of every chunk
    uint32_t warps_fit = 0;  // warps of the score kernel that fit next to one image in shared memory
    uint32_t off_lookup = 16, off_pre = 0, off_rp = 0, off_ew = 0, off_ec = 0;  // byte offsets inside an image
    uint32_t img_bytes = 0;  // image stride (multiple of 128)
};

// Device view of a batch of queries (either CSR or row-major dense).
struct QueryDev {
    const uint64_t* row_ptr;  // CSR: absolute offsets (row_ptr[r] - nnz_base indexes col_idx/val); nullptr for dense
    const uint32_t* col_idx;
    const float* val;         // CSR values, or the dense matrix
    uint64_t nnz_base;
    uint32_t rows;
    uint32_t cols;
    uint32_t max_row_nnz;     // longest row of the batch (sizes the shared-memory staging of the query)
};

struct XLinearStats {  // algorithmic-byte counters of SURVEY.md section 8(d), accumulated by the STATS kernel variant
    unsigned long long chunks;      // (query, chunk) products evaluated
    unsigned long long chunk_rows;  // sum R_p
    unsigned long long matched;     // sum m(q,p)   (bias row included when applied)
    unsigned long long entries;     // sum e(q,p)
    unsigned long long out_cols;    // sum c_p
    unsigned long long query_nnz;   // sum nnz(x_q) (once per query per layer)
    unsigned long long beam_out;    // sum min(k, candidates)
};

struct XLinearLayerProfile {
    double scores_ms = 0.0;  // score kernel of the layer
    double topk_ms = 0.0;    // top-k kernel of the layer
    uint64_t launches = 0;
    int scores_kernel = 0;   // last launch: 0 row-list streaming, 1 feature-map lookup, 2 dense, 3 query-warp, 4 chunk-major, 5 chunk-major without image
    int topk_kernel = 0;     // last launch: 0 block-wide sort, 1 warp arg-max, 2 estimate filter
};

class XLinearEngine {
public:
    XLinearEngine(std::unique_ptr<XLinearHostModel> host, int device);
    ~XLinearEngine();

    const XLinearHostModel& host() const { return *host_; }
    int device() const { return device_; }

    struct Result {  // fixed-stride top-k per query, host side (pinned)
        uint32_t rows = 0;
        uint32_t stride = 0;
        uint32_t out_cols = 0;
        const uint32_t* ids = nullptr;
        const float* vals = nullptr;
        const uint32_t* cnt = nullptr;
    };

    // Host-buffer entry points (H2D + kernels + D2H inside). Exactly one of (row_ptr/col_idx/val) or dense is used.
    Result predict_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows, uint32_t cols,
                       uint32_t beam_size, const char* post_processor, uint32_t only_topk);
    Result predict_drm(const float* dense, uint32_t rows, uint32_t cols, uint32_t beam_size, const char* post_processor,
                       uint32_t only_topk);

    // One layer of the python prediction chain (c_xlinear_single_layer_predict_*, pecos/core/libpecos.cpp:201-235): the
    // engine must hold a one-layer model.  Queries: CSR (row_ptr != nullptr) or row-major dense.  codes_*: the previous
    // layer's prediction as CSR rows x n_chunks, consumed in stored order (nullptr: ones(rows x n_chunks), no combine).
    Result predict_single_layer(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, const float* dense,
                                uint32_t rows, uint32_t cols, const uint64_t* codes_row_ptr, const uint32_t* codes_col_idx,
                                const float* codes_val, const char* post_processor, uint32_t only_topk);

    // predict_on_selected_outputs (c_xlinear_predict_on_selected_outputs_*, pecos/core/libpecos.cpp:179-198): scores of exactly
    // the (query, label) pairs of the CSR pattern `sel_*` (rows x nr_labels), pushed through the hierarchy, no top-k.
    // Result rows have the selected rows' lengths; entry order = the reference's (see xlinear_selected.cuh).
    struct SelectedResult {
        uint32_t rows = 0, cols = 0;
        std::vector<uint64_t> indptr;
        std::vector<uint32_t> indices;
        std::vector<float> data;
    };
    SelectedResult predict_selected(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, const float* dense,
                                    uint32_t rows, uint32_t cols, const uint64_t* sel_ptr, const uint32_t* sel_idx,
                                    uint32_t sel_cols, const char* post_processor, const uint64_t* codes_ptr = nullptr,
                                    const uint32_t* codes_idx = nullptr, const float* codes_val = nullptr);

    // Device-resident queries (bench "value" leg: inputs already in HBM when the timed region starts).
    void resident_upload_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows, uint32_t cols);
    // Runs all layers over the resident batch; results stay in HBM (fetch with resident_fetch). Returns device ms.
    double resident_predict(uint32_t beam_size, const char* post_processor, uint32_t only_topk, bool collect_stats);
    Result resident_fetch();

    // Index sharding (leaf layer split over `world` GPUs).  sharded_local_csr runs every layer on this GPU's
    // shard and writes the LOCAL top-k {key, id, value}[rows][stride] + count[rows] into caller-owned device buffers
    // (the NCCL all-gather send buffers); sharded_merge reduces the gathered [world][rows][stride] lists to the global
    // top-k.  Returns the stride used.
    uint32_t sharded_local_csr(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows, uint32_t cols,
                               uint32_t beam_size, const char* post_processor, uint32_t only_topk, uint32_t stride_capacity,
                               unsigned long long* keys_dev, uint32_t* ids_dev, float* vals_dev, uint32_t* cnt_dev);
    Result sharded_merge(uint32_t world, uint32_t rows, uint32_t stride, uint32_t only_topk, const unsigned long long* g_keys,
                         const uint32_t* g_ids, const float* g_vals, const uint32_t* g_cnt);

    // Packed form: ONE buffer of 16-byte {u64 key, u32 id, f32 value} records [rows][stride] (key == 0: empty slot), so the
    // exchange is a single all-gather; sharded_merge_packed consumes the gathered [world][rows][stride] records.
    uint32_t sharded_local_csr_packed(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val, uint32_t rows, uint32_t cols,
                                      uint32_t beam_size, const char* post_processor, uint32_t only_topk, uint32_t stride_capacity,
                                      void* rec_dev);
    Result sharded_merge_packed(uint32_t world, uint32_t rows, uint32_t stride, uint32_t only_topk, const void* g_rec);

    void set_profile(bool on) { profile_ = on; }
    // false = stream the chunk row lists (first-generation kernel) even when feature maps exist; for A/B tests
    void set_kernel_mode(int mode);
    bool has_feature_maps() const;
    const std::vector<XLinearLayerProfile>& layer_profile() const { return layer_profile_; }
    void reset_profile();
    const std::vector<XLinearStats>& layer_stats() const { return layer_stats_; }
    uint64_t launches() const { return launches_; }
    uint64_t model_bytes() const { return model_bytes_; }

private:
    struct LayerPlan {
        uint32_t k = 0;       // only_topk used for this layer
        uint32_t b_prev = 0;  // beam capacity entering the layer
        uint32_t k_cap = 0;   // beam capacity leaving the layer
        PostProc pp;
    };
    struct LayerStore {
        DeviceBuffer<ChunkHeader> chunks;
        DeviceBuffer<uint32_t> meta;
        DeviceBuffer<uint32_t> rowext;
        DeviceBuffer<uint2> entries;
        DeviceBuffer<uint32_t> label_of_col;
        DeviceBuffer<uint2> featmap;
        uint32_t e_max = 0;  // most entries of one chunk (sizes the chunk-major kernel's shared-memory staging)
        DeviceBuffer<unsigned char> cm_images;  // packed chunk images of the chunk-major kernel (empty: layer not eligible)
        DeviceBuffer<uint32_t> cm_vc_ptr;
        CmShape cm_shape;
        LayerDev view{};
    };

    std::vector<LayerPlan> make_plan_(uint32_t beam_size, const char* post_processor, uint32_t only_topk) const;
    void ensure_workspace_(const std::vector<LayerPlan>& plan, uint32_t tile_rows);
    uint32_t pick_tile_rows_(const std::vector<LayerPlan>& plan, uint32_t rows) const;
    // layers [d_begin, d_end) over one tile of queries; the tile's beam / candidate rows start at row_off_ of the workspace
    void run_tile_(const QueryDev& q, const std::vector<LayerPlan>& plan, bool collect_stats, bool ext_beam = false,
                   int combine_first = 0, size_t d_begin = 0, size_t d_end = static_cast<size_t>(-1));
    uint32_t* bid_(int b) const { return beam_id_[b].get() + static_cast<uint64_t>(row_off_) * beam_stride_; }
    float* bval_(int b) const { return beam_val_[b].get() + static_cast<uint64_t>(row_off_) * beam_stride_; }
    uint32_t* bcnt_(int b) const { return beam_cnt_[b].get() + row_off_; }
    float* cand_at_(uint64_t stride_q) const { return cand_.get() + static_cast<uint64_t>(row_off_) * stride_q; }
    uint32_t row_off_ = 0;
    int score_layer_(size_t d, const QueryDev& q, uint32_t b_prev, int cur, bool collect_stats);
    Result finish_result_(uint32_t rows, uint32_t stride);

    struct SelIndex {  // per layer: label -> (chunk, column offset); built on the first predict_selected call
        std::vector<uint32_t> chunk_of_label, offset_of_label;
    };
    std::vector<SelIndex> sel_index_;

    std::unique_ptr<XLinearHostModel> host_;
    int device_ = 0;
    cudaStream_t stream_ = nullptr;
    std::vector<LayerStore> layers_;
    uint64_t model_bytes_ = 0;

    // per-tile workspace
    DeviceBuffer<uint32_t> beam_id_[2];
    DeviceBuffer<float> beam_val_[2];
    DeviceBuffer<uint32_t> beam_cnt_[2];
    DeviceBuffer<float> cand_;
    DeviceBuffer<unsigned long long> sortbuf_;
    DeviceBuffer<unsigned long long> stats_dev_;
    int final_buf_ = 0;  // which ping-pong buffer holds the last layer's output
    uint32_t beam_stride_ = 0;

    // staged inputs (host-buffer path) and resident batch
    DeviceBuffer<uint64_t> x_row_ptr_;
    DeviceBuffer<uint32_t> x_col_idx_;
    DeviceBuffer<float> x_val_;
    // second staging set + copy stream: predict_csr uploads sub-tile t+1 while sub-tile t is being scored
    DeviceBuffer<uint64_t> x2_row_ptr_;
    DeviceBuffer<uint32_t> x2_col_idx_;
    DeviceBuffer<float> x2_val_;
    cudaStream_t copy_stream_ = nullptr;
    cudaEvent_t up_ev_[2] = {nullptr, nullptr};   // staging set uploaded
    cudaEvent_t use_ev_[2] = {nullptr, nullptr};  // staging set consumed by the score kernels
    bool pipeline_uploads_ = true;
    PinnedBuffer<uint32_t> beam_id_host_;   // single-layer entry point: the given beam, staged per tile
    PinnedBuffer<float> beam_val_host_;
    PinnedBuffer<uint32_t> beam_cnt_host_;
    uint32_t pipeline_parts_ = 4;
    QueryDev resident_{};
    bool has_resident_ = false;
    DeviceBuffer<uint32_t> res_ids_dev_;
    DeviceBuffer<float> res_vals_dev_;
    DeviceBuffer<uint32_t> res_cnt_dev_;
    uint32_t res_rows_ = 0, res_stride_ = 0;
    DeviceBuffer<unsigned long long> shard_keys_;  // local top-k of an index-sharded run before packing
    DeviceBuffer<uint32_t> shard_ids_, shard_cnt_;
    DeviceBuffer<float> shard_vals_;
    unsigned long long* ext_keys_ = nullptr;  // caller-owned leaf outputs of an index-sharded run
    uint32_t* ext_ids_ = nullptr;
    float* ext_vals_ = nullptr;
    uint32_t* ext_cnt_ = nullptr;

    // host result staging (pinned)
    PinnedBuffer<uint32_t> out_ids_;
    PinnedBuffer<float> out_vals_;
    PinnedBuffer<uint32_t> out_cnt_;

    bool profile_ = false;
    bool no_query_warp_ = false;
    bool force_query_warp_ = false;
    bool no_topk_filter_ = false;
    bool chunk_major_ = true;   // chunk-major scoring wherever cm_plan() finds it eligible (kernel mode 6 switches it off)
    bool cm_force_ = false;     // kernel mode 5
    bool cmg_ = false;          // image-less lane-per-pair kernel: opt-in (kernel modes 8, 9, 10); measured slower, DESIGN.md 3.3
    bool cmg_all_ = false;      // kernel modes 8, 10
    bool cm_image_ = true;      // staged-image chunk-major kernel (kernel mode 10 switches it off)
    uint32_t n_sm_ = 148;
    DeviceBuffer<uint32_t> cm_slot_pos_, cm_count_, cm_bucket_ptr_, cm_item_ptr_, cm_pair_q_, cm_pair_pos_;
    bool force_block_topk_ = false;  // A/B switch: first-generation kernels (row-list streaming + block-wide sort)
    std::vector<XLinearLayerProfile> layer_profile_;
    std::vector<XLinearStats> layer_stats_;
    uint64_t launches_ = 0;
    cudaEvent_t ev_[4] = {nullptr, nullptr, nullptr, nullptr};
};

}  // namespace pb200
