// extern "C" boundary of libpecos_b200_float32.so: XR-Linear entry points + pb200_* additions.
// Signatures mirror pecos/core/libpecos.cpp:116-176 (see include/pecos_b200.h for the per-symbol citations).
#include "../../include/pecos_b200.h"

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <exception>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_set>
#include <vector>

#include "hnsw_engine.h"
#include "xlinear_engine.h"

namespace {

std::atomic<int> g_device{0};

[[noreturn]] void die(const char* where, const char* what) {
    std::fprintf(stderr, "pecos_b200 fatal error in %s: %s\n", where, what);
    std::fflush(stderr);
    std::abort();
}

#define PB200_API_BEGIN try {
#define PB200_API_END(name)                                  \
    }                                                        \
    catch (const std::exception& e) { die(name, e.what()); } \
    catch (...) { die(name, "unknown exception"); }

// In-library multi-GPU query fan-out (the reference's analogue is the OpenMP loop over work items,
// pecos/core/xmc/inference.hpp:969-1005): PB200_DEVICES="0,1,2,3" | "all" at LOAD time puts a replica of the model on every
// listed device; a predict call then splits its rows into contiguous blocks balanced by nnz, one host thread + stream per
// device, and the per-device results are concatenated in row order into the single pred_alloc buffers.  A device may be
// listed more than once (two engines on one GPU: how the single-GPU tests emulate a world of 2).  Unset: one engine on the
// device chosen by pb200_set_device.
std::vector<int> device_list() {
    std::vector<int> out;
    const char* env = std::getenv("PB200_DEVICES");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        throw std::runtime_error("no CUDA device visible: pecos_b200 has no CPU fallback");
    if (!env || !*env) { out.push_back(g_device.load()); return out; }
    const std::string v(env);
    if (v == "all") { for (int i = 0; i < n; ++i) out.push_back(i); return out; }
    size_t pos = 0;
    while (pos <= v.size()) {
        const size_t comma = v.find(',', pos);
        const std::string tok = v.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos);
        if (!tok.empty()) {
            char* end = nullptr;
            const long d = std::strtol(tok.c_str(), &end, 10);
            if (*end != 0 || d < 0 || d >= n) throw std::runtime_error("PB200_DEVICES: bad device id '" + tok + "'");
            out.push_back(static_cast<int>(d));
        }
        if (comma == std::string::npos) break;
        pos = comma + 1;
    }
    if (out.empty()) out.push_back(g_device.load());
    return out;
}

// runs fn(i) for i in [0, n) on n host threads (fn(0) on the caller); rethrows the first exception
template <typename F>
void fan_out(size_t n, F&& fn) {
    if (n <= 1) { if (n == 1) fn(0); return; }
    std::vector<std::exception_ptr> err(n);
    std::vector<std::thread> th;
    th.reserve(n - 1);
    for (size_t i = 1; i < n; ++i)
        th.emplace_back([&, i] { try { fn(i); } catch (...) { err[i] = std::current_exception(); } });
    try { fn(0); } catch (...) { err[0] = std::current_exception(); }
    for (auto& t : th) t.join();
    for (auto& e : err) if (e) std::rethrow_exception(e);
}

struct XLinearHandle {
    std::vector<std::unique_ptr<pb200::XLinearEngine>> engines;  // [0] = primary; > 1: replicas for the query fan-out
    // The reference's handles are immutable after load and may be shared between threads (ctypes drops the GIL during a
    // call).  Ours own device workspaces, so calls on ONE handle are serialised; different handles run concurrently.
    std::mutex mu;
};

std::mutex& mutex_of(void* ptr) {
    if (!ptr) throw std::runtime_error("null model handle");
    return static_cast<XLinearHandle*>(ptr)->mu;
}
#define PB200_LOCK_XL(ptr) std::lock_guard<std::mutex> pb200_handle_lock(mutex_of(ptr));

pb200::XLinearEngine& engine_of(void* ptr) {
    if (!ptr) throw std::runtime_error("null model handle");
    return *static_cast<XLinearHandle*>(ptr)->engines.at(0);
}

void emit_results(const std::vector<pb200::XLinearEngine::Result>& parts, py_sparse_allocator_t pred_alloc) {
    // create_pycsr contract (pecos/core/utils/matrix.hpp:300-316): one allocator call, then fill the three arrays.
    uint64_t nnz = 0, rows = 0;
    bool all_full = true;
    for (const auto& r : parts) {
        uint64_t part = 0;
        for (uint32_t i = 0; i < r.rows; ++i) part += r.cnt[i];
        all_full = all_full && (part == static_cast<uint64_t>(r.rows) * r.stride);
        nnz += part;
        rows += r.rows;
    }
    uint32_t* indices = nullptr;
    uint64_t* indptr = nullptr;
    float* data = nullptr;
    pred_alloc(false, rows, parts.empty() ? 0 : parts[0].out_cols, nnz, &indices, &indptr, &data);
    if (!indptr || (nnz && (!indices || !data))) throw std::runtime_error("result allocator returned null buffers");
    uint64_t w = 0, row = 0;
    indptr[0] = 0;
    for (const auto& r : parts) {
        if (all_full) {
            // every row is full: the fixed-stride device layout already is the CSR payload
            const uint64_t n = static_cast<uint64_t>(r.rows) * r.stride;
            if (n) {
                std::memcpy(indices + w, r.ids, n * sizeof(uint32_t));
                std::memcpy(data + w, r.vals, n * sizeof(float));
            }
            for (uint32_t i = 0; i < r.rows; ++i) indptr[row + i + 1] = w + static_cast<uint64_t>(i + 1) * r.stride;
            w += n;
        } else {
            for (uint32_t i = 0; i < r.rows; ++i) {
                const uint32_t c = r.cnt[i];
                std::memcpy(indices + w, r.ids + static_cast<uint64_t>(i) * r.stride, c * sizeof(uint32_t));
                std::memcpy(data + w, r.vals + static_cast<uint64_t>(i) * r.stride, c * sizeof(float));
                w += c;
                indptr[row + i + 1] = w;
            }
        }
        row += r.rows;
    }
}

void emit_result(const pb200::XLinearEngine::Result& r, py_sparse_allocator_t pred_alloc) {
    emit_results(std::vector<pb200::XLinearEngine::Result>{r}, pred_alloc);
}

void* make_engine(std::unique_ptr<pb200::XLinearHostModel> host, bool allow_replicas = true) {
    std::vector<int> devs = device_list();
    if (!allow_replicas || host->shard_world > 1) devs.resize(1);
    auto h = std::make_unique<XLinearHandle>();
    h->engines.resize(devs.size());
    // the engine consumes (and frees) the host arrays while uploading: every replica gets its own copy
    std::vector<std::unique_ptr<pb200::XLinearHostModel>> copies(devs.size());
    for (size_t i = 1; i < devs.size(); ++i) copies[i] = std::make_unique<pb200::XLinearHostModel>(*host);
    copies[0] = std::move(host);
    fan_out(devs.size(), [&](size_t i) {
        h->engines[i] = std::make_unique<pb200::XLinearEngine>(std::move(copies[i]), devs[i]);
    });
    return h.release();
}

// rows [0, rows) cut into n contiguous blocks of (nearly) equal nnz (pecos_b200.distributed.split_rows_by_nnz)
std::vector<uint32_t> split_rows_by_nnz(const uint64_t* row_ptr, uint32_t rows, size_t n) {
    std::vector<uint32_t> cut(n + 1, rows);
    cut[0] = 0;
    const uint64_t base = row_ptr[0], total = row_ptr[rows] - base;
    for (size_t i = 1; i < n; ++i) {
        const uint64_t target = base + total * i / n;
        const uint64_t* p = std::lower_bound(row_ptr, row_ptr + rows + 1, target);
        uint32_t r = static_cast<uint32_t>(p - row_ptr);
        if (total == 0) r = static_cast<uint32_t>(static_cast<uint64_t>(rows) * i / n);
        cut[i] = std::max(cut[i - 1], std::min(r, rows));
    }
    return cut;
}

constexpr uint32_t kFanOutMinRows = 256;  // below this many rows per device a single engine serves the call

pb200::DeviceBuffer<unsigned char>* g_flush_buf = nullptr;
std::mutex g_flush_mutex;

}  // namespace

extern "C" {

// ------------------------------------------------ XR-Linear ------------------------------------------------------
void* c_xlinear_load_model_from_disk(const char* model_path) {
    PB200_API_BEGIN
    return make_engine(pb200::load_xlinear_npz_model(model_path, pb200::LT_BINARY_SEARCH_CHUNKED));
    PB200_API_END("c_xlinear_load_model_from_disk")
}

void* c_xlinear_load_model_from_disk_ext(const char* model_path, int weight_matrix_type) {
    PB200_API_BEGIN
    return make_engine(pb200::load_xlinear_npz_model(model_path, weight_matrix_type));
    PB200_API_END("c_xlinear_load_model_from_disk_ext")
}

void* c_xlinear_load_mmap_model_from_disk(const char* model_path, const bool lazy_load) {
    PB200_API_BEGIN
    return make_engine(pb200::load_xlinear_mmap_model(model_path, lazy_load));
    PB200_API_END("c_xlinear_load_mmap_model_from_disk")
}

void c_xlinear_compile_mmap_model(const char* model_path, const char* mmap_model_path) {
    // host-only (no CUDA calls): npz model folder -> the reference's mmap format (libpecos.cpp:133-138)
    PB200_API_BEGIN
    auto host = pb200::load_xlinear_npz_model(model_path, pb200::LT_BINARY_SEARCH_CHUNKED);
    pb200::write_xlinear_mmap_model(*host, mmap_model_path);
    PB200_API_END("c_xlinear_compile_mmap_model")
}

void c_mlmodel_compile_mmap_model(const char* model_path, const char* mmap_model_path) {
    // host-only (no CUDA calls): one npz layer folder -> the reference's single-layer mmap format (libpecos.cpp:32-36)
    PB200_API_BEGIN
    pb200::compile_mlmodel_mmap(model_path, mmap_model_path);
    PB200_API_END("c_mlmodel_compile_mmap_model")
}

void c_xlinear_destruct_model(void* ptr) {
    PB200_API_BEGIN
    delete static_cast<XLinearHandle*>(ptr);
    PB200_API_END("c_xlinear_destruct_model")
}

uint32_t c_xlinear_get_int_attr(void* ptr, const char* attr) {
    PB200_API_BEGIN
    const auto& m = engine_of(ptr).host();
    if (std::strcmp(attr, "depth") == 0) return m.depth();
    if (std::strcmp(attr, "nr_features") == 0) return m.nr_features();
    if (std::strcmp(attr, "nr_labels") == 0) return m.nr_labels();
    if (std::strcmp(attr, "nr_codes") == 0) return m.nr_codes();
    throw std::runtime_error(std::string(attr) + " is not implemented in get_int_attr.");
    PB200_API_END("c_xlinear_get_int_attr")
}

int c_xlinear_get_layer_type(void* ptr, int layer_depth) {
    PB200_API_BEGIN
    const auto& m = engine_of(ptr).host();
    if (layer_depth < 0 || static_cast<uint32_t>(layer_depth) >= m.depth()) throw std::runtime_error("layer_depth out of range");
    return m.layer_type;
    PB200_API_END("c_xlinear_get_layer_type")
}

void c_xlinear_predict_csr_f32(void* ptr, const ScipyCsrF32* X, const uint32_t overridden_beam_size,
                               const char* overridden_post_processor_str, const uint32_t overridden_only_topk,
                               const int threads, py_sparse_allocator_t pred_alloc) {
    (void)threads;
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    auto& H = *static_cast<XLinearHandle*>(ptr);
    size_t n = H.engines.size();
    if (static_cast<uint64_t>(X->rows) < static_cast<uint64_t>(kFanOutMinRows) * n) n = 1;
    const auto cut = split_rows_by_nnz(X->row_ptr, X->rows, n);
    std::vector<pb200::XLinearEngine::Result> parts(n);
    fan_out(n, [&](size_t i) {
        parts[i] = H.engines[i]->predict_csr(X->row_ptr + cut[i], X->col_idx, X->val, cut[i + 1] - cut[i], X->cols,
                                             overridden_beam_size, overridden_post_processor_str, overridden_only_topk);
    });
    emit_results(parts, pred_alloc);
    PB200_API_END("c_xlinear_predict_csr_f32")
}

void c_xlinear_predict_drm_f32(void* ptr, const ScipyDrmF32* X, const uint32_t overridden_beam_size,
                               const char* overridden_post_processor_str, const uint32_t overridden_only_topk,
                               const int threads, py_sparse_allocator_t pred_alloc) {
    (void)threads;
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    auto& H = *static_cast<XLinearHandle*>(ptr);
    if (X->cols != engine_of(ptr).host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    size_t n = H.engines.size();
    if (static_cast<uint64_t>(X->rows) < static_cast<uint64_t>(kFanOutMinRows) * n) n = 1;
    std::vector<pb200::XLinearEngine::Result> parts(n);
    fan_out(n, [&](size_t i) {
        const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(X->rows) * i / n);
        const uint32_t r1 = static_cast<uint32_t>(static_cast<uint64_t>(X->rows) * (i + 1) / n);
        parts[i] = H.engines[i]->predict_drm(X->val + static_cast<uint64_t>(r0) * X->cols, r1 - r0, X->cols, overridden_beam_size,
                                             overridden_post_processor_str, overridden_only_topk);
    });
    emit_results(parts, pred_alloc);
    PB200_API_END("c_xlinear_predict_drm_f32")
}

// ------------------------------------------------ selected outputs -----------------------------------------------
}  // extern "C"

namespace {

void emit_selected(const pb200::XLinearEngine::SelectedResult& r, py_sparse_allocator_t pred_alloc) {
    uint32_t* indices = nullptr;
    uint64_t* indptr = nullptr;
    float* data = nullptr;
    const uint64_t nnz = r.indptr.empty() ? 0 : r.indptr.back();
    pred_alloc(false, r.rows, r.cols, nnz, &indices, &indptr, &data);
    if (!indptr || (nnz && (!indices || !data))) throw std::runtime_error("result allocator returned null buffers");
    std::memcpy(indptr, r.indptr.data(), r.indptr.size() * sizeof(uint64_t));
    if (nnz) {
        std::memcpy(indices, r.indices.data(), nnz * sizeof(uint32_t));
        std::memcpy(data, r.data.data(), nnz * sizeof(float));
    }
}

void predict_selected(void* ptr, const ScipyCsrF32* Xs, const ScipyDrmF32* Xd, const ScipyCsrF32* sel, const char* pp,
                      py_sparse_allocator_t pred_alloc) {
    PB200_LOCK_XL(ptr)
    auto& eng = engine_of(ptr);
    const uint32_t rows = Xs ? Xs->rows : Xd->rows;
    const uint32_t cols = Xs ? Xs->cols : Xd->cols;
    if (!sel) throw std::runtime_error("selected_outputs_csr is required");
    if (sel->rows != rows) throw std::runtime_error("Instance dimension of query and selected output matrix do not match");
    if (Xd && cols != eng.host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    auto r = eng.predict_selected(Xs ? Xs->row_ptr : nullptr, Xs ? Xs->col_idx : nullptr, Xs ? Xs->val : nullptr,
                                  Xd ? Xd->val : nullptr, rows, cols, sel->row_ptr, sel->col_idx, sel->cols, pp);
    emit_selected(r, pred_alloc);
}

}  // namespace

extern "C" {

void c_xlinear_predict_on_selected_outputs_csr_f32(void* ptr, const ScipyCsrF32* X, const ScipyCsrF32* selected_outputs_csr,
                                                   const char* overridden_post_processor_str, const int threads,
                                                   py_sparse_allocator_t pred_alloc) {
    (void)threads;
    PB200_API_BEGIN
    predict_selected(ptr, X, nullptr, selected_outputs_csr, overridden_post_processor_str, pred_alloc);
    PB200_API_END("c_xlinear_predict_on_selected_outputs_csr_f32")
}

void c_xlinear_predict_on_selected_outputs_drm_f32(void* ptr, const ScipyDrmF32* X, const ScipyCsrF32* selected_outputs_csr,
                                                   const char* overridden_post_processor_str, const int threads,
                                                   py_sparse_allocator_t pred_alloc) {
    (void)threads;
    PB200_API_BEGIN
    predict_selected(ptr, nullptr, X, selected_outputs_csr, overridden_post_processor_str, pred_alloc);
    PB200_API_END("c_xlinear_predict_on_selected_outputs_drm_f32")
}

// ------------------------------------------------ single layer (python chain) ------------------------------------
}  // extern "C"

namespace {

// Cache of one-layer engines.  The python chain re-sends W and C with every call (pecos/xmc/base.py:934-944), and its
// ctypes shim re-creates the index arrays each time, so pointers of col_ptr / row_idx are useless as identity: the key is
// shape + nnz + value pointer + bias + a content fingerprint over strided samples of all five arrays.
struct LayerKey {
    uint32_t w_rows, w_cols, c_rows, c_cols;
    uint64_t w_nnz, c_nnz;
    const float* w_val;
    uint32_t bias_bits;
    uint64_t fingerprint;
    bool operator==(const LayerKey& o) const {
        return w_rows == o.w_rows && w_cols == o.w_cols && c_rows == o.c_rows && c_cols == o.c_cols && w_nnz == o.w_nnz &&
               c_nnz == o.c_nnz && w_val == o.w_val && bias_bits == o.bias_bits && fingerprint == o.fingerprint;
    }
};

inline uint64_t mix64(uint64_t h, uint64_t v) {
    h ^= v + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return h;
}

template <typename T>
uint64_t sample_hash(uint64_t h, const T* a, uint64_t n) {
    if (n == 0) return mix64(h, 0x51ull);
    const uint64_t samples = 4096;
    const uint64_t step = n > samples ? n / samples : 1;
    for (uint64_t i = 0; i < n; i += step) {
        uint64_t v = 0;
        std::memcpy(&v, &a[i], sizeof(T) < 8 ? sizeof(T) : 8);
        h = mix64(h, v + i);
    }
    uint64_t v = 0;
    std::memcpy(&v, &a[n - 1], sizeof(T) < 8 ? sizeof(T) : 8);
    return mix64(h, v);
}

LayerKey make_layer_key(const ScipyCscF32* W, const ScipyCscF32* C, float bias) {
    LayerKey k{};
    k.w_rows = W->rows; k.w_cols = W->cols; k.c_rows = C->rows; k.c_cols = C->cols;
    k.w_nnz = W->col_ptr[W->cols];
    k.c_nnz = C->col_ptr[C->cols];
    k.w_val = W->val;
    std::memcpy(&k.bias_bits, &bias, 4);
    uint64_t h = 0x0123456789ABCDEFull;
    h = sample_hash(h, W->col_ptr, static_cast<uint64_t>(W->cols) + 1);
    h = sample_hash(h, W->row_idx, k.w_nnz);
    h = sample_hash(h, W->val, k.w_nnz);
    h = sample_hash(h, C->col_ptr, static_cast<uint64_t>(C->cols) + 1);
    h = sample_hash(h, C->row_idx, k.c_nnz);
    k.fingerprint = h;
    return k;
}

struct CachedLayer {
    LayerKey key;
    std::shared_ptr<XLinearHandle> handle;
    uint64_t last_use;
};

std::mutex g_layer_cache_mutex;
// heap-allocated and never destroyed: engines must not run CUDA calls from static destructors at process exit
std::vector<CachedLayer>& g_layer_cache = *new std::vector<CachedLayer>();
uint64_t g_layer_clock = 0, g_layer_hits = 0, g_layer_misses = 0;

std::shared_ptr<XLinearHandle> layer_engine(const ScipyCscF32* W, const ScipyCscF32* C, float bias) {
    if (!W || !C) throw std::runtime_error("single layer: W and C are required");
    if (W->cols != C->rows) throw std::runtime_error("single layer: W.cols != C.rows");
    const LayerKey key = make_layer_key(W, C, bias);
    std::lock_guard<std::mutex> lock(g_layer_cache_mutex);
    for (auto& e : g_layer_cache)
        if (e.key == key) { e.last_use = ++g_layer_clock; ++g_layer_hits; return e.handle; }
    ++g_layer_misses;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        throw std::runtime_error("no CUDA device visible: pecos_b200 has no CPU fallback");
    size_t cap = 8;
    if (const char* env = std::getenv("PB200_LAYER_CACHE")) cap = static_cast<size_t>(std::max(1, std::atoi(env)));
    while (g_layer_cache.size() >= cap) {  // evict the least recently used engine (frees its HBM once callers are done)
        size_t victim = 0;
        for (size_t i = 1; i < g_layer_cache.size(); ++i)
            if (g_layer_cache[i].last_use < g_layer_cache[victim].last_use) victim = i;
        g_layer_cache.erase(g_layer_cache.begin() + static_cast<std::ptrdiff_t>(victim));
    }
    const pb200::CscRaw w{W->rows, W->cols, W->col_ptr, W->row_idx, W->val};
    const pb200::CscRaw c{C->rows, C->cols, C->col_ptr, C->row_idx, C->val};
    auto h = std::make_shared<XLinearHandle>();
    h->engines.push_back(std::make_unique<pb200::XLinearEngine>(pb200::make_single_layer_model(w, c, bias), g_device.load()));
    g_layer_cache.push_back(CachedLayer{key, h, ++g_layer_clock});
    return h;
}

void single_layer_predict(const ScipyCsrF32* Xs, const ScipyDrmF32* Xd, const ScipyCsrF32* codes, ScipyCscF32* W,
                          ScipyCscF32* C, const char* pp, uint32_t only_topk, float bias, py_sparse_allocator_t pred_alloc) {
    if (!pp) throw std::runtime_error("single layer: post_processor_str is required");
    const uint32_t rows = Xs ? Xs->rows : Xd->rows;
    const uint32_t cols = Xs ? Xs->cols : Xd->cols;
    auto h = layer_engine(W, C, bias);
    std::lock_guard<std::mutex> lock(h->mu);
    auto& eng = *h->engines[0];
    // MLModel::predict_internal checks (inference.hpp:2041-2051)
    if (codes && codes->rows != rows) throw std::runtime_error("Instance dimension of query and prev_layer_pred matrix do not match");
    if (codes && codes->cols != C->cols) throw std::runtime_error("Label dimension of prev_layer_pred and C matrix do not match");
    if (Xd && cols != eng.host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    auto r = eng.predict_single_layer(Xs ? Xs->row_ptr : nullptr, Xs ? Xs->col_idx : nullptr, Xs ? Xs->val : nullptr,
                                      Xd ? Xd->val : nullptr, rows, cols, codes ? codes->row_ptr : nullptr,
                                      codes ? codes->col_idx : nullptr, codes ? codes->val : nullptr, pp, only_topk);
    emit_result(r, pred_alloc);
}

// c_xlinear_single_layer_predict_on_selected_outputs_* (libpecos.cpp:238-273): one layer handed over by the caller, scores of
// exactly the given (query, label) pairs; prev_layer_pred = csr_codes or all ones (MLModel::predict_on_selected_outputs,
// inference.hpp:2129-2224)
void single_layer_predict_selected(const ScipyCsrF32* Xs, const ScipyDrmF32* Xd, const ScipyCsrF32* sel, const ScipyCsrF32* codes,
                                   ScipyCscF32* W, ScipyCscF32* C, const char* pp, float bias, py_sparse_allocator_t pred_alloc) {
    if (!pp) throw std::runtime_error("single layer: post_processor_str is required");
    if (!sel) throw std::runtime_error("selected_outputs_csr is required");
    const uint32_t rows = Xs ? Xs->rows : Xd->rows;
    const uint32_t cols = Xs ? Xs->cols : Xd->cols;
    auto h = layer_engine(W, C, bias);
    std::lock_guard<std::mutex> lock(h->mu);
    auto& eng = *h->engines[0];
    if (sel->rows != rows) throw std::runtime_error("Instance dimension of query and selected output matrix do not match");
    if (codes && codes->rows != rows) throw std::runtime_error("Instance dimension of query and prev_layer_pred matrix do not match");
    if (codes && codes->cols != C->cols) throw std::runtime_error("Label dimension of prev_layer_pred and C matrix do not match");
    if (Xd && cols != eng.host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    auto r = eng.predict_selected(Xs ? Xs->row_ptr : nullptr, Xs ? Xs->col_idx : nullptr, Xs ? Xs->val : nullptr,
                                  Xd ? Xd->val : nullptr, rows, cols, sel->row_ptr, sel->col_idx, sel->cols, pp,
                                  codes ? codes->row_ptr : nullptr, codes ? codes->col_idx : nullptr, codes ? codes->val : nullptr);
    emit_selected(r, pred_alloc);
}

}  // namespace

extern "C" {

void c_xlinear_single_layer_predict_on_selected_outputs_csr_f32(const ScipyCsrF32* input_x, const ScipyCsrF32* selected_outputs_csr,
                                                                const ScipyCsrF32* csr_codes, ScipyCscF32* W, ScipyCscF32* C,
                                                                const char* post_processor_str, const int num_threads,
                                                                const float bias, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    single_layer_predict_selected(input_x, nullptr, selected_outputs_csr, csr_codes, W, C, post_processor_str, bias, pred_alloc);
    PB200_API_END("c_xlinear_single_layer_predict_on_selected_outputs_csr_f32")
}

void c_xlinear_single_layer_predict_on_selected_outputs_drm_f32(const ScipyDrmF32* input_x, const ScipyCsrF32* selected_outputs_csr,
                                                                const ScipyCsrF32* csr_codes, ScipyCscF32* W, ScipyCscF32* C,
                                                                const char* post_processor_str, const int num_threads,
                                                                const float bias, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    single_layer_predict_selected(nullptr, input_x, selected_outputs_csr, csr_codes, W, C, post_processor_str, bias, pred_alloc);
    PB200_API_END("c_xlinear_single_layer_predict_on_selected_outputs_drm_f32")
}

void c_xlinear_single_layer_predict_csr_f32(const ScipyCsrF32* input_x, const ScipyCsrF32* csr_codes, ScipyCscF32* W,
                                            ScipyCscF32* C, const char* post_processor_str, const uint32_t only_topk,
                                            const int num_threads, const float bias, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    single_layer_predict(input_x, nullptr, csr_codes, W, C, post_processor_str, only_topk, bias, pred_alloc);
    PB200_API_END("c_xlinear_single_layer_predict_csr_f32")
}

void c_xlinear_single_layer_predict_drm_f32(const ScipyDrmF32* input_x, const ScipyCsrF32* csr_codes, ScipyCscF32* W,
                                            ScipyCscF32* C, const char* post_processor_str, const uint32_t only_topk,
                                            const int num_threads, const float bias, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    single_layer_predict(nullptr, input_x, csr_codes, W, C, post_processor_str, only_topk, bias, pred_alloc);
    PB200_API_END("c_xlinear_single_layer_predict_drm_f32")
}

uint32_t pb200_layer_cache_clear(void) {
    PB200_API_BEGIN
    std::lock_guard<std::mutex> lock(g_layer_cache_mutex);
    const uint32_t n = static_cast<uint32_t>(g_layer_cache.size());
    g_layer_cache.clear();
    return n;
    PB200_API_END("pb200_layer_cache_clear")
}

void pb200_layer_cache_info(uint64_t* out) {
    PB200_API_BEGIN
    std::lock_guard<std::mutex> lock(g_layer_cache_mutex);
    out[0] = g_layer_cache.size(); out[1] = g_layer_hits; out[2] = g_layer_misses;
    PB200_API_END("pb200_layer_cache_info")
}

// ------------------------------------------------ single-layer mmap handles (c_mlmodel_*) -------------------------
// pecos/core/libpecos.cpp:37-113: MLModel<csc_t> saved by save_mmap, served here by a one-layer engine.
void* c_mlmodel_load_mmap_model(const char* model_path, const bool lazy_load) {
    PB200_API_BEGIN
    return make_engine(pb200::load_mlmodel_mmap(model_path, lazy_load), false);
    PB200_API_END("c_mlmodel_load_mmap_model")
}

void c_mlmodel_destruct_model(void* ptr) {
    PB200_API_BEGIN
    delete static_cast<XLinearHandle*>(ptr);
    PB200_API_END("c_mlmodel_destruct_model")
}

uint32_t c_mlmodel_get_int_attr(void* ptr, const char* attr) {
    PB200_API_BEGIN
    const auto& m = engine_of(ptr).host();
    // MLModel<csc_t>::label_count() = W.cols: the csc layout is never rearranged, so pruned / permuted trees report every
    // column of W (our chunked layout scores nnz(C) columns; out_cols keeps the reference's count)
    if (std::strcmp(attr, "nr_labels") == 0) return m.layers.back().out_cols;
    if (std::strcmp(attr, "nr_codes") == 0) return m.nr_codes();
    if (std::strcmp(attr, "nr_features") == 0) return m.nr_features();
    throw std::runtime_error(std::string(attr) + " is not implemented in get_int_attr.");
    PB200_API_END("c_mlmodel_get_int_attr")
}

}  // extern "C"

namespace {

void mlmodel_predict(void* ptr, const ScipyCsrF32* Xs, const ScipyDrmF32* Xd, const ScipyCsrF32* codes, const char* pp,
                     uint32_t only_topk, py_sparse_allocator_t pred_alloc) {
    PB200_LOCK_XL(ptr)
    auto& eng = engine_of(ptr);
    const auto& L = eng.host().layers.at(0);
    const uint32_t rows = Xs ? Xs->rows : Xd->rows;
    const uint32_t cols = Xs ? Xs->cols : Xd->cols;
    if (codes && codes->rows != rows) throw std::runtime_error("Instance dimension of query and prev_layer_pred matrix do not match");
    if (codes && codes->cols != L.n_chunks) throw std::runtime_error("Label dimension of prev_layer_pred and C matrix do not match");
    if (Xd && cols != eng.host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    // only_topk_to_use / post_processor_to_use (inference.hpp:2055-2058): the override if given, else the stored value
    const uint32_t k = only_topk > 0 ? only_topk : static_cast<uint32_t>(L.only_topk);
    auto r = eng.predict_single_layer(Xs ? Xs->row_ptr : nullptr, Xs ? Xs->col_idx : nullptr, Xs ? Xs->val : nullptr,
                                      Xd ? Xd->val : nullptr, rows, cols, codes ? codes->row_ptr : nullptr,
                                      codes ? codes->col_idx : nullptr, codes ? codes->val : nullptr, pp, k);
    emit_result(r, pred_alloc);
}

void mlmodel_predict_selected(void* ptr, const ScipyCsrF32* Xs, const ScipyDrmF32* Xd, const ScipyCsrF32* sel,
                              const ScipyCsrF32* codes, const char* pp, py_sparse_allocator_t pred_alloc) {
    PB200_LOCK_XL(ptr)
    auto& eng = engine_of(ptr);
    const auto& L = eng.host().layers.at(0);
    const uint32_t rows = Xs ? Xs->rows : Xd->rows;
    const uint32_t cols = Xs ? Xs->cols : Xd->cols;
    if (!sel) throw std::runtime_error("selected_outputs_csr is required");
    if (sel->rows != rows) throw std::runtime_error("Instance dimension of query and selected output matrix do not match");
    if (codes && codes->rows != rows) throw std::runtime_error("Instance dimension of query and prev_layer_pred matrix do not match");
    if (codes && codes->cols != L.n_chunks) throw std::runtime_error("Label dimension of prev_layer_pred and C matrix do not match");
    if (Xd && cols != eng.host().nr_features()) throw std::runtime_error("dense query width != nr_features");
    auto r = eng.predict_selected(Xs ? Xs->row_ptr : nullptr, Xs ? Xs->col_idx : nullptr, Xs ? Xs->val : nullptr,
                                  Xd ? Xd->val : nullptr, rows, cols, sel->row_ptr, sel->col_idx, sel->cols, pp,
                                  codes ? codes->row_ptr : nullptr, codes ? codes->col_idx : nullptr, codes ? codes->val : nullptr);
    emit_selected(r, pred_alloc);
}

}  // namespace

extern "C" {

void c_mlmodel_predict_csr_f32(void* ptr, const ScipyCsrF32* input_x, const ScipyCsrF32* csr_codes,
                               const char* overridden_post_processor, const uint32_t overridden_only_topk, const int num_threads,
                               py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    mlmodel_predict(ptr, input_x, nullptr, csr_codes, overridden_post_processor, overridden_only_topk, pred_alloc);
    PB200_API_END("c_mlmodel_predict_csr_f32")
}

void c_mlmodel_predict_drm_f32(void* ptr, const ScipyDrmF32* input_x, const ScipyCsrF32* csr_codes,
                               const char* overridden_post_processor, const uint32_t overridden_only_topk, const int num_threads,
                               py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    mlmodel_predict(ptr, nullptr, input_x, csr_codes, overridden_post_processor, overridden_only_topk, pred_alloc);
    PB200_API_END("c_mlmodel_predict_drm_f32")
}

void c_mlmodel_predict_on_selected_outputs_csr_f32(void* ptr, const ScipyCsrF32* input_x, const ScipyCsrF32* selected_outputs_csr,
                                                   const ScipyCsrF32* csr_codes, const char* overridden_post_processor,
                                                   const int num_threads, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    mlmodel_predict_selected(ptr, input_x, nullptr, selected_outputs_csr, csr_codes, overridden_post_processor, pred_alloc);
    PB200_API_END("c_mlmodel_predict_on_selected_outputs_csr_f32")
}

void c_mlmodel_predict_on_selected_outputs_drm_f32(void* ptr, const ScipyDrmF32* input_x, const ScipyCsrF32* selected_outputs_csr,
                                                   const ScipyCsrF32* csr_codes, const char* overridden_post_processor,
                                                   const int num_threads, py_sparse_allocator_t pred_alloc) {
    (void)num_threads;
    PB200_API_BEGIN
    mlmodel_predict_selected(ptr, nullptr, input_x, selected_outputs_csr, csr_codes, overridden_post_processor, pred_alloc);
    PB200_API_END("c_mlmodel_predict_on_selected_outputs_drm_f32")
}

// ------------------------------------------------ additions ------------------------------------------------------
const char* pb200_version(void) { return "pecos_b200 0.1 (sm_100a)"; }

int pb200_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int pb200_set_device(int device) {
    int n = pb200_device_count();
    if (device < 0 || device >= n) return 1;
    g_device.store(device);
    return cudaSetDevice(device) == cudaSuccess ? 0 : 1;
}

int pb200_get_device(void) { return g_device.load(); }

void* pb200_host_alloc(size_t bytes) {
    PB200_API_BEGIN
    void* p = nullptr;
    PB200_CUDA(cudaSetDevice(g_device.load()));
    PB200_CUDA(cudaMallocHost(&p, bytes ? bytes : 1));
    return p;
    PB200_API_END("pb200_host_alloc")
}

void pb200_host_free(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

void pb200_l2_flush(void) {
    PB200_API_BEGIN
    std::lock_guard<std::mutex> lock(g_flush_mutex);
    PB200_CUDA(cudaSetDevice(g_device.load()));
    if (!g_flush_buf) g_flush_buf = new pb200::DeviceBuffer<unsigned char>();
    const uint64_t bytes = 512ull << 20;  // 4x the 126 MB L2
    g_flush_buf->reserve(bytes);
    static int tick = 0;
    PB200_CUDA(cudaMemset(g_flush_buf->get(), (++tick) & 0xFF, bytes));
    PB200_CUDA(cudaDeviceSynchronize());
    PB200_API_END("pb200_l2_flush")
}

void pb200_xlinear_resident_upload_csr(void* ptr, const ScipyCsrF32* X) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    engine_of(ptr).resident_upload_csr(X->row_ptr, X->col_idx, X->val, X->rows, X->cols);
    PB200_API_END("pb200_xlinear_resident_upload_csr")
}

double pb200_xlinear_resident_predict(void* ptr, uint32_t beam, const char* pp, uint32_t topk, int collect_stats) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    return engine_of(ptr).resident_predict(beam, pp, topk, collect_stats != 0);
    PB200_API_END("pb200_xlinear_resident_predict")
}

void pb200_xlinear_resident_fetch(void* ptr, py_sparse_allocator_t pred_alloc) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    emit_result(engine_of(ptr).resident_fetch(), pred_alloc);
    PB200_API_END("pb200_xlinear_resident_fetch")
}

void* pb200_xlinear_load_sharded(const char* model_path, int weight_matrix_type, uint32_t shard_rank, uint32_t shard_world) {
    PB200_API_BEGIN
    if (weight_matrix_type < 0) return make_engine(pb200::load_xlinear_mmap_model(model_path, false, shard_rank, shard_world), false);
    return make_engine(pb200::load_xlinear_npz_model(model_path, weight_matrix_type, shard_rank, shard_world), false);
    PB200_API_END("pb200_xlinear_load_sharded")
}

void pb200_xlinear_get_shard(void* ptr, uint32_t* out) {
    PB200_API_BEGIN
    const auto& m = engine_of(ptr).host();
    out[0] = m.shard_rank; out[1] = m.shard_world; out[2] = m.leaf_chunk_begin; out[3] = m.leaf_chunk_end;
    PB200_API_END("pb200_xlinear_get_shard")
}

uint32_t pb200_xlinear_sharded_local_csr(void* ptr, const ScipyCsrF32* X, uint32_t beam, const char* pp, uint32_t topk,
                                         uint32_t stride_capacity, void* keys_dev, void* ids_dev, void* vals_dev, void* cnt_dev) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    return engine_of(ptr).sharded_local_csr(X->row_ptr, X->col_idx, X->val, X->rows, X->cols, beam, pp, topk, stride_capacity,
                                            static_cast<unsigned long long*>(keys_dev), static_cast<uint32_t*>(ids_dev),
                                            static_cast<float*>(vals_dev), static_cast<uint32_t*>(cnt_dev));
    PB200_API_END("pb200_xlinear_sharded_local_csr")
}

void pb200_xlinear_sharded_merge(void* ptr, uint32_t world, uint32_t rows, uint32_t stride, uint32_t topk, const void* g_keys,
                                 const void* g_ids, const void* g_vals, const void* g_cnt, py_sparse_allocator_t pred_alloc) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    auto r = engine_of(ptr).sharded_merge(world, rows, stride, topk, static_cast<const unsigned long long*>(g_keys),
                                          static_cast<const uint32_t*>(g_ids), static_cast<const float*>(g_vals),
                                          static_cast<const uint32_t*>(g_cnt));
    emit_result(r, pred_alloc);
    PB200_API_END("pb200_xlinear_sharded_merge")
}

uint32_t pb200_xlinear_sharded_local_csr_packed(void* ptr, const ScipyCsrF32* X, uint32_t beam, const char* pp, uint32_t topk,
                                                uint32_t stride_capacity, void* rec_dev) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    return engine_of(ptr).sharded_local_csr_packed(X->row_ptr, X->col_idx, X->val, X->rows, X->cols, beam, pp, topk, stride_capacity, rec_dev);
    PB200_API_END("pb200_xlinear_sharded_local_csr_packed")
}

void pb200_xlinear_sharded_merge_packed(void* ptr, uint32_t world, uint32_t rows, uint32_t stride, uint32_t topk, const void* g_rec,
                                        py_sparse_allocator_t pred_alloc) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    emit_result(engine_of(ptr).sharded_merge_packed(world, rows, stride, topk, g_rec), pred_alloc);
    PB200_API_END("pb200_xlinear_sharded_merge_packed")
}

void pb200_xlinear_set_profile(void* ptr, int on) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    engine_of(ptr).set_profile(on != 0);
    PB200_API_END("pb200_xlinear_set_profile")
}

int pb200_xlinear_set_lookup(void* ptr, int on) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    engine_of(ptr).set_kernel_mode(on);
    return engine_of(ptr).has_feature_maps() ? 1 : 0;
    PB200_API_END("pb200_xlinear_set_lookup")
}

void pb200_xlinear_reset_profile(void* ptr) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    engine_of(ptr).reset_profile();
    PB200_API_END("pb200_xlinear_reset_profile")
}

void pb200_xlinear_get_profile(void* ptr, double* out) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    const auto& p = engine_of(ptr).layer_profile();
    for (size_t d = 0; d < p.size(); ++d) { out[2 * d] = p[d].scores_ms; out[2 * d + 1] = p[d].topk_ms; }
    PB200_API_END("pb200_xlinear_get_profile")
}

void pb200_xlinear_get_kernel_ids(void* ptr, int* out) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    const auto& p = engine_of(ptr).layer_profile();
    for (size_t d = 0; d < p.size(); ++d) { out[2 * d] = p[d].scores_kernel; out[2 * d + 1] = p[d].topk_kernel; }
    PB200_API_END("pb200_xlinear_get_kernel_ids")
}

void pb200_xlinear_get_stats(void* ptr, uint64_t* out) {
    PB200_API_BEGIN
    PB200_LOCK_XL(ptr)
    const auto& s = engine_of(ptr).layer_stats();
    for (size_t d = 0; d < s.size(); ++d) {
        out[7 * d + 0] = s[d].chunks; out[7 * d + 1] = s[d].chunk_rows; out[7 * d + 2] = s[d].matched;
        out[7 * d + 3] = s[d].entries; out[7 * d + 4] = s[d].out_cols; out[7 * d + 5] = s[d].query_nnz;
        out[7 * d + 6] = s[d].beam_out;
    }
    PB200_API_END("pb200_xlinear_get_stats")
}

uint64_t pb200_xlinear_launches(void* ptr) {
    PB200_API_BEGIN
    return engine_of(ptr).launches();
    PB200_API_END("pb200_xlinear_launches")
}

uint32_t pb200_xlinear_replicas(void* ptr) {
    PB200_API_BEGIN
    if (!ptr) throw std::runtime_error("null model handle");
    return static_cast<uint32_t>(static_cast<XLinearHandle*>(ptr)->engines.size());
    PB200_API_END("pb200_xlinear_replicas")
}

uint64_t pb200_xlinear_model_bytes(void* ptr) {
    PB200_API_BEGIN
    return engine_of(ptr).model_bytes();
    PB200_API_END("pb200_xlinear_model_bytes")
}

// host-only inspection (no CUDA calls)
void* pb200_xlinear_host_load(const char* model_path, int kind) {
    PB200_API_BEGIN
    std::unique_ptr<pb200::XLinearHostModel> m =
        kind == 1 ? pb200::load_xlinear_mmap_model(model_path, false)
                  : pb200::load_xlinear_npz_model(model_path, pb200::LT_BINARY_SEARCH_CHUNKED);
    return m.release();
    PB200_API_END("pb200_xlinear_host_load")
}

void* pb200_xlinear_host_from_csc(const ScipyCscF32* W, const ScipyCscF32* C, float bias) {
    PB200_API_BEGIN
    const pb200::CscRaw w{W->rows, W->cols, W->col_ptr, W->row_idx, W->val};
    const pb200::CscRaw c{C->rows, C->cols, C->col_ptr, C->row_idx, C->val};
    return pb200::make_single_layer_model(w, c, bias).release();
    PB200_API_END("pb200_xlinear_host_from_csc")
}

void pb200_xlinear_host_free(void* hptr) { delete static_cast<pb200::XLinearHostModel*>(hptr); }

uint32_t pb200_xlinear_host_depth(void* hptr) { return static_cast<pb200::XLinearHostModel*>(hptr)->depth(); }

void pb200_xlinear_host_layer_dims(void* hptr, uint32_t layer, uint64_t* out) {
    PB200_API_BEGIN
    const auto& L = static_cast<pb200::XLinearHostModel*>(hptr)->layers.at(layer);
    out[0] = L.w_rows; out[1] = L.n_cols; out[2] = L.out_cols; out[3] = L.n_chunks; out[4] = L.c_max;
    out[5] = L.meta.size(); out[6] = L.entries.size(); out[7] = L.label_of_col.size();
    PB200_API_END("pb200_xlinear_host_layer_dims")
}

void pb200_xlinear_host_layer_export(void* hptr, uint32_t layer, void* chunks32, uint32_t* meta, void* entries8,
                                     uint32_t* label_of_col) {
    PB200_API_BEGIN
    const auto& L = static_cast<pb200::XLinearHostModel*>(hptr)->layers.at(layer);
    if (chunks32) std::memcpy(chunks32, L.chunks.data(), L.chunks.size() * sizeof(pb200::ChunkHeader));
    if (meta) std::memcpy(meta, L.meta.data(), L.meta.size() * 4);
    if (entries8) std::memcpy(entries8, L.entries.data(), L.entries.size() * 8);
    if (label_of_col) std::memcpy(label_of_col, L.label_of_col.data(), L.label_of_col.size() * 4);
    PB200_API_END("pb200_xlinear_host_layer_export")
}

}  // extern "C"

// ------------------------------------------------ HNSW ----------------------------------------------------------
namespace {

struct HnswHandle {
    std::string model_dir;                                     // <model>/c_model the index was loaded from (c_ann_hnsw_save copies it)
    std::vector<std::unique_ptr<pb200::HnswEngine>> engines;  // [0] = primary; > 1: PB200_DEVICES replicas (query fan-out)
    std::mutex mu;  // per-warp search scratch lives with the engine: calls on one handle are serialised
};

std::mutex& hnsw_mutex_of(void* ptr) {
    if (!ptr) throw std::runtime_error("null HNSW handle");
    return static_cast<HnswHandle*>(ptr)->mu;
}
#define PB200_LOCK_HNSW(ptr) std::lock_guard<std::mutex> pb200_handle_lock(hnsw_mutex_of(ptr));

struct HnswSearchers {  // the reference hands out a vector<Searcher>; our scratch lives with the engine (per warp)
    HnswHandle* owner;
    uint32_t num_searcher;
};

pb200::HnswEngine& hnsw_of(void* ptr) {
    if (!ptr) throw std::runtime_error("null HNSW handle");
    return *static_cast<HnswHandle*>(ptr)->engines.at(0);
}

void* hnsw_load(const char* model_dir, bool lazy_load, int metric, bool sparse) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0)
        throw std::runtime_error("no CUDA device visible: pecos_b200 has no CPU fallback");
    const std::vector<int> devs = device_list();
    auto h = std::make_unique<HnswHandle>();
    h->engines.resize(devs.size());
    // the host index is a view of the memory-mapped file: every replica maps it again (shared page cache)
    std::vector<std::unique_ptr<pb200::HnswHostIndex>> views(devs.size());
    for (size_t i = 0; i < devs.size(); ++i) views[i] = pb200::load_hnsw_index(model_dir, metric, lazy_load, sparse);
    fan_out(devs.size(), [&](size_t i) { h->engines[i] = std::make_unique<pb200::HnswEngine>(std::move(views[i]), devs[i]); });
    h->model_dir = model_dir;
    return h.release();
}

void hnsw_predict(void* model_ptr, const ScipyDrmF32* pX, uint32_t* ret_idx, float* ret_val, uint32_t efS, uint32_t topk,
                  int metric) {
    PB200_LOCK_HNSW(model_ptr)
    auto& H = *static_cast<HnswHandle*>(model_ptr);
    if (hnsw_of(model_ptr).metric() != metric) throw std::runtime_error("HNSW handle was loaded with a different metric");
    size_t n = H.engines.size();  // replicas: contiguous row blocks, each engine writes its slice of the caller's arrays
    if (static_cast<uint64_t>(pX->rows) < static_cast<uint64_t>(kFanOutMinRows) * n) n = 1;
    fan_out(n, [&](size_t i) {
        const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(pX->rows) * i / n);
        const uint32_t r1 = static_cast<uint32_t>(static_cast<uint64_t>(pX->rows) * (i + 1) / n);
        H.engines[i]->predict(pX->val + static_cast<uint64_t>(r0) * pX->cols, r1 - r0, pX->cols, efS, topk,
                              ret_idx + static_cast<uint64_t>(r0) * topk, ret_val + static_cast<uint64_t>(r0) * topk);
    });
}

// sparse index, csr queries: the same row fan-out; every engine gets its rows of the caller's csr arrays
void hnsw_predict(void* model_ptr, const ScipyCsrF32* pX, uint32_t* ret_idx, float* ret_val, uint32_t efS, uint32_t topk,
                  int metric) {
    PB200_LOCK_HNSW(model_ptr)
    auto& H = *static_cast<HnswHandle*>(model_ptr);
    if (hnsw_of(model_ptr).metric() != metric) throw std::runtime_error("HNSW handle was loaded with a different metric");
    size_t n = H.engines.size();
    if (static_cast<uint64_t>(pX->rows) < static_cast<uint64_t>(kFanOutMinRows) * n) n = 1;
    fan_out(n, [&](size_t i) {
        const uint32_t r0 = static_cast<uint32_t>(static_cast<uint64_t>(pX->rows) * i / n);
        const uint32_t r1 = static_cast<uint32_t>(static_cast<uint64_t>(pX->rows) * (i + 1) / n);
        H.engines[i]->predict_csr(pX->row_ptr + r0, pX->col_idx, pX->val, r1 - r0, pX->cols, efS, topk,
                                  ret_idx + static_cast<uint64_t>(r0) * topk, ret_val + static_cast<uint64_t>(r0) * topk);
    });
}

}  // namespace

extern "C" {

// Handles are library-specific.  Indices TRAINED by the reference (c_ann_hnsw_train_* stays on the reference library) are
// reference handles, yet once the overlay has re-pointed destruct / predict / searchers_* / save at this library the reference's
// Python hands them to us.  Live handles and searcher tokens created HERE are therefore registered; anything else is forwarded
// to the reference's own function, which the overlay registers through pb200_hnsw_set_foreign (without it: a clear fatal error
// instead of undefined behaviour).
typedef void (*hnsw_destruct_fn)(void*);
typedef void* (*hnsw_searchers_create_fn)(void*, uint32_t);
typedef void (*hnsw_predict_fn)(void*, const void*, uint32_t*, float*, uint32_t, uint32_t, int32_t, void*);
typedef void (*hnsw_save_fn)(void*, const char*);
struct HnswForeign {
    hnsw_destruct_fn destruct = nullptr;
    hnsw_searchers_create_fn searchers_create = nullptr;
    hnsw_destruct_fn searchers_destruct = nullptr;
    hnsw_predict_fn predict = nullptr;
    hnsw_save_fn save = nullptr;
};
}  // extern "C"

namespace {
std::mutex g_hnsw_reg_mutex;
std::unordered_set<void*>& g_hnsw_models = *new std::unordered_set<void*>();
std::unordered_set<void*>& g_hnsw_tokens = *new std::unordered_set<void*>();
HnswForeign g_hnsw_foreign[4];  // [metric + 2 * sparse]

bool hnsw_is_ours(void* p, bool token = false) {
    std::lock_guard<std::mutex> lock(g_hnsw_reg_mutex);
    return (token ? g_hnsw_tokens : g_hnsw_models).count(p) != 0;
}
void hnsw_register(void* p, bool token, bool add) {
    std::lock_guard<std::mutex> lock(g_hnsw_reg_mutex);
    auto& s = token ? g_hnsw_tokens : g_hnsw_models;
    if (add) s.insert(p); else s.erase(p);
}
[[noreturn]] void hnsw_foreign_missing(const char* what) {
    throw std::runtime_error(std::string(what) + ": the handle was not created by pecos_b200 (an index trained by the reference "
                             "library?) and no reference functions were registered with pb200_hnsw_set_foreign");
}
void hnsw_save_copy(void* model_ptr, const char* model_dir) {
    // an index loaded here is a view of <dir>/index.mmap_store + config.json written by the reference: saving = copying them
    const std::string src = static_cast<HnswHandle*>(model_ptr)->model_dir, dst(model_dir);
    if (system(("mkdir -p '" + dst + "'").c_str()) != 0) throw std::runtime_error("cannot create " + dst);
    for (const char* f : {"/config.json", "/index.mmap_store"}) {
        std::FILE* in = std::fopen((src + f).c_str(), "rb");
        if (!in) throw std::runtime_error("cannot read " + src + f);
        std::FILE* out = std::fopen((dst + f).c_str(), "wb");
        if (!out) { std::fclose(in); throw std::runtime_error("cannot write " + dst + f); }
        std::vector<char> buf(1 << 20);
        size_t n;
        while ((n = std::fread(buf.data(), 1, buf.size(), in)) > 0) std::fwrite(buf.data(), 1, n, out);
        std::fclose(in);
        std::fclose(out);
    }
}
}  // namespace

extern "C" {

void pb200_hnsw_set_foreign(int metric, void* destruct, void* searchers_create, void* searchers_destruct, void* predict, void* save) {
    if (metric < 0 || metric > 3) return;  // 0 / 1: dense ip / l2; 2 / 3: sparse (csr) ip / l2
    HnswForeign& f = g_hnsw_foreign[metric];
    f.destruct = reinterpret_cast<hnsw_destruct_fn>(destruct);
    f.searchers_create = reinterpret_cast<hnsw_searchers_create_fn>(searchers_create);
    f.searchers_destruct = reinterpret_cast<hnsw_destruct_fn>(searchers_destruct);
    f.predict = reinterpret_cast<hnsw_predict_fn>(predict);
    f.save = reinterpret_cast<hnsw_save_fn>(save);
}

#define PB200_HNSW_API(SUFFIX, METRIC, MAT_T, SPARSE)                                                                              \
    void* c_ann_hnsw_load##SUFFIX(const char* model_dir, const bool lazy_load) {                                        \
        PB200_API_BEGIN                                                                                                 \
        void* h = hnsw_load(model_dir, lazy_load, METRIC, SPARSE);                                                              \
        hnsw_register(h, false, true);                                                                                  \
        return h;                                                                                                       \
        PB200_API_END("c_ann_hnsw_load" #SUFFIX)                                                                        \
    }                                                                                                                   \
    void c_ann_hnsw_destruct##SUFFIX(void* model_ptr) {                                                                 \
        PB200_API_BEGIN                                                                                                 \
        if (!model_ptr) return;                                                                                         \
        if (!hnsw_is_ours(model_ptr)) {                                                                                 \
            if (!g_hnsw_foreign[METRIC + 2 * SPARSE].destruct) hnsw_foreign_missing("c_ann_hnsw_destruct" #SUFFIX);                  \
            g_hnsw_foreign[METRIC + 2 * SPARSE].destruct(model_ptr);                                                                 \
            return;                                                                                                     \
        }                                                                                                               \
        hnsw_register(model_ptr, false, false);                                                                         \
        delete static_cast<HnswHandle*>(model_ptr);                                                                     \
        PB200_API_END("c_ann_hnsw_destruct" #SUFFIX)                                                                    \
    }                                                                                                                   \
    void* c_ann_hnsw_searchers_create##SUFFIX(void* model_ptr, uint32_t num_searcher) {                                 \
        PB200_API_BEGIN                                                                                                 \
        if (!hnsw_is_ours(model_ptr)) {                                                                                 \
            if (!g_hnsw_foreign[METRIC + 2 * SPARSE].searchers_create) hnsw_foreign_missing("c_ann_hnsw_searchers_create" #SUFFIX);  \
            return g_hnsw_foreign[METRIC + 2 * SPARSE].searchers_create(model_ptr, num_searcher);                                    \
        }                                                                                                               \
        void* t = new HnswSearchers{static_cast<HnswHandle*>(model_ptr), num_searcher};                                 \
        hnsw_register(t, true, true);                                                                                   \
        return t;                                                                                                       \
        PB200_API_END("c_ann_hnsw_searchers_create" #SUFFIX)                                                            \
    }                                                                                                                   \
    void c_ann_hnsw_searchers_destruct##SUFFIX(void* searchers_ptr) {                                                   \
        PB200_API_BEGIN                                                                                                 \
        if (!searchers_ptr) return;                                                                                     \
        if (!hnsw_is_ours(searchers_ptr, true)) {                                                                       \
            if (!g_hnsw_foreign[METRIC + 2 * SPARSE].searchers_destruct) hnsw_foreign_missing("c_ann_hnsw_searchers_destruct" #SUFFIX); \
            g_hnsw_foreign[METRIC + 2 * SPARSE].searchers_destruct(searchers_ptr);                                                   \
            return;                                                                                                     \
        }                                                                                                               \
        hnsw_register(searchers_ptr, true, false);                                                                      \
        delete static_cast<HnswSearchers*>(searchers_ptr);                                                              \
        PB200_API_END("c_ann_hnsw_searchers_destruct" #SUFFIX)                                                          \
    }                                                                                                                   \
    void c_ann_hnsw_predict##SUFFIX(void* model_ptr, const MAT_T* pX, uint32_t* ret_idx, float* ret_val,                \
                                    uint32_t efS, uint32_t topk, int32_t threads, void* searchers_ptr) {                \
        PB200_API_BEGIN                                                                                                 \
        if (!hnsw_is_ours(model_ptr)) {                                                                                 \
            if (!g_hnsw_foreign[METRIC + 2 * SPARSE].predict) hnsw_foreign_missing("c_ann_hnsw_predict" #SUFFIX);                    \
            g_hnsw_foreign[METRIC + 2 * SPARSE].predict(model_ptr, pX, ret_idx, ret_val, efS, topk, threads, searchers_ptr);         \
            return;                                                                                                     \
        }                                                                                                               \
        hnsw_predict(model_ptr, pX, ret_idx, ret_val, efS, topk, METRIC);                                               \
        PB200_API_END("c_ann_hnsw_predict" #SUFFIX)                                                                     \
    }                                                                                                                   \
    void c_ann_hnsw_save##SUFFIX(void* model_ptr, const char* model_dir) {                                              \
        PB200_API_BEGIN                                                                                                 \
        if (!hnsw_is_ours(model_ptr)) {                                                                                 \
            if (!g_hnsw_foreign[METRIC + 2 * SPARSE].save) hnsw_foreign_missing("c_ann_hnsw_save" #SUFFIX);                          \
            g_hnsw_foreign[METRIC + 2 * SPARSE].save(model_ptr, model_dir);                                                          \
            return;                                                                                                     \
        }                                                                                                               \
        hnsw_save_copy(model_ptr, model_dir);                                                                           \
        PB200_API_END("c_ann_hnsw_save" #SUFFIX)                                                                        \
    }

PB200_HNSW_API(_drm_ip_f32, pb200::HNSW_IP, ScipyDrmF32, 0)
PB200_HNSW_API(_drm_l2_f32, pb200::HNSW_L2, ScipyDrmF32, 0)
PB200_HNSW_API(_csr_ip_f32, pb200::HNSW_IP, ScipyCsrF32, 1)
PB200_HNSW_API(_csr_l2_f32, pb200::HNSW_L2, ScipyCsrF32, 1)

void pb200_hnsw_resident_upload_csr(void* model_ptr, const ScipyCsrF32* pX) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    hnsw_of(model_ptr).resident_upload_csr(pX->row_ptr, pX->col_idx, pX->val, pX->rows, pX->cols);
    PB200_API_END("pb200_hnsw_resident_upload_csr")
}

void pb200_hnsw_resident_upload(void* model_ptr, const ScipyDrmF32* pX) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    hnsw_of(model_ptr).resident_upload(pX->val, pX->rows, pX->cols);
    PB200_API_END("pb200_hnsw_resident_upload")
}

double pb200_hnsw_resident_predict(void* model_ptr, uint32_t efS, uint32_t topk) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    return hnsw_of(model_ptr).resident_predict(efS, topk);
    PB200_API_END("pb200_hnsw_resident_predict")
}

void pb200_hnsw_resident_fetch(void* model_ptr, uint32_t* ret_idx, float* ret_val) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    hnsw_of(model_ptr).resident_fetch(ret_idx, ret_val);
    PB200_API_END("pb200_hnsw_resident_fetch")
}

int pb200_hnsw_set_stages(void* model_ptr, int stages) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    hnsw_of(model_ptr).set_stages(stages);
    return hnsw_of(model_ptr).stages();
    PB200_API_END("pb200_hnsw_set_stages")
}

void pb200_hnsw_get_counters(void* model_ptr, uint64_t* out) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    auto c = hnsw_of(model_ptr).counters();
    out[0] = c.n_dist; out[1] = c.n_expand; out[2] = c.n_hops; out[3] = c.n_queries;
    PB200_API_END("pb200_hnsw_get_counters")
}

uint64_t pb200_hnsw_sparse_entries(void* model_ptr) {
    PB200_API_BEGIN
    PB200_LOCK_HNSW(model_ptr)
    return hnsw_of(model_ptr).counters().n_entries;
    PB200_API_END("pb200_hnsw_sparse_entries")
}

uint32_t pb200_hnsw_vcap_retries(void* ptr) {
    PB200_API_BEGIN
    return hnsw_of(ptr).vcap_retries();
    PB200_API_END("pb200_hnsw_vcap_retries")
}

uint32_t pb200_hnsw_replicas(void* ptr) {
    PB200_API_BEGIN
    if (!ptr) throw std::runtime_error("null HNSW handle");
    return static_cast<uint32_t>(static_cast<HnswHandle*>(ptr)->engines.size());
    PB200_API_END("pb200_hnsw_replicas")
}

// host-only ingest of an HNSW index (no CUDA calls): validates config.json + index.mmap_store exactly like the loader does and
// reports what it found.  Returns 0, or 1 with the reason on stderr (wrong index type, version, truncated records ...).
int pb200_hnsw_host_info(const char* model_dir, int metric, int sparse, uint64_t* out) {
    try {
        auto ix = pb200::load_hnsw_index(model_dir, metric, false, sparse != 0);
        uint64_t entries = 0, degree_sum = 0;
        for (uint32_t i = 0; i < ix->num_node; ++i) {
            degree_sum += std::min(ix->l0_neighborhood(i)[0], ix->l0_max_degree);
            if (ix->sparse) {
                const float* v; const uint32_t* c;
                const uint32_t len = ix->l0_sparse_row(i, &v, &c);
                for (uint32_t j = 1; j < len; ++j)
                    if (c[j] <= c[j - 1]) throw std::runtime_error("hnsw index: a stored row does not have strictly ascending indices");
                if (len && c[len - 1] >= ix->feat_dim) throw std::runtime_error("hnsw index: a stored index is out of range");
                entries += len;
            } else {
                entries += ix->feat_dim;
            }
        }
        out[0] = ix->num_node; out[1] = ix->feat_dim; out[2] = ix->maxM; out[3] = ix->maxM0; out[4] = ix->max_level;
        out[5] = ix->init_node; out[6] = entries; out[7] = degree_sum;
        return 0;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "pb200_hnsw_host_info: %s\n", e.what());
        return 1;
    }
}

void pb200_hnsw_get_info(void* model_ptr, uint64_t* out) {
    PB200_API_BEGIN
    auto& e = hnsw_of(model_ptr);
    const auto& h = e.host();
    out[0] = h.num_node; out[1] = h.feat_dim; out[2] = h.maxM; out[3] = h.maxM0; out[4] = h.max_level; out[5] = h.init_node;
    out[6] = e.index_bytes(); out[7] = e.launches();
    PB200_API_END("pb200_hnsw_get_info")
}

}  // extern "C"
