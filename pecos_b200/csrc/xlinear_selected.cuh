// predict_on_selected_outputs on the device.
//
// Replaces HierarchicalMLModel::predict_on_selected_outputs (pecos/core/xmc/inference.hpp:2507-2571), per layer
// MLModel::predict_on_selected_outputs_internal (:2129-2180) with prolongate_sparse_predictions (:1302-1358); C ABI
// c_xlinear_predict_on_selected_outputs_{csr,drm}_f32 (pecos/core/libpecos.cpp:179-198).
//
// What the reference computes: the scores of exactly the given (query, label) pairs pushed through the hierarchy, no
// top-k.  The selected set of layer l-1 is the SORTED set of parents of layer l's selected set; a row of layer l holds,
// for every entry of the previous layer's row IN ORDER, its children in C's column order that belong to the layer's
// selected set; value = transform(raw score) combined with the parent's value (not at layer 0).
//
// Split used here: everything STRUCTURAL (selected sets, entry order, which candidate position and which parent entry an
// entry reads) depends only on C and the selected pattern, so the host computes it per query; the device does the
// arithmetic with the SAME validated score kernels as beam search -- the previous layer's entry list plays the beam's
// role, so cand[] holds every child of every listed parent in prolongation order -- followed by one small gather kernel
// (transform + combine of the selected candidates).  Raw scores are therefore bit-identical to predict()'s.
//
// This file is included twice by xlinear_engine.cu: once inside its anonymous namespace for the kernel
// (PB200_SELECTED_KERNELS) and once at namespace scope for the engine method (PB200_SELECTED_ENGINE).
#if defined(PB200_SELECTED_KERNELS)

// one CTA per query: its entries [ent_ptr[q], ent_ptr[q + 1]) read candidate ent_pos[e] of the query's row and, when
// `combine`, the value of entry prev_ptr[q] + ent_parent[e] of the previous layer
__global__ void __launch_bounds__(128)
xl_selected_gather_kernel(const float* __restrict__ cand, const uint64_t cand_stride_q, const uint64_t* __restrict__ ent_ptr,
                          const uint32_t* __restrict__ ent_pos, const uint32_t* __restrict__ ent_parent,
                          const uint64_t* __restrict__ prev_ptr, const float* __restrict__ prev_val, float* __restrict__ cur_val,
                          const int pp_kind, const int pp_p, const int combine) {
    const uint32_t q = blockIdx.x;
    const uint64_t b = ent_ptr[q], e = ent_ptr[q + 1];
    const float* cq = cand + static_cast<uint64_t>(q) * cand_stride_q;
    const uint64_t pb = combine ? prev_ptr[q] : 0;
    for (uint64_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        float v = xl_transform(cq[ent_pos[i]], pp_kind, pp_p);
        if (combine) v = xl_combine(v, prev_val[pb + ent_parent[i]], pp_kind);
        cur_val[i] = v;
    }
}

#elif defined(PB200_SELECTED_ENGINE)

namespace {

struct SelLayerIndex {  // label (original numbering of the layer) -> the chunk (= parent node) and column offset that holds it
    std::vector<uint32_t> chunk_of_label;   // 0xFFFFFFFF: the label has no parent (pruned tree)
    std::vector<uint32_t> offset_of_label;
};

SelLayerIndex build_sel_index(const ChunkedLayerHost& L) {
    SelLayerIndex ix;
    ix.chunk_of_label.assign(L.out_cols, 0xFFFFFFFFu);
    ix.offset_of_label.assign(L.out_cols, 0u);
    for (uint32_t p = 0; p < L.n_chunks; ++p) {
        const ChunkHeader& h = L.chunks[p];
        for (uint32_t j = 0; j < h.n_cols; ++j) {
            const uint32_t col = h.col_begin + j;
            const uint32_t label = L.reordered ? L.label_of_col[col] : col;
            if (label < L.out_cols) { ix.chunk_of_label[label] = p; ix.offset_of_label[label] = j; }
        }
    }
    return ix;
}

}  // namespace

XLinearEngine::SelectedResult XLinearEngine::predict_selected(const uint64_t* row_ptr, const uint32_t* col_idx, const float* val,
                                                              const float* dense, uint32_t rows, uint32_t cols,
                                                              const uint64_t* sel_ptr, const uint32_t* sel_idx, uint32_t sel_cols,
                                                              const char* post_processor, const uint64_t* codes_ptr,
                                                              const uint32_t* codes_idx, const float* codes_val) {
    PB200_CUDA(cudaSetDevice(device_));
    // codes_* (single-layer handles only, c_mlmodel_predict_on_selected_outputs_*): the previous layer's prediction; its
    // rows replace the root as the first layer's parent list and its values are combined with the first layer's scores
    const bool have_codes = codes_ptr != nullptr;
    if (have_codes && layers_.size() != 1) throw std::runtime_error("pecos_b200: csr_codes needs a one-layer model");
    const size_t depth = layers_.size();
    const auto& HL = host_->layers;
    if (sel_cols != HL.back().out_cols) throw std::runtime_error("pecos_b200: selected_outputs_csr.cols != nr_labels");
    if (sel_index_.empty()) {
        sel_index_.resize(depth);
        for (size_t d = 0; d < depth; ++d) {
            SelLayerIndex ix = build_sel_index(HL[d]);
            sel_index_[d].chunk_of_label = std::move(ix.chunk_of_label);
            sel_index_[d].offset_of_label = std::move(ix.offset_of_label);
        }
    }
    SelectedResult out;
    out.rows = rows;
    out.cols = sel_cols;
    out.indptr.assign(sel_ptr, sel_ptr + rows + 1);
    const uint64_t sel_base = sel_ptr[0];
    for (auto& v : out.indptr) v -= sel_base;
    const uint64_t total = out.indptr[rows];
    out.indices.assign(total, 0u);
    out.data.assign(total, 0.0f);
    if (rows == 0) return out;

    // ---- structure, per query (host threads): entry lists of every layer
    struct Lists {
        std::vector<uint64_t> ptr;       // [rows + 1]
        std::vector<uint32_t> id;        // label of the entry (original numbering of the layer)
        std::vector<uint32_t> pos;       // candidate position inside the query's row of this layer
        std::vector<uint32_t> parent;    // index of the parent entry inside the query's row of the previous layer
    };
    std::vector<Lists> lists(depth);
    {
        std::vector<std::vector<uint32_t>> q_id(static_cast<size_t>(rows) * depth), q_pos(static_cast<size_t>(rows) * depth),
            q_par(static_cast<size_t>(rows) * depth);
        const unsigned hw = std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
        const unsigned n_thr = static_cast<unsigned>(std::min<uint64_t>(hw, std::max<uint64_t>(1, rows / 64)));
        std::vector<std::thread> pool;
        std::atomic<uint32_t> next{0};
        std::vector<std::exception_ptr> errs(n_thr);
        auto work = [&](unsigned t) {
            try {
                std::vector<std::vector<uint32_t>> sel(depth);
                for (;;) {
                    const uint32_t q0 = next.fetch_add(64);
                    if (q0 >= rows) break;
                    for (uint32_t q = q0; q < std::min(rows, q0 + 64); ++q) {
                        // selected sets, leaf upwards (sorted, unique)
                        sel[depth - 1].assign(sel_idx + sel_ptr[q], sel_idx + sel_ptr[q + 1]);
                        for (uint32_t lab : sel[depth - 1])
                            if (lab >= HL.back().out_cols) throw std::runtime_error("pecos_b200: selected label id out of range");
                        std::sort(sel[depth - 1].begin(), sel[depth - 1].end());
                        const uint32_t n_leaf = static_cast<uint32_t>(sel[depth - 1].size());
                        for (size_t d = depth - 1; d > 0; --d) {
                            auto& up = sel[d - 1];
                            up.clear();
                            for (uint32_t lab : sel[d]) {
                                const uint32_t par = sel_index_[d].chunk_of_label[lab];
                                if (par != 0xFFFFFFFFu) up.push_back(par);
                            }
                            std::sort(up.begin(), up.end());
                            up.erase(std::unique(up.begin(), up.end()), up.end());
                        }
                        (void)n_leaf;
                        // entry lists, root downwards; first parent list: the given codes row, else every code of the first
                        // layer (= the root for a hierarchical model; ones(rows x nr_codes) for a single layer, libpecos.cpp:96-99)
                        std::vector<uint32_t> prev_id;
                        if (have_codes) prev_id.assign(codes_idx + codes_ptr[q], codes_idx + codes_ptr[q + 1]);
                        else { prev_id.resize(HL[0].n_chunks); for (uint32_t p = 0; p < HL[0].n_chunks; ++p) prev_id[p] = p; }
                        for (size_t d = 0; d < depth; ++d) {
                            auto& ids = q_id[static_cast<size_t>(q) * depth + d];
                            auto& pos = q_pos[static_cast<size_t>(q) * depth + d];
                            auto& par = q_par[static_cast<size_t>(q) * depth + d];
                            const auto& L = HL[d];
                            const auto& S = sel[d];
                            uint32_t slot_base = 0;
                            for (uint32_t i = 0; i < prev_id.size(); ++i) {
                                const uint32_t p = prev_id[i];
                                if (p >= L.n_chunks) throw std::runtime_error("pecos_b200: selected outputs: parent id out of range");
                                const ChunkHeader& h = L.chunks[p];
                                for (uint32_t j = 0; j < h.n_cols; ++j) {
                                    const uint32_t col = h.col_begin + j;
                                    const uint32_t label = L.reordered ? L.label_of_col[col] : col;
                                    if (ids.size() >= S.size() || !std::binary_search(S.begin(), S.end(), label)) continue;
                                    ids.push_back(label);
                                    pos.push_back(slot_base + j);
                                    par.push_back(i);
                                }
                                slot_base += h.n_cols;
                            }
                            prev_id = ids;
                        }
                    }
                }
            } catch (...) { errs[t] = std::current_exception(); }
        };
        for (unsigned t = 1; t < n_thr; ++t) pool.emplace_back(work, t);
        work(0);
        for (auto& th : pool) th.join();
        for (auto& e : errs) if (e) std::rethrow_exception(e);
        for (size_t d = 0; d < depth; ++d) {
            auto& Ls = lists[d];
            Ls.ptr.assign(static_cast<size_t>(rows) + 1, 0);
            for (uint32_t q = 0; q < rows; ++q) Ls.ptr[q + 1] = Ls.ptr[q] + q_id[static_cast<size_t>(q) * depth + d].size();
            Ls.id.resize(Ls.ptr[rows]);
            Ls.pos.resize(Ls.ptr[rows]);
            Ls.parent.resize(Ls.ptr[rows]);
            for (uint32_t q = 0; q < rows; ++q) {
                const size_t k = static_cast<size_t>(q) * depth + d;
                std::copy(q_id[k].begin(), q_id[k].end(), Ls.id.begin() + Ls.ptr[q]);
                std::copy(q_pos[k].begin(), q_pos[k].end(), Ls.pos.begin() + Ls.ptr[q]);
                std::copy(q_par[k].begin(), q_par[k].end(), Ls.parent.begin() + Ls.ptr[q]);
            }
        }
    }

    // ---- pseudo plan: the beam entering layer d is the entry list of layer d - 1
    std::vector<LayerPlan> plan(depth);
    for (size_t d = 0; d < depth; ++d) {
        uint32_t b = 1;
        if (d > 0)
            for (uint32_t q = 0; q < rows; ++q) b = std::max<uint32_t>(b, static_cast<uint32_t>(lists[d - 1].ptr[q + 1] - lists[d - 1].ptr[q]));
        else if (have_codes)
            for (uint32_t q = 0; q < rows; ++q) b = std::max<uint32_t>(b, static_cast<uint32_t>(codes_ptr[q + 1] - codes_ptr[q]));
        else
            b = std::max<uint32_t>(1u, HL[0].n_chunks);
        plan[d].b_prev = b;
        plan[d].k = 1;
        plan[d].k_cap = 1;
        plan[d].pp = post_processor ? parse_post_processor(post_processor) : HL[d].post_processor;
        if (b > 32768u) throw std::runtime_error("pecos_b200: selected outputs: more than 32768 selected nodes in one layer of one query");
    }
    uint32_t tile = pick_tile_rows_(plan, rows);
    if (dense) {
        const uint64_t max_dense_rows = std::max<uint64_t>(1, (4ull << 30) / (static_cast<uint64_t>(std::max<uint32_t>(cols, 1u)) * 4));
        tile = static_cast<uint32_t>(std::min<uint64_t>(tile, max_dense_rows));
    }
    ensure_workspace_(plan, tile);
    beam_id_host_.reserve(static_cast<uint64_t>(tile) * beam_stride_ + 1);
    beam_cnt_host_.reserve(static_cast<uint64_t>(tile) + 1);

    DeviceBuffer<uint64_t> d_ptr[2];
    DeviceBuffer<uint32_t> d_pos, d_par;
    DeviceBuffer<float> d_val[2];
    std::vector<uint64_t> rel_ptr;
    std::vector<float> leaf_vals;
    for (uint32_t r0 = 0; r0 < rows; r0 += tile) {
        const uint32_t tr = std::min(tile, rows - r0);
        QueryDev qd{};
        if (dense) {
            x_val_.upload(dense + static_cast<uint64_t>(r0) * cols, static_cast<uint64_t>(tr) * cols, stream_);
            qd = QueryDev{nullptr, nullptr, x_val_.get(), 0, tr, cols, cols};
        } else {
            const uint64_t base = row_ptr[r0], end = row_ptr[r0 + tr];
            x_row_ptr_.upload(row_ptr + r0, static_cast<uint64_t>(tr) + 1, stream_);
            x_col_idx_.upload(col_idx + base, end - base, stream_);
            x_val_.upload(val + base, end - base, stream_);
            qd = QueryDev{x_row_ptr_.get(), x_col_idx_.get(), x_val_.get(), base, tr, cols, max_row_nnz(row_ptr + r0, tr)};
        }
        int cur = 0;  // which d_ptr / d_val set holds the previous layer
        if (have_codes) {  // the given previous prediction plays "layer -1"
            const uint64_t c0 = codes_ptr[r0], c1 = codes_ptr[r0 + tr];
            rel_ptr.resize(static_cast<size_t>(tr) + 1);
            for (uint32_t r = 0; r <= tr; ++r) rel_ptr[r] = codes_ptr[r0 + r] - c0;
            d_ptr[cur].upload(rel_ptr.data(), rel_ptr.size(), stream_);
            d_val[cur].upload(codes_val + c0, c1 - c0, stream_);
            PB200_CUDA(cudaStreamSynchronize(stream_));
        }
        for (size_t d = 0; d < depth; ++d) {
            // beam = the previous layer's entry list (layer 0: the root)
            for (uint32_t r = 0; r < tr; ++r) {
                uint32_t* ids = beam_id_host_.get() + static_cast<uint64_t>(r) * beam_stride_;
                if (d == 0 && have_codes) {
                    const uint64_t b = codes_ptr[r0 + r], n = codes_ptr[r0 + r + 1] - b;
                    std::memcpy(ids, codes_idx + b, n * 4);
                    beam_cnt_host_.get()[r] = static_cast<uint32_t>(n);
                } else if (d == 0) {
                    for (uint32_t p = 0; p < HL[0].n_chunks; ++p) ids[p] = p;
                    beam_cnt_host_.get()[r] = HL[0].n_chunks;
                } else {
                    const auto& P = lists[d - 1];
                    const uint64_t b = P.ptr[r0 + r], n = P.ptr[r0 + r + 1] - b;
                    std::memcpy(ids, P.id.data() + b, n * 4);
                    beam_cnt_host_.get()[r] = static_cast<uint32_t>(n);
                }
            }
            PB200_CUDA(cudaMemcpyAsync(beam_id_[0].get(), beam_id_host_.get(), static_cast<uint64_t>(tr) * beam_stride_ * 4, cudaMemcpyHostToDevice, stream_));
            PB200_CUDA(cudaMemcpyAsync(beam_cnt_[0].get(), beam_cnt_host_.get(), static_cast<uint64_t>(tr) * 4, cudaMemcpyHostToDevice, stream_));
            score_layer_(d, qd, plan[d].b_prev, 0, false);
            const auto& Ls = lists[d];
            const uint64_t e0 = Ls.ptr[r0], e1 = Ls.ptr[r0 + tr];
            rel_ptr.resize(static_cast<size_t>(tr) + 1);
            for (uint32_t r = 0; r <= tr; ++r) rel_ptr[r] = Ls.ptr[r0 + r] - e0;
            const int nxt = cur ^ 1;
            d_ptr[nxt].upload(rel_ptr.data(), rel_ptr.size(), stream_);
            d_pos.upload(Ls.pos.data() + e0, e1 - e0, stream_);
            d_par.upload(Ls.parent.data() + e0, e1 - e0, stream_);
            d_val[nxt].reserve(std::max<uint64_t>(e1 - e0, 1));
            const uint64_t cand_stride_q = static_cast<uint64_t>(plan[d].b_prev) * std::max<uint32_t>(layers_[d].view.c_max, 1u);
            xl_selected_gather_kernel<<<tr, 128, 0, stream_>>>(cand_.get(), cand_stride_q, d_ptr[nxt].get(), d_pos.get(), d_par.get(),
                                                             d_ptr[cur].get(), d_val[cur].get(), d_val[nxt].get(), plan[d].pp.kind,
                                                             plan[d].pp.p, (d > 0 || have_codes) ? 1 : 0);
            PB200_CUDA(cudaGetLastError());
            ++launches_;
            PB200_CUDA(cudaStreamSynchronize(stream_));  // the pinned beam staging area and rel_ptr are refilled next
            cur = nxt;
        }
        // leaf values of this tile -> result rows (the reference copies the selected row's LENGTH; entries it could not reach
        // -- labels without a path to the root -- stay zero, inference.hpp:2560-2568)
        const auto& Lf = lists[depth - 1];
        const uint64_t e0 = Lf.ptr[r0], e1 = Lf.ptr[r0 + tr];
        leaf_vals.resize(e1 - e0);
        if (e1 > e0) PB200_CUDA(cudaMemcpy(leaf_vals.data(), d_val[cur].get(), (e1 - e0) * 4, cudaMemcpyDeviceToHost));
        for (uint32_t r = 0; r < tr; ++r) {
            const uint64_t ob = out.indptr[r0 + r], on = out.indptr[r0 + r + 1] - ob;
            const uint64_t lb = Lf.ptr[r0 + r], ln = Lf.ptr[r0 + r + 1] - lb;
            for (uint64_t i = 0; i < std::min(on, ln); ++i) {
                out.indices[ob + i] = Lf.id[lb + i];
                out.data[ob + i] = leaf_vals[lb - e0 + i];
            }
        }
    }
    return out;
}

#endif
