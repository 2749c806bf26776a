// Forward: input staging, per-layer launches, CUDA-graph capture / replay of the conv stack.  Included by fvy_api.cu.
namespace fvy {

// ------------------------------------------------------------------------------------------ forward
static int stage_input(fvy_handle* h, const void* images, int dtype, int batch, const void** dev_images) {
    const size_t es = dtype == FVY_F64 ? 8 : (dtype == FVY_U8 ? 1 : 4);
    const size_t bytes = (size_t)batch * h->cfg.net_h * h->cfg.net_w * 3 * es;
    h->last_slot = -1;
    if (is_device_ptr(images)) { *dev_images = images; return FVY_OK; }
    if (h->input_bytes < bytes) {
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->h2d_stream));
        for (int i = 0; i < 2; ++i) {
            if (h->d_input[i]) cudaFree(h->d_input[i]);
            h->d_input[i] = nullptr;
        }
        h->input_bytes = 0;
        for (int i = 0; i < 2; ++i) CUDA_TRY(cudaMalloc(&h->d_input[i], bytes));
        h->input_bytes = bytes;
    }
    const int slot = (int)(h->stage_slot++ & 1u);
    CUDA_TRY(cudaStreamWaitEvent(h->h2d_stream, h->ev_consumed[slot], 0));     // the stem of two calls ago has read this slot
    CUDA_TRY(cudaMemcpyAsync(h->d_input[slot], images, bytes, cudaMemcpyHostToDevice, h->h2d_stream));
    CUDA_TRY(cudaEventRecord(h->ev_ready[slot], h->h2d_stream));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_ready[slot], 0));
    h->last_slot = slot;
    *dev_images = h->d_input[slot];
    return FVY_OK;
}

static int run_layers(fvy_handle* h, int batch, int first, int last) {
    // Tile dependencies pay off where a CTA only gets a few (long) tiles: the per-tile flag check is a global round trip
    // (~0.5 us) on the A producer, the gain is the overlap of one layer's last wave / drain with the next layer's start.
    static const int flags_max_tiles = [] { const char* v = getenv("FVY_FLAGS_MAX_TILES"); return v && *v ? atoi(v) : 8; }();
    std::vector<char> wait_live(h->layers.size(), 0), sig_live(h->layers.size(), 0);
    if (h->flags_live)
        for (size_t i = 0; i < h->layers.size(); ++i) {
            const Layer& C = h->layers[i];
            if (C.wait_on < 0) continue;
            const int m_tiles = (batch * C.p.dom_plane + kBlockM - 1) / kBlockM;
            const int tiles = C.cta2 ? ((m_tiles + 1) / 2) * C.num_n_tiles : m_tiles * C.num_n_tiles;
            const int ctas = std::max(1, std::min(tiles, C.cta2 ? h->num_sms / 2 : h->num_sms));
            if (h->fuse_stem && C.wait_on == 1) continue;          // conv_1 runs inside the fused stem kernel: no counters
            if ((tiles + ctas - 1) / ctas <= flags_max_tiles) { wait_live[i] = 1; sig_live[C.wait_on] = 1; }
        }
    for (int i = first; i < last; ++i) {
        Layer& L = h->layers[i];
        if (h->flags_live && h->use_chain && L.chain >= 0 && L.chain_pos == 0 && first == 0 && last == (int)h->layers.size()) {
            const fvy_handle::Chain& ch = h->chains[L.chain];
            if (int e = launch_chain(h, ch)) return e;
            i += ch.count - 1;
            continue;
        }
        if (L.s.src == -1 && h->fuse_stem && first == 0 && last == (int)h->layers.size()) {
            // conv_0 + conv_1 in one kernel: conv_0's activation never leaves shared memory
            if (!h->cur_img) return fail(FVY_E_STATE, "no input image resident for conv_0");
            FuseParams f = h->fuse;
            f.batch = batch;
            const int items = batch * f.tiles_w * f.segs;
            const int grid = std::min(items, h->num_sms);
            if (h->cur_dtype == FVY_F32)
                stem_conv1_fused_kernel<float><<<grid, kFuseThreads, kFuseSmem, h->stream>>>(h->tmap_w1f, (const float*)h->cur_img, f);
            else if (h->cur_dtype == FVY_U8)
                stem_conv1_fused_kernel<unsigned char><<<grid, kFuseThreads, kFuseSmem, h->stream>>>(h->tmap_w1f, (const unsigned char*)h->cur_img, f);
            else
                stem_conv1_fused_kernel<double><<<grid, kFuseThreads, kFuseSmem, h->stream>>>(h->tmap_w1f, (const double*)h->cur_img, f);
            CUDA_TRY(cudaGetLastError());
            ++h->launches;
            h->stem_phase_valid = false;
            if (h->last_slot >= 0 && !h->capturing) { CUDA_TRY(cudaEventRecord(h->ev_consumed[h->last_slot], h->stream)); h->last_slot = -1; }
            ++i;                   // conv_1 is done as well
            continue;
        }
        if (L.s.src == -1) {       // conv_0: stem_strip_kernel straight from the image
            if (!h->cur_img) return fail(FVY_E_STATE, "no input image resident for conv_0");
            h->stem_phase_valid = true;
            const long long total = (long long)batch * (h->cfg.net_w / 16) * h->cfg.net_h;     // (image, strip, row) triples
            const int blocks = (int)std::min<long long>((total + kStripWarps - 1) / kStripWarps, (long long)h->num_sms * h->stem_blocks_per_sm);
            __nv_bfloat16* dst = (__nv_bfloat16*)L.p.out[0].ptr;
            if (h->cur_dtype == FVY_F32)
                stem_strip_kernel<float><<<blocks, kStripWarps * 32, 0, h->stream>>>((const float*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                     h->d_stem_w2, L.bias, dst);
            else if (h->cur_dtype == FVY_U8)
                stem_strip_kernel<unsigned char><<<blocks, kStripWarps * 32, 0, h->stream>>>((const unsigned char*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w,
                                                                                             h->cfg.max_batch, h->d_stem_w2, L.bias, dst);
            else
                stem_strip_kernel<double><<<blocks, kStripWarps * 32, 0, h->stream>>>((const double*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                      h->d_stem_w2, L.bias, dst);
            CUDA_TRY(cudaGetLastError());
            ++h->launches;
            if (h->last_slot >= 0 && !h->capturing) { CUDA_TRY(cudaEventRecord(h->ev_consumed[h->last_slot], h->stream)); h->last_slot = -1; }
            continue;
        }
        if (h->conv_mode) L.p.out[0].ptr = h->conv_out;          // fvy_conv_run: the caller's output tensor
        else if (!L.s.bn && L.head_slot >= 0)     // head logits of this call's set
            L.p.out[0].ptr = h->logit_set ? h->d_logits_alt[L.head_slot] : h->d_logits[L.head_slot];
        L.p.m_total = batch * L.p.dom_plane;
        L.p.num_m_tiles = (L.p.m_total + kBlockM - 1) / kBlockM;
        static const int nowork = [] { const char* v = getenv("FVY_NOWORK"); return v && *v ? atoi(v) : 0; }();   // profiling aid: launch cost only
        // SMs the persistent conv grids may occupy; the rest is left to the overlapped post-processing of the previous call
        static const int conv_sms_env = [] { const char* v = getenv("FVY_CONV_SMS"); return v && *v ? atoi(v) : 0; }();
        const int conv_sms = conv_sms_env > 0 ? std::min(conv_sms_env, h->num_sms) : h->num_sms;
        int grid;
        if (L.cta2) {
            const int tiles = ((L.p.num_m_tiles + 1) / 2) * L.p.num_n_tiles;
            grid = std::min(2 * tiles, conv_sms & ~1);
        } else {
            grid = std::min(L.p.num_m_tiles * L.p.num_n_tiles, conv_sms * L.occ);
        }
        {   // column-split epilogue (both groups drain every tile): when the tile's K loop hides the drain anyway, or when a CTA
            // only gets a few tiles and the drain of the last one is what the layer waits for
            const int tiles = L.cta2 ? ((L.p.num_m_tiles + 1) / 2) * L.p.num_n_tiles : L.p.num_m_tiles * L.p.num_n_tiles;
            const int ctas = L.cta2 ? grid / 2 : grid;
            static const int split_env = [] { const char* v = getenv("FVY_SPLIT"); return v && *v ? atoi(v) : -1; }();
            L.p.epi_split = split_env >= 0 ? split_env : ((L.BN >= 128 && (L.deep_k || tiles <= 3 * ctas)) ? 1 : 0);
            // Tail split (CTA pairs, FVY_TAIL_SPLIT=1): when the last wave of tiles would occupy at most half of the pairs, each
            // of its tiles is processed as two column halves by two pairs.  Measured: a half tile is latency-bound on the
            // operand ring (6 taps in flight at 256 clk per tap) and takes 0.96 of a full tile's time - 2.5 us off a 26^2
            // layer in isolation, nothing in the chained forward - so it is OFF by default.
            static const int tail_env = [] { const char* v = getenv("FVY_TAIL_SPLIT"); return v && *v ? atoi(v) : 0; }();
            L.p.split_from = tiles;
            if (L.cta2 && tail_env && L.p.epi_split && tiles > ctas) {
                const int full = (tiles / ctas) * ctas, t = tiles - full;
                if (t > 0 && 2 * t <= ctas) L.p.split_from = full;
            }
        }
        {   // tile dependencies are live only inside a whole forward (every producer runs in the same pass)
            L.p.sig_flags = (sig_live[i] && L.signals) ? L.flags : nullptr;
            L.p.wait_flags = nullptr;
            if (wait_live[i]) {
                const Layer& P = h->layers[L.wait_on];
                const bool p_chained = h->use_chain && P.chain >= 0 && first == 0 && last == (int)h->layers.size();   // conv_chain_kernel: always column-split
                const int pgroups = p_chained ? 2 : ((P.p.epi_groups == 2 && P.BN >= 64 && P.p.epi_split != 0) ? 2 : 1);
                L.p.wait_flags = P.flags;
                L.p.wait_expected = P.num_n_tiles * pgroups;
                L.p.wait_margin = L.s.k == 3 ? L.p.dom_w + 1 : 0;
                L.p.wait_blocks = (batch * P.p.dom_plane + 127) / 128;      // row blocks the producer really writes in this call
            }
            if (L.p.sig_flags) L.p.split_from = L.cta2 ? ((L.p.num_m_tiles + 1) / 2) * L.p.num_n_tiles : L.p.num_m_tiles * L.p.num_n_tiles;
        }
        if (nowork) L.p.num_m_tiles = 0;
        if (nowork == 2) L.p.m_total = -1;
        if (int e = launch_conv(h, L, grid)) return e;
    }
    return FVY_OK;
}

static int forward_enqueue(fvy_handle* h, const void* images, int dtype, int batch) {
    if (h->conv_mode) return fail(FVY_E_STATE, "single-convolution handle: use fvy_conv_run");
    if (!h->weights_loaded) return fail(FVY_E_STATE, "fvy_forward before fvy_load_weights");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (dtype != FVY_F32 && dtype != FVY_F64 && dtype != FVY_U8) return fail(FVY_E_INVALID, "dtype %d", dtype);
    const void* dimg = nullptr;
    if (int e = stage_input(h, images, dtype, batch, &dimg)) return e;
    h->cur_img = dimg; h->cur_dtype = dtype;
    const int nl = (int)h->layers.size();
    if (h->use_chain && h->use_flags && h->d_flags)
        if (int e = prepare_chains(h, batch)) return e;
    auto run_all = [&]() -> int {          // one whole forward: the tile-dependency counters start at zero and are live
        if (h->use_flags && h->d_flags) {
            CUDA_TRY(cudaMemsetAsync(h->d_flags, 0, h->flags_bytes, h->stream));
            h->flags_live = true;
        }
        // FVY_TRACE=1 (with FVY_GRAPH=0): %globaltimer milestones of every layer of this forward, printed to stderr
        static const bool trace = getenv("FVY_TRACE") != nullptr;
        unsigned long long* d = nullptr;
        const size_t per = (size_t)h->num_sms * 32;
        if (trace && !h->capturing) {
            CUDA_TRY(cudaMalloc(&d, per * nl * 8));
            CUDA_TRY(cudaMemsetAsync(d, 0, per * nl * 8, h->stream));
            for (int i = 0; i < nl; ++i) h->layers[i].p.dbg = d + per * i;
        }
        const int e = run_layers(h, batch, 0, nl);
        h->flags_live = false;
        if (d) {
            for (int i = 0; i < nl; ++i) h->layers[i].p.dbg = nullptr;
            std::vector<unsigned long long> v(per * nl);
            cudaMemcpyAsync(v.data(), d, per * nl * 8, cudaMemcpyDeviceToHost, h->stream);
            cudaStreamSynchronize(h->stream);
            cudaFree(d);
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < nl; ++i)
                for (int c = 0; c < h->num_sms; ++c) { const unsigned long long t = v[per * i + c * 32 + 16]; if (t) t0 = std::min(t0, t); }
            double prev_end = 0;
            for (int i = 0; i < nl; ++i) {
                unsigned long long lo[3] = {~0ull, ~0ull, ~0ull}, hi[3] = {0, 0, 0};
                const int slot[3] = {16, 19, 22};      // CTA start, first operands landed (leader CTAs), CTA end
                for (int c = 0; c < h->num_sms; ++c)
                    for (int k = 0; k < 3; ++k) { const unsigned long long t = v[per * i + c * 32 + slot[k]]; if (t) { lo[k] = std::min(lo[k], t); hi[k] = std::max(hi[k], t); } }
                if (hi[0] == 0) continue;
                auto us = [&](unsigned long long t) { return ((double)t - (double)t0) / 1e3; };
                const Layer& L = h->layers[i];
                fprintf(stderr, "trace conv_%-4d wait_on=%2d sig=%d | start %8.2f..%8.2f | first operands %8.2f..%8.2f | end %8.2f..%8.2f | since prev end %+7.2f | span %7.2f\n",
                        L.s.idx, L.p.wait_flags ? L.wait_on : -1, L.p.sig_flags ? 1 : 0, us(lo[0]), us(hi[0]), us(lo[1]), us(hi[1]), us(lo[2]), us(hi[2]),
                        us(hi[2]) - prev_end, us(hi[2]) - us(lo[0]));
                prev_end = us(hi[2]);
            }
        }
        return e;
    };
    if (!h->use_graph) return run_all();
    const fvy_handle::GraphKey key{batch, dtype, h->logit_set, dimg};
    auto it = h->graphs.find(key);
    if (it == h->graphs.end()) {
        if (h->graphs.size() >= 16) {          // callers that keep changing device pointers: do not hoard graphs
            for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
            h->graphs.clear(); h->graph_launches.clear();
        }
        cudaGraph_t g = nullptr;
        cudaGraphExec_t ge = nullptr;
        const long long launches0 = h->launches;
        CUDA_TRY(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
        h->capturing = true;
        const int e = run_all();
        h->capturing = false;
        const cudaError_t ce = cudaStreamEndCapture(h->stream, &g);
        if (e) { if (g) cudaGraphDestroy(g); return e; }
        if (ce != cudaSuccess) return fail(FVY_E_CUDA, "graph capture of the conv stack failed: %s", cudaGetErrorString(ce));
        const cudaError_t ie = cudaGraphInstantiate(&ge, g, 0);
        cudaGraphDestroy(g);
        if (ie != cudaSuccess) return fail(FVY_E_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie));
        h->graph_launches[key] = h->launches - launches0;
        h->launches = launches0;
        it = h->graphs.emplace(key, ge).first;
    }
    CUDA_TRY(cudaGraphLaunch(it->second, h->stream));
    h->launches += h->graph_launches[key];
    if (h->last_slot >= 0) { CUDA_TRY(cudaEventRecord(h->ev_consumed[h->last_slot], h->stream)); h->last_slot = -1; }
    return FVY_OK;
}

static int copy_out(fvy_handle* h, const void* dev, void* dst, size_t bytes) {
    if (!dst) return FVY_OK;
    CUDA_TRY(cudaMemcpyAsync(dst, dev, bytes, is_device_ptr(dst) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, h->stream));
    return FVY_OK;
}

}  // namespace fvy
