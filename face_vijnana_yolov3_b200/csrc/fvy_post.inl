// Post-processing host side: scratch plan, decode / NMS / assemble launches, status checks.  Included by fvy_api.cu.
namespace fvy {

static int total_cands(const fvy_handle* h) {
    if (h->cfg.head == FVY_HEAD_FD6) return (h->cfg.net_h / 32) * (h->cfg.net_w / 32);
    int t = 0;
    for (int lvl : {32, 16, 8}) t += 3 * (h->cfg.net_h / lvl) * (h->cfg.net_w / lvl);
    return t;
}

static int build_post(fvy_handle* h) {
    const fvy_config& c = h->cfg;
    if (h->conv_mode) return FVY_OK;          // no decode / NMS behind a single convolution
    const int B = c.max_batch;
    if (c.head == FVY_HEAD_NONE || c.head == FVY_HEAD_YOLO3) {
        int i = 0;
        for (int lvl : {32, 16, 8}) { h->gh[i] = c.net_h / lvl; h->gw[i] = c.net_w / lvl; ++i; }
        h->head_c = 3 * (5 + c.nb_class);
    }
    if (c.head == FVY_HEAD_NONE) {   // post-processing-only handle still stages logits it is given
        for (int i = 0; i < 3; ++i)
            if (int e = dev_alloc(h, (void**)&h->d_logits[i], (size_t)B * h->gh[i] * h->gw[i] * h->head_c * 4, true)) return e;
    }
    h->cap = c.max_cands > 0 ? c.max_cands : total_cands(h);
    h->capP = (h->cap + 63) / 64 * 64;
    h->words = h->capP / 64;
    h->np2max = 64;
    while (h->np2max < h->cap) h->np2max <<= 1;
    h->smem_keys = std::min(h->np2max, 16384);
    const int nc = std::max(1, c.nb_class);
    const size_t n = (size_t)B * h->cap;
    if (int e = dev_alloc(h, (void**)&h->d_nbox, n * 4 * 8, false)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_ibox, n * 16, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_obj, n * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_cls, n * nc * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_cand, n * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_counts, (size_t)B * 4, true)) return e;
    // status word (16 bytes) + the decode kernel's per-image chunk tickets / counts, zeroed together before every decode
    h->status_bytes = 16 + (size_t)B * (1 + kDecodeMaxChunks) * sizeof(int);
    if (int e = dev_alloc(h, (void**)&h->d_status, h->status_bytes, true)) return e;
    if ((total_cands(h) + kDecodeChunk - 1) / kDecodeChunk > kDecodeMaxChunks)
        return fail(FVY_E_INVALID, "network input %dx%d has more candidate slots than the decode kernel's %d chunks cover", c.net_h, c.net_w, kDecodeMaxChunks);
    if (int e = dev_alloc(h, (void**)&h->d_image_hw, (size_t)B * 8, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_order, (size_t)B * h->capP * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_sbox, (size_t)B * h->capP * 16, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_srow, (size_t)B * h->capP * 16, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_sflag, (size_t)B * (h->capP / 32), true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_mask, (size_t)B * h->capP * h->words * 8, false)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_rowflag, (size_t)B * h->words * 8, true)) return e;
    if (h->np2max > h->smem_keys)
        if (int e = dev_alloc(h, (void**)&h->d_gkeys, (size_t)B * h->np2max * 8, false)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_kept, n * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_kept_counts, (size_t)B * 4, true)) return e;
    h->dets_cap = h->cap;
    if (int e = dev_alloc(h, (void**)&h->d_dets, n * sizeof(FvyDet), true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_det_counts, (size_t)B * 4, true)) return e;
    // a function attribute is per device, not per handle: always raise it to the largest key buffer any handle may use
    CUDA_TRY(cudaFuncSetAttribute(sort_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 8));
    CUDA_TRY(cudaFuncSetAttribute(nms_sweep_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sweep_small_smem(kSweepMaxBlocks)));
    return FVY_OK;
}

// ------------------------------------------------------------------------------------------ post
static int check_pp(const fvy_handle* h, const fvy_post_params* pp) {
    if (!pp) return fail(FVY_E_INVALID, "fvy_post_params is NULL");
    if (pp->arith != FVY_ARITH_F64 && pp->arith != FVY_ARITH_F32) return fail(FVY_E_INVALID, "arith %d", pp->arith);
    (void)h;
    return FVY_OK;
}

// logits given by the caller (host or device) or resident; returns device pointers
static int resolve_logits(fvy_handle* h, const float* o0, const float* o1, const float* o2, int batch, const float* dev[3]) {
    const float* in[3] = {o0, o1, o2};
    const int nheads = h->cfg.head == FVY_HEAD_FD6 ? 1 : 3;
    for (int i = 0; i < nheads; ++i) {
        if (in[i] == nullptr) { dev[i] = h->logit_set ? h->d_logits_alt[i] : h->d_logits[i]; continue; }
        if (is_device_ptr(in[i])) { dev[i] = in[i]; continue; }
        const size_t bytes = (size_t)batch * h->gh[i] * h->gw[i] * h->head_c * 4;
        CUDA_TRY(cudaMemcpyAsync(h->d_logits[i], in[i], bytes, cudaMemcpyHostToDevice, h->stream));
        dev[i] = h->d_logits[i];
    }
    return FVY_OK;
}

static int upload_image_hw(fvy_handle* h, const int* image_hw, int batch, const int** dev) {
    if (!image_hw) { *dev = nullptr; return FVY_OK; }
    if (is_device_ptr(image_hw)) { *dev = image_hw; return FVY_OK; }
    CUDA_TRY(cudaMemcpyAsync(h->d_image_hw, image_hw, (size_t)batch * 8, cudaMemcpyHostToDevice, h->ps));
    *dev = h->d_image_hw;
    return FVY_OK;
}

static int decode_enqueue(fvy_handle* h, const float* dev[3], int batch, const fvy_post_params* pp, const int* d_hw, bool want_nbox) {
    CUDA_TRY(cudaMemsetAsync(h->d_status, 0, h->status_bytes, h->ps));
    if (h->cfg.head == FVY_HEAD_FD6) {
        DecodeFd6Args a;
        a.cands = dev[0]; a.grid = h->gh[0]; a.image_size = h->cfg.net_h; a.cell_px = h->cfg.net_h / 13;
        a.face_conf_th = pp->obj_thresh; a.arith = pp->arith; a.cap = h->cap;
        a.ibox = h->d_ibox; a.objness = h->d_obj; a.score = h->d_cls; a.cand = h->d_cand; a.counts = h->d_counts;
        decode_fd6_kernel<<<batch, 512, 0, h->ps>>>(a);
    } else {
        DecodeArgs a;
        for (int i = 0; i < 3; ++i) { a.out[i] = dev[i]; a.gh[i] = h->gh[i]; a.gw[i] = h->gw[i]; }
        a.nb_class = h->cfg.nb_class;
        memcpy(a.anchors, pp->anchors, sizeof(a.anchors));
        a.anchor_mask = pp->anchor_mask; a.obj_thresh = pp->obj_thresh;
        a.net_h = h->cfg.net_h; a.net_w = h->cfg.net_w; a.arith = pp->arith;
        a.image_hw = d_hw; a.cap = h->cap;
        a.nbox = want_nbox ? h->d_nbox : nullptr; a.ibox = h->d_ibox; a.objness = h->d_obj; a.classes = h->d_cls;
        a.cand = h->d_cand; a.counts = h->d_counts; a.status = h->d_status;
        const int chunks = (total_cands(h) + kDecodeChunk - 1) / kDecodeChunk;
        decode_yolo_kernel<<<dim3(chunks, batch), kDecodeThreads, 0, h->ps>>>(a, h->d_status + 4);
    }
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    return FVY_OK;
}

// The greedy sweep: segments of at most kSweepMaxBlocks x 64 boxes (the handle's capacity decides: headline configurations, fd6)
// take nms_sweep_small_kernel (everything the per-block chain needs staged in shared memory), larger ones nms_sweep_kernel.
static int launch_sweep(fvy_handle* h, const SweepArgs& w, int batch, int max_n, cudaStream_t st) {
    static const int force_big = [] { const char* v = getenv("FVY_SWEEP_BIG"); return v && *v ? atoi(v) : 0; }();
    const int nb_max = (std::min(std::min(w.seg_stride, h->cap), std::max(1, max_n)) + 63) / 64;      // max_n: a bound on the boxes per segment
    if (nb_max <= kSweepMaxBlocks && !force_big) {
        nms_sweep_small_kernel<<<batch, kSweepThreads, sweep_small_smem(nb_max), st>>>(w, nb_max);
    } else {
        static const int sweep_threads = [] { const char* v = getenv("FVY_SWEEP_THREADS"); const int t = v && *v ? atoi(v) : 1024; return t >= 128 && t <= 1024 && t % 32 == 0 ? t : 1024; }();
        nms_sweep_kernel<<<batch, sweep_threads, (size_t)h->words * 8, st>>>(w);
    }
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

// NMS over device-resident segments (ibox/classes with stride `seg_stride`, counts on device)
static int nms_enqueue(fvy_handle* h, const int* d_ibox, float* d_cls, const int* d_counts, int batch, int seg_stride, int nb_class,
                       double th, int max_n) {
    for (int c = 0; c < nb_class; ++c) {
        SortArgs s;
        s.ibox = d_ibox; s.classes = d_cls; s.counts = d_counts; s.seg_stride = seg_stride; s.nb_class = nb_class; s.cls = c;
        s.capP = h->capP; s.descending = 1; s.order = h->d_order; s.sbox = h->d_sbox; s.gkeys = h->d_gkeys;
        s.smem_keys = h->smem_keys; s.np2max = h->np2max;
        s.srow = h->d_srow; s.sflag = h->d_sflag; s.rowflag = h->d_rowflag;      // the sort kernel also zeroes the row flags
        sort_scores_kernel<<<batch, 1024, (size_t)h->smem_keys * 8, h->ps>>>(s);
        CUDA_TRY(cudaGetLastError());
        MaskArgs m;
        m.sbox = h->d_sbox; m.counts = d_counts; m.seg_stride = seg_stride; m.batch = batch; m.capP = h->capP; m.words = h->words;
        m.th = th; m.mask = h->d_mask; m.rowflag = h->d_rowflag; m.srow = h->d_srow; m.sflag = h->d_sflag;
        nms_mask_kernel<<<h->num_sms * 16, 64, 0, h->ps>>>(m);
        CUDA_TRY(cudaGetLastError());
        SweepArgs w;
        w.mask = h->d_mask; w.order = h->d_order; w.counts = d_counts; w.seg_stride = seg_stride; w.capP = h->capP; w.words = h->words;
        w.nb_class = nb_class; w.cls = c; w.classes = d_cls; w.rowflag = h->d_rowflag;
        if (int e = launch_sweep(h, w, batch, max_n, h->ps)) return e;
        h->launches += 3;
    }
    return FVY_OK;
}

static int post_enqueue(fvy_handle* h, const float* dev[3], int batch, const fvy_post_params* pp, const int* d_hw, int max_out) {
    if (int e = decode_enqueue(h, dev, batch, pp, d_hw, false)) return e;
    const int nc = h->cfg.head == FVY_HEAD_FD6 ? 1 : h->cfg.nb_class;
    // the most candidates the anchor mask lets an image have (the sweep sizes its shared memory from it)
    int max_n = h->cap;
    if (h->cfg.head != FVY_HEAD_FD6) {
        max_n = 0;
        for (int s = 0; s < 3; ++s) max_n += __builtin_popcount((pp->anchor_mask >> (3 * s)) & 7u) * h->gh[s] * h->gw[s];
        max_n = std::min(max_n, h->cap);
    }
    if (int e = nms_enqueue(h, h->d_ibox, h->d_cls, h->d_counts, batch, h->cap, nc, pp->nms_thresh, max_n)) return e;
    AssembleArgs a;
    a.ibox = h->d_ibox; a.objness = h->d_obj; a.classes = h->d_cls; a.cand = h->d_cand; a.counts = h->d_counts;
    a.seg_stride = h->cap; a.nb_class = nc; a.max_out = max_out;
    a.limit = pp->num_cands > 0 ? std::min(pp->num_cands, max_out) : max_out;
    a.kept_idx = nullptr; a.kept_counts = nullptr; a.dets = h->d_dets; a.det_counts = h->d_det_counts;
    if (h->cfg.head == FVY_HEAD_FD6) {
        SortArgs s;
        s.ibox = h->d_ibox; s.classes = h->d_cls; s.counts = h->d_counts; s.seg_stride = h->cap; s.nb_class = 1; s.cls = 0;
        s.capP = h->capP; s.descending = 0; s.order = h->d_order; s.sbox = nullptr; s.gkeys = h->d_gkeys;
        s.smem_keys = h->smem_keys; s.np2max = h->np2max; s.srow = nullptr; s.sflag = nullptr; s.rowflag = nullptr;
        sort_scores_kernel<<<batch, 1024, (size_t)h->smem_keys * 8, h->ps>>>(s);
        CUDA_TRY(cudaGetLastError());
        assemble_fd6_kernel<<<batch, 512, 0, h->ps>>>(a, h->d_order, h->capP);
        h->launches += 2;
    } else {
        assemble_yolo_kernel<<<batch, 1024, 0, h->ps>>>(a);
        ++h->launches;
    }
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

static int check_post_status(fvy_handle* h, int batch, const int* counts_host) {
    int st = 0;
    CUDA_TRY(cudaMemcpy(&st, h->d_status, 4, cudaMemcpyDeviceToHost));
    if (st & 1) return fail(FVY_E_RANGE, "decoded coordinate outside +-2^30");
    if (counts_host)
        for (int b = 0; b < batch; ++b)
            if (counts_host[b] > h->cap) return fail(FVY_E_CAPACITY, "image %d: %d candidates exceed capacity %d", b, counts_host[b], h->cap);
    return FVY_OK;
}

// Deferred errors of asynchronous calls (see fvy_handle::h_async): waits for the post-processing of logit set `slot` and turns
// its status word / candidate counts into the error the synchronous path would have returned.
static int harvest_async(fvy_handle* h, int slot) {
    if (h->async_batch[slot] == 0) return FVY_OK;
    CUDA_TRY(cudaEventSynchronize(h->ev_post_done[slot]));
    const int batch = h->async_batch[slot];
    h->async_batch[slot] = 0;
    const int* s = h->h_async[slot];
    if (s[0] & 1) return fail(FVY_E_RANGE, "an earlier asynchronous call: decoded coordinate outside +-2^30");
    for (int b = 0; b < batch; ++b)
        if (s[1 + b] > h->cap)
            return fail(FVY_E_CAPACITY, "an earlier asynchronous call: image %d: %d candidates exceed capacity %d", b, s[1 + b], h->cap);
    return FVY_OK;
}
static int harvest_async_all(fvy_handle* h) {
    const int e0 = harvest_async(h, 0);
    const int e1 = e0 ? FVY_OK : harvest_async(h, 1);
    if (e0) h->async_batch[1] = 0;          // one report per synchronisation point
    return e0 ? e0 : e1;
}

}  // namespace fvy
