// libfvy.so - C ABI (include/fvy.h) over the sm_100a kernels.  This translation unit is the whole library: the kernels
// (conv_igemm_sm100.cuh, conv_chain_sm100.cuh, stem_kernel.cuh, postproc_kernels.cuh, prepost_kernels.cuh), the handle (fvy_handle.h),
// the planner (fvy_planner.inl), the forward (fvy_forward.inl) and post-processing (fvy_post.inl) host code, and the entry points below.
// No torch types, no CPU fallback.
#include "fvy_handle.h"
#include "fvy_planner.inl"
#include "fvy_post.inl"
#include "fvy_forward.inl"

using namespace fvy;

// ============================================================================================ C ABI
extern "C" {

#define NO_CONV_HANDLE(h) do { if ((h) && (h)->conv_mode) return fail(FVY_E_STATE, "single-convolution handle: only fvy_conv_set_weights / fvy_conv_run / fvy_destroy apply"); } while (0)

const char* fvy_last_error(void) { return g_err; }
const char* fvy_version(void) { return "fvy 0.1 (sm_100a: tcgen05/TMEM/TMA implicit-GEMM conv, bitmask NMS)"; }

void fvy_destroy(fvy_handle* h) {
    if (!h) return;
    cudaSetDevice(h->cfg.device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->post_stream) cudaStreamSynchronize(h->post_stream);
    if (h->h2d_stream) cudaStreamSynchronize(h->h2d_stream);
    if (h->d2h_stream) cudaStreamSynchronize(h->d2h_stream);
    for (void* p : h->allocs) cudaFree(p);
    if (h->d_lb_src) cudaFree(h->d_lb_src);
    for (int* p : h->h_async) if (p) cudaFreeHost(p);
    for (auto& pr : h->ev_tf) for (cudaEvent_t e : pr) if (e) cudaEventDestroy(e);
    for (auto& pr : h->ev_tp) for (cudaEvent_t e : pr) if (e) cudaEventDestroy(e);
    for (void* p : h->d_input) if (p) cudaFree(p);
    for (auto& e : h->ev) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : {h->ev_ready[0], h->ev_ready[1], h->ev_consumed[0], h->ev_consumed[1], h->ev_post, h->ev_d2h}) if (e) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    if (h->h2d_stream) cudaStreamDestroy(h->h2d_stream);
    if (h->d2h_stream) cudaStreamDestroy(h->d2h_stream);
    if (h->post_stream) cudaStreamDestroy(h->post_stream);
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
    for (cudaEvent_t e : {h->ev_fwd_done[0], h->ev_fwd_done[1], h->ev_post_done[0], h->ev_post_done[1]}) if (e) cudaEventDestroy(e);
    delete h;
}

// Shared tail of fvy_create / fvy_conv_create: device checks, streams and events, plan, scratch.  conv_k > 0 = single-convolution handle.
static int create_impl(const fvy_config* cfg, int conv_cin, int conv_cout, int conv_k, fvy_handle** out, int conv_stride = 1);

int fvy_create(const fvy_config* cfg, fvy_handle** out) {
    if (!cfg || !out) return fail(FVY_E_INVALID, "NULL argument");
    *out = nullptr;
    if (cfg->head != FVY_HEAD_YOLO3 && cfg->head != FVY_HEAD_FD6 && cfg->head != FVY_HEAD_NONE) return fail(FVY_E_INVALID, "head %d", cfg->head);
    if (cfg->net_h <= 0 || cfg->net_w <= 0 || cfg->net_h % 32 || cfg->net_w % 32) return fail(FVY_E_INVALID, "net size %dx%d must be a positive multiple of 32", cfg->net_h, cfg->net_w);
    if (cfg->max_batch < 1 || cfg->max_batch > 1024) return fail(FVY_E_INVALID, "max_batch %d outside [1, 1024]", cfg->max_batch);
    if (cfg->head != FVY_HEAD_FD6 && (cfg->nb_class < 1 || cfg->nb_class > 80)) return fail(FVY_E_INVALID, "nb_class %d outside [1, 80]", cfg->nb_class);
    if (cfg->head == FVY_HEAD_FD6 && cfg->bb_info_c_size != 6) return fail(FVY_E_INVALID, "bb_info_c_size must be 6 (FaceDetector.detect reads channels 0..5)");
    if (cfg->tile_n_max != 0 && cfg->tile_n_max != 32 && cfg->tile_n_max != 64 && cfg->tile_n_max != 128 && cfg->tile_n_max != 256)
        return fail(FVY_E_INVALID, "tile_n_max %d", cfg->tile_n_max);
    return create_impl(cfg, 0, 0, 0, out);
}

static int create_impl(const fvy_config* cfg, int conv_cin, int conv_cout, int conv_k, fvy_handle** out, int conv_stride) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { cudaGetLastError(); return fail(FVY_E_CUDA, "no CUDA device (there is no CPU fallback)"); }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FVY_E_INVALID, "device %d of %d", cfg->device, ndev);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10) return fail(FVY_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", cfg->device, prop.major, prop.minor);
    fvy_handle* h = new fvy_handle();
    h->cfg = *cfg;
    h->num_sms = prop.multiProcessorCount;
    h->conv_mode = conv_k > 0; h->conv_cin = conv_cin; h->conv_cout = conv_cout; h->conv_k = conv_k; h->conv_stride = conv_stride;
    int e = FVY_OK;
    do {
        if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) { e = fail(FVY_E_CUDA, "cudaStreamCreate failed"); break; }
        bool ok = cudaStreamCreateWithFlags(&h->h2d_stream, cudaStreamNonBlocking) == cudaSuccess &&
                  cudaStreamCreateWithFlags(&h->d2h_stream, cudaStreamNonBlocking) == cudaSuccess;
        {   // post-processing stream at the LOWEST priority: it only takes what the conv kernels leave
            int lo = 0, hi = 0;
            cudaDeviceGetStreamPriorityRange(&lo, &hi);
            ok = ok && cudaStreamCreateWithPriority(&h->post_stream, cudaStreamNonBlocking, lo) == cudaSuccess;
            h->ps = h->stream;
            const char* v = getenv("FVY_OVERLAP_POST");
            h->overlap_post = !(v && *v && atoi(v) == 0) && !(cfg->flags & FVY_CFG_NO_OVERLAP_POST);
            const char* g = getenv("FVY_GRAPH");
            h->use_graph = !(g && *g && atoi(g) == 0) && !(cfg->flags & FVY_CFG_NO_GRAPH);
        }
        for (cudaEvent_t* ev : {&h->ev_fwd_done[0], &h->ev_fwd_done[1], &h->ev_post_done[0], &h->ev_post_done[1]})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
        for (auto& ev : h->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        for (cudaEvent_t* ev : {&h->ev_ready[0], &h->ev_ready[1], &h->ev_consumed[0], &h->ev_consumed[1], &h->ev_post, &h->ev_d2h})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { e = fail(FVY_E_CUDA, "cudaEventCreate failed"); break; }
        for (auto& pr : h->ev_tf) for (cudaEvent_t& ev : pr) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        for (auto& pr : h->ev_tp) for (cudaEvent_t& ev : pr) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        for (int*& p : h->h_async)
            ok = ok && cudaHostAlloc((void**)&p, (size_t)(1 + cfg->max_batch) * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
        if (!ok) { e = fail(FVY_E_CUDA, "cudaHostAlloc (async status) failed"); break; }
        if (cfg->head != FVY_HEAD_NONE && (e = build_plan(h))) break;
        if ((e = build_post(h))) break;
        if (cudaStreamSynchronize(h->stream) != cudaSuccess) { e = fail(FVY_E_CUDA, "sync after create failed: %s", cudaGetErrorString(cudaGetLastError())); break; }
    } while (0);
    if (e) { fvy_destroy(h); return e; }
    *out = h;
    return FVY_OK;
}

long long fvy_weight_count(const fvy_handle* h) { return h ? h->weight_count : 0; }

int fvy_load_weights(fvy_handle* h, const float* stream, size_t n_floats) {
    NO_CONV_HANDLE(h);
    if (!h || !stream) return fail(FVY_E_INVALID, "NULL argument");
    if (h->cfg.head == FVY_HEAD_NONE) return fail(FVY_E_STATE, "handle has no network");
    if ((long long)n_floats != h->weight_count) return fail(FVY_E_INVALID, "weight stream has %zu floats, network needs %lld", n_floats, h->weight_count);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    std::vector<uint16_t> wbuf;
    std::vector<float> bbuf;
    for (Layer& L : h->layers) {
        const ConvSpec& s = L.s;
        const float* p = stream + L.stream_off;
        const float *beta = nullptr, *gamma = nullptr, *mean = nullptr, *var = nullptr, *bias = nullptr;
        if (s.bn) { beta = p; gamma = p + s.cout; mean = p + 2 * s.cout; var = p + 3 * s.cout; p += 4 * s.cout; }   // yolov3_detect.py:97-101
        else { bias = p; p += s.cout; }                                                                              // :108-109
        const float* kern = p;   // (Cout, Cin, kh, kw)  :112, :117
        const size_t kdim = (size_t)L.taps * L.cin_pad;
        wbuf.assign((size_t)L.cout_pad * kdim, 0);
        bbuf.assign(L.cout_pad, 0.f);
        const bool stem = s.src == -1;
        for (int o = 0; o < s.cout; ++o) {
            float scale = 1.f;
            if (s.bn) {
                scale = gamma[o] / sqrtf(var[o] + 0.001f);          // BatchNormalization(epsilon=0.001) :212, folded in fp32
                bbuf[o] = beta[o] - mean[o] * scale;
            } else bbuf[o] = bias[o];
            for (int ci = 0; ci < s.cin; ++ci)
                for (int r = 0; r < s.k; ++r)
                    for (int q = 0; q < s.k; ++q) {
                        const float v = kern[(((size_t)o * s.cin + ci) * s.k + r) * s.k + q] * scale;
                        const int qp = L.tap_perm ? (q == 0 ? 0 : (q == 2 ? 1 : 2)) : q;      // stored position of column tap q
                        const size_t kk = stem ? (size_t)(r * 3 + q) * 3 + ci : (size_t)(r * s.k + qp) * L.cin_pad + ci;
                        wbuf[(size_t)o * kdim + kk] = f32_to_bf16_rn(v);
                    }
        }
        if (stem) {      // the same folded weights in stem_strip_kernel's K order: k = 10 r + (3 q + ci)
            std::vector<uint16_t> w2((size_t)32 * 32, 0);
            for (int o = 0; o < s.cout && o < 32; ++o)
                for (int r = 0; r < 3; ++r)
                    for (int j = 0; j < 9; ++j) w2[(size_t)o * 32 + r * 10 + j] = wbuf[(size_t)o * kdim + r * 9 + j];
            CUDA_TRY(cudaMemcpyAsync(h->d_stem_w2, w2.data(), w2.size() * 2, cudaMemcpyHostToDevice, h->stream));
            CUDA_TRY(cudaStreamSynchronize(h->stream));
        }
        CUDA_TRY(cudaMemcpyAsync(L.w, wbuf.data(), wbuf.size() * 2, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(L.bias, bbuf.data(), bbuf.size() * 4, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
    }
    h->weights_loaded = true;
    return FVY_OK;
}

int fvy_forward(fvy_handle* h, const void* images, int dtype, int batch, float* out0, float* out1, float* out2) {
    NO_CONV_HANDLE(h);
    if (!h || !images) return fail(FVY_E_INVALID, "NULL argument");
    if (h->cfg.head == FVY_HEAD_NONE) return fail(FVY_E_STATE, "handle has no network");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // Like the synchronous detect: logit set 0, and nothing of an earlier asynchronous call may still read (or later overwrite) it.
    if (int e = harvest_async_all(h)) return e;
    h->logit_set = 0;
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[0], 0));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[1], 0));
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    if (int e = forward_enqueue(h, images, dtype, batch)) return e;
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    float* outs[3] = {out0, out1, out2};
    const int nheads = h->cfg.head == FVY_HEAD_FD6 ? 1 : 3;
    for (int i = 0; i < nheads; ++i)
        if (int e = copy_out(h, h->d_logits[i], outs[i], (size_t)batch * h->gh[i] * h->gw[i] * h->head_c * 4)) return e;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventElapsedTime(&h->last_fwd_ms, h->ev[0], h->ev[1]));
    return FVY_OK;
}

int fvy_decode(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch, const fvy_post_params* pp,
               const int* image_hw, int cap, double* nbox, int32_t* ibox, float* objness, float* classes, int32_t* cand,
               int32_t* counts) {
    NO_CONV_HANDLE(h);
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    if (int e = check_pp(h, pp)) return e;
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (cap != h->cap) return fail(FVY_E_INVALID, "cap %d must equal the handle's candidate capacity %d", cap, h->cap);
    if (ibox && !image_hw && h->cfg.head != FVY_HEAD_FD6) return fail(FVY_E_INVALID, "ibox needs image_hw");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const float* dev[3] = {nullptr, nullptr, nullptr};
    if (int e = resolve_logits(h, out0, out1, out2, batch, dev)) return e;
    const int* d_hw = nullptr;
    if (int e = upload_image_hw(h, image_hw, batch, &d_hw)) return e;
    if (int e = decode_enqueue(h, dev, batch, pp, d_hw, nbox != nullptr)) return e;
    const int nc = h->cfg.head == FVY_HEAD_FD6 ? 1 : h->cfg.nb_class;
    const size_t n = (size_t)batch * h->cap;
    if (int e = copy_out(h, h->d_nbox, nbox, n * 32)) return e;
    if (int e = copy_out(h, h->d_ibox, ibox, n * 16)) return e;
    if (int e = copy_out(h, h->d_obj, objness, n * 4)) return e;
    if (int e = copy_out(h, h->d_cls, classes, n * nc * 4)) return e;
    if (int e = copy_out(h, h->d_cand, cand, n * 4)) return e;
    std::vector<int> hc(batch);
    CUDA_TRY(cudaMemcpyAsync(hc.data(), h->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    if (counts) {
        if (is_device_ptr(counts)) CUDA_TRY(cudaMemcpy(counts, h->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToDevice));
        else memcpy(counts, hc.data(), (size_t)batch * 4);
    }
    return check_post_status(h, batch, hc.data());
}

int fvy_correct_boxes(fvy_handle* h, const double* nbox, int n, int image_h, int image_w, int net_h, int net_w, int arith,
                      int32_t* ibox) {
    NO_CONV_HANDLE(h);
    if (!h || !nbox || !ibox) return fail(FVY_E_INVALID, "NULL argument");
    if (n < 0 || (size_t)n > (size_t)h->cfg.max_batch * h->cap) return fail(FVY_E_INVALID, "n %d exceeds scratch capacity", n);
    if (n == 0) return FVY_OK;
    if (image_h <= 0 || image_w <= 0 || net_h <= 0 || net_w <= 0) return fail(FVY_E_INVALID, "non-positive size");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaMemsetAsync(h->d_status, 0, 4, h->stream));
    const double* src = nbox;
    if (!is_device_ptr(nbox)) {
        CUDA_TRY(cudaMemcpyAsync(h->d_nbox, nbox, (size_t)n * 32, cudaMemcpyHostToDevice, h->stream));
        src = h->d_nbox;
    }
    correct_boxes_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(src, n, image_h, image_w, net_h, net_w, arith, h->d_ibox, h->d_status);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    if (int e = copy_out(h, h->d_ibox, ibox, (size_t)n * 16)) return e;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return check_post_status(h, 0, nullptr);
}

int fvy_nms(fvy_handle* h, const int32_t* ibox, const int32_t* counts, int batch, int seg_stride, int nb_class, double nms_thresh,
            float* classes, int32_t* kept_idx, int32_t* kept_counts) {
    NO_CONV_HANDLE(h);
    if (!h || !ibox || !counts || !classes) return fail(FVY_E_INVALID, "NULL argument");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (seg_stride < 1 || seg_stride > h->cap) return fail(FVY_E_INVALID, "seg_stride %d outside [1, %d]", seg_stride, h->cap);
    if (nb_class < 1 || nb_class > std::max(1, h->cfg.nb_class)) return fail(FVY_E_INVALID, "nb_class %d exceeds the handle's %d", nb_class, h->cfg.nb_class);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)batch * seg_stride;
    const bool dev_in = is_device_ptr(ibox);
    if (dev_in != is_device_ptr(classes) || dev_in != is_device_ptr(counts)) return fail(FVY_E_INVALID, "ibox/classes/counts must all be host or all be device pointers");
    const int* d_ibox = ibox; float* d_cls = classes; const int* d_counts = counts;
    if (!dev_in) {
        for (int b = 0; b < batch; ++b)
            if (counts[b] < 0 || counts[b] > seg_stride) return fail(FVY_E_INVALID, "counts[%d] = %d outside [0, %d]", b, counts[b], seg_stride);
        CUDA_TRY(cudaMemcpyAsync(h->d_ibox, ibox, n * 16, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_cls, classes, n * nb_class * 4, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaMemcpyAsync(h->d_counts, counts, (size_t)batch * 4, cudaMemcpyHostToDevice, h->stream));
        d_ibox = h->d_ibox; d_cls = h->d_cls; d_counts = h->d_counts;
    }
    CUDA_TRY(cudaEventRecord(h->ev[2], h->stream));
    if (int e = nms_enqueue(h, d_ibox, d_cls, d_counts, batch, seg_stride, nb_class, nms_thresh, seg_stride)) return e;
    if (kept_idx || kept_counts) {
        AssembleArgs a;
        a.ibox = d_ibox; a.objness = nullptr; a.classes = d_cls; a.cand = nullptr; a.counts = d_counts;
        a.seg_stride = seg_stride; a.nb_class = nb_class; a.max_out = 0; a.limit = 0;
        a.kept_idx = h->d_kept; a.kept_counts = h->d_kept_counts; a.dets = nullptr; a.det_counts = nullptr;
        assemble_yolo_kernel<<<batch, 1024, 0, h->stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        ++h->launches;
    }
    CUDA_TRY(cudaEventRecord(h->ev[3], h->stream));
    if (!dev_in) if (int e = copy_out(h, d_cls, classes, n * nb_class * 4)) return e;
    if (int e = copy_out(h, h->d_kept, kept_idx, n * 4)) return e;
    if (int e = copy_out(h, h->d_kept_counts, kept_counts, (size_t)batch * 4)) return e;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaEventElapsedTime(&h->last_post_ms, h->ev[2], h->ev[3]));
    return FVY_OK;
}

int fvy_bbox_iou(fvy_handle* h, const int32_t* a, const int32_t* b, int n, double* out) {
    NO_CONV_HANDLE(h);
    if (!h || !a || !b || !out) return fail(FVY_E_INVALID, "NULL argument");
    if (n < 0 || (size_t)n * 2 > (size_t)h->cfg.max_batch * h->cap) return fail(FVY_E_INVALID, "n %d exceeds scratch capacity", n);
    if (n == 0) return FVY_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    int4* da = reinterpret_cast<int4*>(h->d_ibox);
    int4* db = da + n;
    double* dout = h->d_nbox;
    CUDA_TRY(cudaMemcpyAsync(da, a, (size_t)n * 16, cudaMemcpyDefault, h->stream));
    CUDA_TRY(cudaMemcpyAsync(db, b, (size_t)n * 16, cudaMemcpyDefault, h->stream));
    bbox_iou_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(da, db, n, dout);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDefault, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FVY_OK;
}

int fvy_bbox_iou_fp(fvy_handle* h, const double* a, const double* b, int n, int arith, double* out) {
    NO_CONV_HANDLE(h);
    if (!h || !a || !b || !out) return fail(FVY_E_INVALID, "NULL argument");
    if (arith != FVY_ARITH_F64 && arith != FVY_ARITH_F32) return fail(FVY_E_INVALID, "arith %d", arith);
    if (n < 0 || (size_t)n * 3 > (size_t)h->cfg.max_batch * h->cap) return fail(FVY_E_INVALID, "n %d exceeds scratch capacity", n);
    if (n == 0) return FVY_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    double* da = h->d_nbox; double* db = da + 4 * (size_t)n; double* dout = db + 4 * (size_t)n;       // 9 n doubles <= 4 cap doubles
    CUDA_TRY(cudaMemcpyAsync(da, a, (size_t)n * 32, cudaMemcpyDefault, h->stream));
    CUDA_TRY(cudaMemcpyAsync(db, b, (size_t)n * 32, cudaMemcpyDefault, h->stream));
    bbox_iou_fp_kernel<<<(n + 255) / 256, 256, 0, h->stream>>>(da, db, n, arith, dout);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    CUDA_TRY(cudaMemcpyAsync(out, dout, (size_t)n * 8, cudaMemcpyDefault, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FVY_OK;
}

int fvy_nms_fp(fvy_handle* h, const double* box, const int32_t* counts, int batch, int seg_stride, int nb_class, double nms_thresh,
               int arith, float* classes, int32_t* kept_idx, int32_t* kept_counts) {
    NO_CONV_HANDLE(h);
    if (!h || !box || !counts || !classes) return fail(FVY_E_INVALID, "NULL argument");
    if (arith != FVY_ARITH_F64 && arith != FVY_ARITH_F32) return fail(FVY_E_INVALID, "arith %d", arith);
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (seg_stride < 1 || seg_stride > h->cap) return fail(FVY_E_INVALID, "seg_stride %d outside [1, %d]", seg_stride, h->cap);
    if (nb_class < 1 || nb_class > std::max(1, h->cfg.nb_class)) return fail(FVY_E_INVALID, "nb_class %d exceeds the handle's %d", nb_class, h->cfg.nb_class);
    if (is_device_ptr(box) || is_device_ptr(classes) || is_device_ptr(counts)) return fail(FVY_E_INVALID, "fvy_nms_fp takes host pointers");
    for (int b = 0; b < batch; ++b)
        if (counts[b] < 0 || counts[b] > seg_stride) return fail(FVY_E_INVALID, "counts[%d] = %d outside [0, %d]", b, counts[b], seg_stride);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const size_t n = (size_t)batch * seg_stride;
    CUDA_TRY(cudaMemcpyAsync(h->d_nbox, box, n * 32, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_cls, classes, n * nb_class * 4, cudaMemcpyHostToDevice, h->stream));
    CUDA_TRY(cudaMemcpyAsync(h->d_counts, counts, (size_t)batch * 4, cudaMemcpyHostToDevice, h->stream));
    for (int c = 0; c < nb_class; ++c) {
        SortArgs s;
        s.ibox = nullptr; s.classes = h->d_cls; s.counts = h->d_counts; s.seg_stride = seg_stride; s.nb_class = nb_class; s.cls = c;
        s.capP = h->capP; s.descending = 1; s.order = h->d_order; s.sbox = nullptr; s.gkeys = h->d_gkeys;
        s.smem_keys = h->smem_keys; s.np2max = h->np2max; s.srow = nullptr; s.sflag = nullptr; s.rowflag = nullptr;
        sort_scores_kernel<<<batch, 1024, (size_t)h->smem_keys * 8, h->stream>>>(s);
        CUDA_TRY(cudaGetLastError());
        MaskFpArgs m;
        m.box = h->d_nbox; m.order = h->d_order; m.counts = h->d_counts; m.seg_stride = seg_stride; m.batch = batch; m.capP = h->capP;
        m.words = h->words; m.arith = arith; m.th = nms_thresh; m.mask = h->d_mask; m.rowflag = h->d_rowflag;
        CUDA_TRY(cudaMemsetAsync(h->d_rowflag, 0, (size_t)batch * h->words * 8, h->stream));
        nms_mask_fp_kernel<<<h->num_sms * 8, 64, 0, h->stream>>>(m);
        CUDA_TRY(cudaGetLastError());
        SweepArgs w;
        w.mask = h->d_mask; w.order = h->d_order; w.counts = h->d_counts; w.seg_stride = seg_stride; w.capP = h->capP; w.words = h->words;
        w.nb_class = nb_class; w.cls = c; w.classes = h->d_cls; w.rowflag = h->d_rowflag;
        if (int e = launch_sweep(h, w, batch, seg_stride, h->stream)) return e;
        h->launches += 3;
    }
    if (kept_idx || kept_counts) {
        AssembleArgs a;
        a.ibox = nullptr; a.objness = nullptr; a.classes = h->d_cls; a.cand = nullptr; a.counts = h->d_counts;
        a.seg_stride = seg_stride; a.nb_class = nb_class; a.max_out = 0; a.limit = 0;
        a.kept_idx = h->d_kept; a.kept_counts = h->d_kept_counts; a.dets = nullptr; a.det_counts = nullptr;
        assemble_yolo_kernel<<<batch, 1024, 0, h->stream>>>(a);
        CUDA_TRY(cudaGetLastError());
        ++h->launches;
    }
    if (int e = copy_out(h, h->d_cls, classes, n * nb_class * 4)) return e;
    if (int e = copy_out(h, h->d_kept, kept_idx, n * 4)) return e;
    if (int e = copy_out(h, h->d_kept_counts, kept_counts, (size_t)batch * 4)) return e;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FVY_OK;
}

int fvy_map_match(fvy_handle* h, const double* gt_box, const int32_t* gt_off, const double* det_box, const int32_t* det_off, int n_img,
                  double* det_iou, int32_t* img_any) {
    NO_CONV_HANDLE(h);
    if (!h || !gt_off || !det_off || !det_iou || !img_any) return fail(FVY_E_INVALID, "NULL argument");
    if (n_img < 0) return fail(FVY_E_INVALID, "n_img %d", n_img);
    if (n_img == 0) return FVY_OK;
    if (is_device_ptr(gt_off) || is_device_ptr(det_off) || is_device_ptr(det_iou)) return fail(FVY_E_INVALID, "fvy_map_match takes host pointers");
    const int n_gt = gt_off[n_img], n_det = det_off[n_img];
    if (gt_off[0] != 0 || det_off[0] != 0 || n_gt < 0 || n_det < 0 || (n_gt > 0 && !gt_box) || (n_det > 0 && !det_box)) return fail(FVY_E_INVALID, "bad offsets");
    std::vector<long long> pair_off(n_img + 1, 0);
    for (int i = 0; i < n_img; ++i) {
        const long long G = gt_off[i + 1] - gt_off[i], D = det_off[i + 1] - det_off[i];
        if (G < 0 || D < 0) return fail(FVY_E_INVALID, "offsets must be non-decreasing");
        pair_off[i + 1] = pair_off[i] + G * D;
    }
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    // transient device buffers: the scorer runs once per result file, off the hot path
    const size_t b_gt = (size_t)std::max(1, n_gt) * 32, b_det = (size_t)std::max(1, n_det) * 32, b_off = (size_t)(n_img + 1) * 4;
    const size_t b_po = (size_t)(n_img + 1) * 8, b_pairs = (size_t)std::max<long long>(1, pair_off[n_img]) * 8, b_iou = (size_t)std::max(1, n_det) * 8;
    const size_t b_any = (size_t)n_img * 4;
    char* d = nullptr;
    const size_t offs[8] = {0, b_gt, b_gt + b_det, b_gt + b_det + b_off, b_gt + b_det + 2 * b_off, 0, 0, 0};
    size_t total = b_gt + b_det + 2 * b_off;
    total = (total + 7) & ~size_t(7); const size_t o_po = total; total += b_po;
    const size_t o_pairs = total; total += b_pairs;
    const size_t o_iou = total; total += b_iou;
    const size_t o_any = total; total += b_any;
    CUDA_TRY(cudaMalloc((void**)&d, total));
    cudaError_t e = cudaSuccess;
    auto up = [&](size_t off, const void* src, size_t bytes) { if (e == cudaSuccess && bytes && src) e = cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, h->stream); };
    up(offs[0], gt_box, (size_t)n_gt * 32); up(offs[1], det_box, (size_t)n_det * 32); up(offs[2], gt_off, b_off); up(offs[3], det_off, b_off);
    up(o_po, pair_off.data(), b_po);
    MapMatchArgs a;
    a.gt = (const double*)(d + offs[0]); a.det = (const double*)(d + offs[1]); a.gt_off = (const int*)(d + offs[2]); a.det_off = (const int*)(d + offs[3]);
    a.pair_off = (const long long*)(d + o_po); a.pair_iou = (double*)(d + o_pairs); a.det_iou = (double*)(d + o_iou); a.img_any = (int*)(d + o_any);
    if (e == cudaSuccess) { map_match_kernel<<<n_img, 256, 0, h->stream>>>(a); e = cudaGetLastError(); ++h->launches; }
    if (e == cudaSuccess && n_det) e = cudaMemcpyAsync(det_iou, d + o_iou, (size_t)n_det * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(img_any, d + o_any, b_any, cudaMemcpyDeviceToHost, h->stream);
    const cudaError_t e2 = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess || e2 != cudaSuccess) return fail(FVY_E_CUDA, "fvy_map_match failed: %s", cudaGetErrorString(e != cudaSuccess ? e : e2));
    return FVY_OK;
}

int fvy_netout_sigmoid(fvy_handle* h, float* netout, long long n_boxes, int nb_class) {
    NO_CONV_HANDLE(h);
    if (!h || !netout) return fail(FVY_E_INVALID, "NULL argument");
    if (n_boxes < 0 || nb_class < 1) return fail(FVY_E_INVALID, "bad size");
    if (n_boxes == 0) return FVY_OK;
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const int ch = 5 + nb_class;
    const size_t bytes = (size_t)n_boxes * ch * 4;
    const bool dev = is_device_ptr(netout);
    float* d = netout;
    if (!dev) {            // a host array of the caller: staged through a transient device buffer (not on the hot path)
        CUDA_TRY(cudaMalloc((void**)&d, bytes));
        if (cudaMemcpyAsync(d, netout, bytes, cudaMemcpyHostToDevice, h->stream) != cudaSuccess) { cudaFree(d); return fail(FVY_E_CUDA, "netout upload failed"); }
    }
    const int blocks = (int)std::min<long long>((n_boxes * ch + 255) / 256, (long long)h->num_sms * 8);
    netout_sigmoid_kernel<<<blocks, 256, 0, h->stream>>>(d, n_boxes, ch);
    cudaError_t e1 = cudaGetLastError();
    ++h->launches;
    if (!dev && e1 == cudaSuccess) e1 = cudaMemcpyAsync(netout, d, bytes, cudaMemcpyDeviceToHost, h->stream);
    const cudaError_t e2 = cudaStreamSynchronize(h->stream);
    if (!dev) cudaFree(d);
    if (e1 != cudaSuccess || e2 != cudaSuccess) return fail(FVY_E_CUDA, "fvy_netout_sigmoid failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
    return FVY_OK;
}

static int postprocess_common(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch,
                              const fvy_post_params* pp, const int* image_hw, int max_out, fvy_det* dets, int32_t* det_counts,
                              bool sync) {
    if (int e = check_pp(h, pp)) return e;
    if (h->cfg.head == FVY_HEAD_NONE && (!out0 || !out1 || !out2)) return fail(FVY_E_INVALID, "post-processing-only handle needs all three logit tensors");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (max_out < 1 || max_out > h->dets_cap) return fail(FVY_E_INVALID, "max_out %d outside [1, %d]", max_out, h->dets_cap);
    if (!dets || !det_counts) return fail(FVY_E_INVALID, "NULL output");
    if (!image_hw && h->cfg.head != FVY_HEAD_FD6) return fail(FVY_E_INVALID, "image_hw is required (correct_yolo_boxes needs the image size)");
    const float* dev[3] = {nullptr, nullptr, nullptr};
    if (int e = resolve_logits(h, out0, out1, out2, batch, dev)) return e;
    const int* d_hw = nullptr;
    if (int e = upload_image_hw(h, image_hw, batch, &d_hw)) return e;
    CUDA_TRY(cudaStreamWaitEvent(h->ps, h->ev_d2h, 0));          // the previous call's detections have left the device buffers
    CUDA_TRY(cudaEventRecord(h->ev[2], h->ps));
    if (h->time_slot >= 0) CUDA_TRY(cudaEventRecord(h->ev_tp[h->time_slot][0], h->ps));
    if (int e = post_enqueue(h, dev, batch, pp, d_hw, max_out)) return e;
    if (h->time_slot >= 0) CUDA_TRY(cudaEventRecord(h->ev_tp[h->time_slot][1], h->ps));
    CUDA_TRY(cudaEventRecord(h->ev[3], h->ps));
    CUDA_TRY(cudaEventRecord(h->ev_post, h->ps));
    CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_post, 0));
    CUDA_TRY(cudaMemcpyAsync(dets, h->d_dets, (size_t)batch * max_out * sizeof(FvyDet), cudaMemcpyDefault, h->d2h_stream));
    CUDA_TRY(cudaMemcpyAsync(det_counts, h->d_det_counts, (size_t)batch * 4, cudaMemcpyDefault, h->d2h_stream));
    CUDA_TRY(cudaEventRecord(h->ev_d2h, h->d2h_stream));
    if (!sync) {       // the error report of this call: examined at the next synchronisation point (harvest_async)
        int* hs = h->h_async[h->logit_set];
        CUDA_TRY(cudaMemcpyAsync(hs, h->d_status, 4, cudaMemcpyDeviceToHost, h->ps));
        CUDA_TRY(cudaMemcpyAsync(hs + 1, h->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->ps));
        h->async_batch[h->logit_set] = batch;
        return FVY_OK;
    }
    std::vector<int> hc(batch);
    CUDA_TRY(cudaMemcpyAsync(hc.data(), h->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    CUDA_TRY(cudaEventElapsedTime(&h->last_post_ms, h->ev[2], h->ev[3]));
    return check_post_status(h, batch, hc.data());
}

int fvy_postprocess(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch, const fvy_post_params* pp,
                    const int* image_hw, int max_out, fvy_det* dets, int32_t* det_counts) {
    NO_CONV_HANDLE(h);
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int e = harvest_async_all(h)) return e;
    return postprocess_common(h, out0, out1, out2, batch, pp, image_hw, max_out, dets, det_counts, true);
}

// Adds the elapsed times of ring slot `slot` to the running sums (its events completed long ago, or the streams were just synchronised).
static void harvest_time_slot(fvy_handle* h, int slot) {
    if (!h->ring_used[slot]) return;
    h->ring_used[slot] = false;
    float a = 0.f, b = 0.f;
    if (cudaEventElapsedTime(&a, h->ev_tf[slot][0], h->ev_tf[slot][1]) != cudaSuccess ||
        cudaEventElapsedTime(&b, h->ev_tp[slot][0], h->ev_tp[slot][1]) != cudaSuccess) { cudaGetLastError(); return; }
    h->acc_fwd_ms += a; h->acc_post_ms += b; ++h->acc_calls;
}

static int detect_common(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp, const int* image_hw,
                         int max_out, fvy_det* dets, int32_t* det_counts, bool sync) {
    if (!h || !images) return fail(FVY_E_INVALID, "NULL argument");
    if (h->cfg.head == FVY_HEAD_NONE) return fail(FVY_E_STATE, "handle has no network");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const bool overlap = !sync && h->overlap_post;
    if (overlap) {
        h->logit_set ^= 1;
        if (int e = harvest_async(h, h->logit_set)) return e;                         // the deferred error report of that call
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[h->logit_set], 0));   // the post-processing that read this set (two calls ago) is done
    } else {
        if (int e = harvest_async_all(h)) return e;
        h->logit_set = 0;
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[0], 0));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[1], 0));              // nothing of an earlier asynchronous call is still in flight
    }
    const int slot = (int)(h->call_count++ % fvy_handle::kTimeRing);
    harvest_time_slot(h, slot);
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    CUDA_TRY(cudaEventRecord(h->ev_tf[slot][0], h->stream));
    if (int e = forward_enqueue(h, images, dtype, batch)) return e;
    CUDA_TRY(cudaEventRecord(h->ev_tf[slot][1], h->stream));
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    h->time_slot = slot;
    h->ring_used[slot] = true;
    int e = FVY_OK;
    if (overlap) {
        CUDA_TRY(cudaEventRecord(h->ev_fwd_done[h->logit_set], h->stream));
        CUDA_TRY(cudaStreamWaitEvent(h->post_stream, h->ev_fwd_done[h->logit_set], 0));
        h->ps = h->post_stream;
        e = postprocess_common(h, nullptr, nullptr, nullptr, batch, pp, image_hw, max_out, dets, det_counts, false);
        h->ps = h->stream;
        if (e == FVY_OK) CUDA_TRY(cudaEventRecord(h->ev_post_done[h->logit_set], h->post_stream));
    } else {
        e = postprocess_common(h, nullptr, nullptr, nullptr, batch, pp, image_hw, max_out, dets, det_counts, sync);
        if (e == FVY_OK && !sync) CUDA_TRY(cudaEventRecord(h->ev_post_done[0], h->stream));
    }
    h->time_slot = -1;
    if (e) return e;
    if (sync) CUDA_TRY(cudaEventElapsedTime(&h->last_fwd_ms, h->ev[0], h->ev[1]));
    return FVY_OK;
}

int fvy_detect(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp, const int* image_hw, int max_out,
               fvy_det* dets, int32_t* det_counts) {
    return detect_common(h, images, dtype, batch, pp, image_hw, max_out, dets, det_counts, true);
}

int fvy_detect_async(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp, const int* image_hw,
                     int max_out, fvy_det* dets, int32_t* det_counts) {
    return detect_common(h, images, dtype, batch, pp, image_hw, max_out, dets, det_counts, false);
}

int fvy_sync(fvy_handle* h) {
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->h2d_stream));
    CUDA_TRY(cudaStreamSynchronize(h->post_stream));
    CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    return harvest_async_all(h);       // FVY_E_RANGE / FVY_E_CAPACITY of an asynchronous call surface here
}

int fvy_num_layers(const fvy_handle* h) { return h ? (int)h->layers.size() : 0; }

int fvy_layer_info(const fvy_handle* h, int layer, int* info) {
    if (!h || !info || layer < 0 || layer >= (int)h->layers.size()) return fail(FVY_E_INVALID, "bad layer %d", layer);
    const Layer& L = h->layers[layer];
    const int m_total = h->cfg.max_batch * L.p.dom_plane;
    const int tiles = ((m_total + kBlockM - 1) / kBlockM) * L.num_n_tiles;
    const int v[12] = {L.s.idx, L.s.cin, L.s.cout, L.s.k, L.s.stride, L.Hout, L.Wout, L.BN, L.BK,
                       L.stages + 100 * L.b_stages + 10000 * L.b_resident + 100000 * (L.cta2 ? 1 : 0) + 1000000 * L.p.a_slab,
                       std::min(tiles, h->num_sms * L.occ), tiles};
    memcpy(info, v, sizeof(v));
    return FVY_OK;
}

int fvy_layer_output(fvy_handle* h, int layer, int batch, float* dst_host) {
    NO_CONV_HANDLE(h);
    if (!h || !dst_host || layer < 0 || layer >= (int)h->layers.size()) return fail(FVY_E_INVALID, "bad argument");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d", batch);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const Layer& L = h->layers[layer];
    if (layer == 0 && h->fuse_stem && !h->stem_phase_valid)      // the fused forward keeps conv_0's activation on chip: make it now
        if (int e = run_layers(h, batch, 0, 1)) return e;
    const size_t n = (size_t)batch * L.Hout * L.Wout * L.s.cout;
    float* tmp = nullptr;
    CUDA_TRY(cudaMalloc(&tmp, n * 4));
    unpack_kernel<<<h->num_sms * 8, 256, 0, h->stream>>>(L.primary, batch, L.Hout, L.Wout, L.s.cout, tmp);
    cudaError_t e1 = cudaGetLastError();
    cudaError_t e2 = cudaMemcpyAsync(dst_host, tmp, n * 4, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e3 = cudaStreamSynchronize(h->stream);
    cudaFree(tmp);
    if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess)
        return fail(FVY_E_CUDA, "layer_output failed: %s", cudaGetErrorString(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3)));
    return FVY_OK;
}

float* fvy_staged_images(fvy_handle* h) {
    if (h && h->conv_mode) { fail(FVY_E_STATE, "single-convolution handle"); return nullptr; }
    if (!h) return nullptr;
    if (!h->d_staged) {
        cudaSetDevice(h->cfg.device);
        const size_t bytes = (size_t)h->cfg.max_batch * h->cfg.net_h * h->cfg.net_w * 3 * sizeof(float);
        void* p = nullptr;
        if (dev_alloc(h, &p, bytes, true) != FVY_OK) return nullptr;
        h->d_staged = (float*)p;
    }
    return h->d_staged;
}

int fvy_read_staged(fvy_handle* h, int batch, float* dst) {
    NO_CONV_HANDLE(h);
    if (!h || !dst) return fail(FVY_E_INVALID, "fvy_read_staged: NULL argument");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "fvy_read_staged: batch %d outside [1, %d]", batch, h->cfg.max_batch);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const float* staged = fvy_staged_images(h);
    if (!staged) return fail(FVY_E_CUDA, "fvy_read_staged: staged batch allocation failed");
    if (int e = copy_out(h, staged, dst, (size_t)batch * h->cfg.net_h * h->cfg.net_w * 3 * sizeof(float))) return e;
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    return FVY_OK;
}

int fvy_letterbox_u8(fvy_handle* h, const unsigned char* src, int src_h, int src_w, int w_p, int h_p, int pad_t, int pad_l, int index) {
    NO_CONV_HANDLE(h);
    if (!h || !src) return fail(FVY_E_INVALID, "fvy_letterbox_u8: NULL argument");
    if (src_h <= 0 || src_w <= 0 || w_p <= 0 || h_p <= 0 || pad_t < 0 || pad_l < 0 || pad_t + h_p > h->cfg.net_h || pad_l + w_p > h->cfg.net_w)
        return fail(FVY_E_INVALID, "fvy_letterbox_u8: %dx%d -> %dx%d at (%d, %d) does not fit the %dx%d network input", src_w, src_h, w_p, h_p, pad_l, pad_t,
                    h->cfg.net_w, h->cfg.net_h);
    if (index < 0 || index >= h->cfg.max_batch) return fail(FVY_E_INVALID, "fvy_letterbox_u8: image slot %d outside [0, %d)", index, h->cfg.max_batch);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    float* staged = fvy_staged_images(h);
    if (!staged) return fail(FVY_E_CUDA, "fvy_letterbox_u8: staged batch allocation failed");
    const unsigned char* dsrc = src;
    if (!is_device_ptr(src)) {
        const size_t bytes = (size_t)src_h * src_w * 3;
        if (bytes > h->lb_src_bytes) {                    // grows to the largest image seen (stream-ordered: earlier kernels have been enqueued)
            CUDA_TRY(cudaStreamSynchronize(h->stream));
            if (h->d_lb_src) cudaFree(h->d_lb_src);
            h->d_lb_src = nullptr; h->lb_src_bytes = 0;
            CUDA_TRY(cudaMalloc((void**)&h->d_lb_src, bytes));
            h->lb_src_bytes = bytes;
        }
        CUDA_TRY(cudaMemcpyAsync(h->d_lb_src, src, bytes, cudaMemcpyHostToDevice, h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));       // pageable source: the caller may reuse it, and the scratch is reused per image
        dsrc = h->d_lb_src;
    }
    const dim3 grid((h->cfg.net_w + 127) / 128, h->cfg.net_h);
    letterbox_u8_kernel<<<grid, 128, 0, h->stream>>>(dsrc, src_h, src_w, w_p, h_p, pad_t, pad_l, h->cfg.net_h, h->cfg.net_w,
                                                     staged + (size_t)index * h->cfg.net_h * h->cfg.net_w * 3);
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    if (!is_device_ptr(src)) CUDA_TRY(cudaStreamSynchronize(h->stream));   // the scratch is free for the next image
    return FVY_OK;
}

long long fvy_launch_count(const fvy_handle* h) { return h ? h->launches : 0; }

int fvy_last_timing(const fvy_handle* h, float* forward_ms, float* post_ms) {
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    if (forward_ms) *forward_ms = h->last_fwd_ms;
    if (post_ms) *post_ms = h->last_post_ms;
    return FVY_OK;
}

int fvy_profile_layers(fvy_handle* h, int batch, int iters, float* ms) {
    NO_CONV_HANDLE(h);
    if (!h || !ms) return fail(FVY_E_INVALID, "NULL argument");
    if (!h->weights_loaded) return fail(FVY_E_STATE, "weights not loaded");
    if (batch < 1 || batch > h->cfg.max_batch || iters < 1) return fail(FVY_E_INVALID, "bad batch/iters");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (size_t i = 0; i < h->layers.size(); ++i) {
        if (int e = run_layers(h, batch, (int)i, (int)i + 1)) return e;   // warm
        CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
        for (int it = 0; it < iters; ++it)
            if (int e = run_layers(h, batch, (int)i, (int)i + 1)) return e;
        CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
        CUDA_TRY(cudaStreamSynchronize(h->stream));
        float t = 0.f;
        CUDA_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
        ms[i] = t / iters;
    }
    return FVY_OK;
}

int fvy_run_layer(fvy_handle* h, int layer, int batch, int iters, float* ms) {
    if (!h || !ms) return fail(FVY_E_INVALID, "NULL argument");
    if (!h->weights_loaded) return fail(FVY_E_STATE, "weights not loaded");
    if (layer < 0 || layer >= (int)h->layers.size() || batch < 1 || batch > h->cfg.max_batch || iters < 1) return fail(FVY_E_INVALID, "bad layer/batch/iters");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    if (int e = run_layers(h, batch, layer, layer + 1)) return e;
    if (getenv("FVY_DBG") && h->layers[layer].s.src != -1) {
        // cycle counters of the single-thread roles and wall-clock (globaltimer) milestones, from two back-to-back launches
        Layer& L = h->layers[layer];
        unsigned long long* d = nullptr;
        const int n = h->num_sms * 32;
        CUDA_TRY(cudaMalloc(&d, 2 * n * 8));
        CUDA_TRY(cudaMemsetAsync(d, 0, 2 * n * 8, h->stream));
        L.p.dbg = d;
        int e = run_layers(h, batch, layer, layer + 1);
        L.p.dbg = d + n;
        if (!e) e = run_layers(h, batch, layer, layer + 1);
        L.p.dbg = nullptr;
        std::vector<unsigned long long> v2(2 * n);
        cudaMemcpyAsync(v2.data(), d, 2 * n * 8, cudaMemcpyDeviceToHost, h->stream);
        cudaStreamSynchronize(h->stream);
        cudaFree(d);
        if (e) return e;
        const unsigned long long* v = v2.data() + n;           // second launch
        double s[16] = {0}; int cnt = 0, ecnt = 0;
        for (int c = 0; c < h->num_sms; ++c)
            if (v[c * 32] || v[c * 32 + 4] || v[c * 32 + 8]) {
                for (int k = 0; k < 16; ++k) s[k] += (double)v[c * 32 + k];
                if (v[c * 32]) ++cnt;
                if (v[c * 32 + 8]) ++ecnt;
            }
        if (ecnt)
            fprintf(stderr, "fvy dbg conv_%d epilogue group 0: total %.0f clk  chunks %.0f  => %.0f clk/chunk: wait_tmem_full %.0f  wait_res %.0f  "
                            "named_barrier %.0f  direct_stores %.0f  tile_setup(per chunk) %.0f  body(ld..fence, incl. wait_res) %.0f\n",
                    L.s.idx, s[8] / ecnt, s[13] / ecnt, s[8] / std::max(1.0, s[13]), s[9] / std::max(1.0, s[13]), s[10] / std::max(1.0, s[13]),
                    s[11] / std::max(1.0, s[13]), s[12] / std::max(1.0, s[13]), s[14] / std::max(1.0, s[13]), s[15] / std::max(1.0, s[13]));
        if (cnt)
            fprintf(stderr, "fvy dbg conv_%d: issuing CTAs %d  total %.0f clk  wait_full %.0f  wait_tmem_empty %.0f  taps %.0f  => %.0f clk/tap "
                            "(%.0f outside waits); producers wait_empty A %.0f B %.0f\n",
                    L.s.idx, cnt, s[0] / cnt, s[1] / cnt, s[2] / cnt, s[3] / cnt, s[0] / std::max(1.0, s[3]),
                    (s[0] - s[1] - s[2]) / std::max(1.0, s[3]), s[4] / cnt, s[5] / cnt);
        // timeline in microseconds relative to the first CTA start of the second launch: [min / max over CTAs]
        auto mm = [&](const unsigned long long* base, int slot, unsigned long long* lo, unsigned long long* hi) {
            *lo = ~0ull; *hi = 0;
            for (int c = 0; c < h->num_sms; ++c) { const unsigned long long t = base[c * 32 + slot]; if (t) { *lo = std::min(*lo, t); *hi = std::max(*hi, t); } }
        };
        unsigned long long lo[7], hi[7], plo, phi;
        for (int k = 0; k < 7; ++k) mm(v, 16 + k, &lo[k], &hi[k]);
        mm(v2.data(), 22, &plo, &phi);
        const double t0 = (double)lo[0];
        auto us = [&](unsigned long long t) { return t == ~0ull || t == 0 ? -1.0 : ((double)t - t0) / 1e3; };
        fprintf(stderr, "fvy dbg conv_%d timeline [us, min..max over CTAs; 0 = first CTA start]: previous launch's last CTA end %.2f | start %.2f..%.2f | "
                        "prologue done %.2f..%.2f | after griddepcontrol.wait %.2f..%.2f | first operands landed %.2f..%.2f | MMA loop done %.2f..%.2f | "
                        "epilogue done %.2f..%.2f | CTA end %.2f..%.2f\n",
                L.s.idx, ((double)phi - t0) / 1e3, us(lo[0]), us(hi[0]), us(lo[1]), us(hi[1]), us(lo[2]), us(hi[2]), us(lo[3]), us(hi[3]),
                us(lo[4]), us(hi[4]), us(lo[5]), us(hi[5]), us(lo[6]), us(hi[6]));
    }
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    for (int it = 0; it < iters; ++it)
        if (int e = run_layers(h, batch, layer, layer + 1)) return e;
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    float t = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&t, h->ev[0], h->ev[1]));
    *ms = t / iters;
    return FVY_OK;
}

int fvy_timer_start(fvy_handle* h) {
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    for (bool& u : h->ring_used) u = false;
    h->acc_fwd_ms = h->acc_post_ms = 0.0; h->acc_calls = 0;
    CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
    return FVY_OK;
}
int fvy_timer_breakdown(fvy_handle* h, float* forward_ms_mean, float* post_ms_mean, int* calls) {
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->post_stream));
    for (int s = 0; s < fvy_handle::kTimeRing; ++s) harvest_time_slot(h, s);
    const double n = (double)std::max<long long>(1, h->acc_calls);
    if (forward_ms_mean) *forward_ms_mean = (float)(h->acc_fwd_ms / n);
    if (post_ms_mean) *post_ms_mean = (float)(h->acc_post_ms / n);
    if (calls) *calls = (int)h->acc_calls;
    return FVY_OK;
}
int fvy_timer_stop(fvy_handle* h, float* ms) {
    if (!h || !ms) return fail(FVY_E_INVALID, "NULL argument");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_d2h, 0));      // include the last result copy in the measured interval
    CUDA_TRY(cudaEventRecord(h->ev[5], h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    CUDA_TRY(cudaEventElapsedTime(ms, h->ev[4], h->ev[5]));
    return FVY_OK;
}

int fvy_adam_step(float* param, const float* grad, float* m, float* v, long long n, float lr_t, float beta_1, float beta_2,
                  float epsilon, float grad_scale, void* cuda_stream) {
    if (!param || !grad || !m || !v || n < 0) return fail(FVY_E_INVALID, "fvy_adam_step: bad argument");
    if (n == 0) return FVY_OK;
    if (!is_device_ptr(param) || !is_device_ptr(grad) || !is_device_ptr(m) || !is_device_ptr(v))
        return fail(FVY_E_INVALID, "fvy_adam_step takes device pointers (there is no CPU path)");
    if ((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15)
        return fail(FVY_E_INVALID, "fvy_adam_step: buffers must be 16-byte aligned");
    int dev = 0, sms = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    CUDA_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long want = ((n >> 2) + 255) / 256;
    const int blocks = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * 8));
    adam_step_kernel<<<blocks, 256, 0, (cudaStream_t)cuda_stream>>>(param, grad, m, v, n, lr_t, beta_1, beta_2, epsilon, grad_scale);
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

static int bn_check(const void* x, long long rows, int C, const void* ws) {
    if (!x || !ws || rows < 1 || C < 4 || (C & 3)) return fail(FVY_E_INVALID, "fvy_bn_leaky_*: bad argument (rows %lld, C %d: C must be a positive multiple of 4)", rows, C);
    if (!is_device_ptr(x) || !is_device_ptr(ws)) return fail(FVY_E_INVALID, "fvy_bn_leaky_* take device pointers (there is no CPU path)");
    if (reinterpret_cast<uintptr_t>(x) & 15) return fail(FVY_E_INVALID, "fvy_bn_leaky_*: activations must be 16-byte aligned");
    return FVY_OK;
}
static void bn_grids(long long rows, int C, dim3* sums, int* apply) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int cy = (C + 127) / 128;
    const long long want = (rows + 7) / 8;
    *sums = dim3((unsigned)std::max<long long>(1, std::min<long long>(want, (long long)sms * 8 / cy + 1)), cy);
    *apply = (int)std::max<long long>(1, std::min<long long>((rows * C / 4 + 255) / 256, (long long)sms * 16));
}

int fvy_bn_leaky_train_forward(const float* x, long long rows, int C, const float* gamma, const float* beta, float eps, float momentum, float slope,
                               float* running_mean, float* running_var, float* y, float* save_mean, float* save_invstd, double* workspace,
                               void* cuda_stream) {
    if (int e = bn_check(x, rows, C, workspace)) return e;
    if (!gamma || !beta || !y || !save_mean || !save_invstd) return fail(FVY_E_INVALID, "fvy_bn_leaky_train_forward: NULL argument");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    dim3 gs; int ga;
    bn_grids(rows, C, &gs, &ga);
    CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)2 * C * sizeof(double), st));
    bn_sums_kernel<0><<<gs, kBnThreads, 0, st>>>(x, nullptr, rows, C, nullptr, nullptr, nullptr, nullptr, slope, workspace);
    bn_fwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(workspace, rows, C, eps, momentum, running_mean, running_var, save_mean, save_invstd);
    bn_fwd_apply_kernel<<<ga, 256, 0, st>>>(x, rows * C / 4, C, gamma, beta, save_mean, save_invstd, slope, y);
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

int fvy_bn_leaky_train_backward(const float* x, const float* dy, long long rows, int C, const float* gamma, const float* beta, const float* save_mean,
                                const float* save_invstd, float slope, float* dx, float* dgamma, float* dbeta, double* workspace, void* cuda_stream) {
    if (int e = bn_check(x, rows, C, workspace)) return e;
    if (!dy || !gamma || !beta || !save_mean || !save_invstd || !dx || !dgamma || !dbeta) return fail(FVY_E_INVALID, "fvy_bn_leaky_train_backward: NULL argument");
    if (reinterpret_cast<uintptr_t>(dy) & 15) return fail(FVY_E_INVALID, "fvy_bn_leaky_train_backward: gradients must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    dim3 gs; int ga;
    bn_grids(rows, C, &gs, &ga);
    CUDA_TRY(cudaMemsetAsync(workspace, 0, (size_t)2 * C * sizeof(double), st));
    bn_sums_kernel<1><<<gs, kBnThreads, 0, st>>>(x, dy, rows, C, gamma, beta, save_mean, save_invstd, slope, workspace);
    bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, st>>>(workspace, C, dgamma, dbeta);
    bn_bwd_apply_kernel<<<ga, 256, 0, st>>>(x, dy, rows * C / 4, rows, C, gamma, beta, save_mean, save_invstd, slope, workspace, dx);
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

// ------------------------------------------------------------------------------------------ single convolution (row f-1: dgrad)
int fvy_conv_create(int device, int height, int width, int cin, int cout, int ksize, int stride, int max_batch, fvy_handle** out) {
    if (!out) return fail(FVY_E_INVALID, "NULL argument");
    *out = nullptr;
    if (height < 1 || width < 1 || height > 4096 || width > 4096) return fail(FVY_E_INVALID, "feature map %dx%d", height, width);
    if (ksize != 1 && ksize != 3) return fail(FVY_E_INVALID, "kernel size %d (1 or 3)", ksize);
    if (stride != 1 && !(stride == 2 && ksize == 3 && height % 2 == 0 && width % 2 == 0 && cin % 64 == 0))
        return fail(FVY_E_INVALID, "stride %d (1, or 2 for a 3 x 3 filter over an even-sized map with Cin a multiple of 64)", stride);
    if (cin < 32 || cin % 32 || cin > 2048) return fail(FVY_E_INVALID, "Cin %d must be a multiple of 32 in [32, 2048]", cin);
    if (cout < 1 || cout > kMaxCout) return fail(FVY_E_INVALID, "Cout %d outside [1, %d]", cout, kMaxCout);
    if (max_batch < 1 || max_batch > 1024) return fail(FVY_E_INVALID, "max_batch %d outside [1, 1024]", max_batch);
    fvy_config cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.device = device; cfg.net_h = height; cfg.net_w = width; cfg.head = FVY_HEAD_YOLO3; cfg.nb_class = 1; cfg.bb_info_c_size = 6;
    cfg.max_batch = max_batch; cfg.flags = FVY_CFG_NO_GRAPH | FVY_CFG_NO_CHAIN | FVY_CFG_NO_TILE_FLAGS;
    return create_impl(&cfg, cin, cout, ksize, out, stride);
}

int fvy_conv_set_weights(fvy_handle* h, const float* w_dev, int dgrad, void* stream) {
    if (!h || !w_dev) return fail(FVY_E_INVALID, "NULL argument");
    if (!h->conv_mode) return fail(FVY_E_STATE, "not a single-convolution handle");
    if (dgrad && h->conv_stride != 1) return fail(FVY_E_INVALID, "dgrad weights on a stride-2 handle");
    if (!is_device_ptr(w_dev)) return fail(FVY_E_INVALID, "fvy_conv_set_weights takes a device pointer");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    Layer& L = h->layers[0];
    const long long total = (long long)L.cout_pad * L.taps * L.cin_pad;
    conv_weight_kernel<<<(unsigned)std::min<long long>((total + 255) / 256, 4096), 256, 0, (cudaStream_t)stream>>>(
        w_dev, h->conv_cin, h->conv_cout, L.cout_pad, L.taps, dgrad != 0, L.tap_perm, L.w);
    CUDA_TRY(cudaGetLastError());
    h->weights_loaded = true;
    return FVY_OK;
}

int fvy_conv_run(fvy_handle* h, const float* x_dev, int batch, float* y_dev, void* stream) {
    if (!h || !x_dev || !y_dev) return fail(FVY_E_INVALID, "NULL argument");
    if (!h->conv_mode) return fail(FVY_E_STATE, "not a single-convolution handle");
    if (!h->weights_loaded) return fail(FVY_E_STATE, "fvy_conv_run before fvy_conv_set_weights");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (!is_device_ptr(x_dev) || !is_device_ptr(y_dev)) return fail(FVY_E_INVALID, "fvy_conv_run takes device pointers");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    Layer& L = h->layers[0];
    const cudaStream_t st = (cudaStream_t)stream, saved = h->stream;
    const long long groups = (long long)batch * L.Hin * L.Win * (h->conv_cin / 8);
    if (h->conv_stride == 2)
        pack_phase_kernel<<<(unsigned)std::min<long long>((groups + 255) / 256, (long long)h->num_sms * 32), 256, 0, st>>>(
            x_dev, batch, L.Hin, L.Win, h->conv_cin, L.p.dom_w, L.p.dom_plane, (long long)h->cfg.max_batch * L.p.dom_plane, h->d_conv_in);
    else
        pack_padded_kernel<<<(unsigned)std::min<long long>((groups + 255) / 256, (long long)h->num_sms * 32), 256, 0, st>>>(
            x_dev, batch, L.Hin, L.Win, h->conv_cin, L.p.dom_w, L.p.dom_plane, h->d_conv_in);
    CUDA_TRY(cudaGetLastError());
    h->conv_out = y_dev;
    h->stream = st;                       // the layer is enqueued on the caller's stream, behind the pack kernel
    const int e = run_layers(h, batch, 0, 1);
    h->stream = saved;
    return e;
}

// ------------------------------------------------------------------------------------------ weight gradient (row f-1: wgrad)
static void wgrad_geometry(int batch, int H, int W, int* pitch, long long* plane, int* lead, long long* rows_k, long long* rows_total) {
    *pitch = W + 1; *plane = (long long)(H + 1) * (W + 1);
    *lead = (*pitch + 1 + 7) & ~7;
    *rows_k = ((long long)batch * *plane + kWtKC - 1) / kWtKC * kWtKC;        // a multiple of both kernels' K chunk (64 and 32 pixels)
    *rows_total = *lead + *rows_k + *pitch + 2 + 8;
}
long long fvy_conv_wgrad_scratch_rows(int batch, int height, int width) {
    int pitch, lead; long long plane, rows_k, total;
    wgrad_geometry(batch, height, width, &pitch, &plane, &lead, &rows_k, &total);
    return total;
}
int fvy_conv_wgrad(const float* x_dev, const float* dy_dev, int batch, int height, int width, int cin, int cout, int ksize, int stride,
                   void* x_scratch, void* dy_scratch, float* dw_dev, float* dw_work, void* cuda_stream) {
    if (!x_dev || !dy_dev || !x_scratch || !dy_scratch || !dw_dev) return fail(FVY_E_INVALID, "fvy_conv_wgrad: NULL argument");
    if (ksize != 1 && ksize != 3) return fail(FVY_E_INVALID, "fvy_conv_wgrad: kernel size %d (1 or 3)", ksize);
    if (cin < 64 || cin % 64 || cout < 64 || cout % 64) return fail(FVY_E_INVALID, "fvy_conv_wgrad: Cin %d / Cout %d must be multiples of 64", cin, cout);
    if (batch < 1 || height < 1 || width < 1) return fail(FVY_E_INVALID, "fvy_conv_wgrad: batch %d, map %dx%d", batch, height, width);
    if (stride != 1 && !(stride == 2 && ksize == 3 && cout % 128 == 0)) return fail(FVY_E_INVALID, "fvy_conv_wgrad: stride %d (2 needs a 3 x 3 filter and Cout a multiple of 128)", stride);
    int pitch, lead; long long plane, rows_k, total;
    wgrad_geometry(batch, height, width, &pitch, &plane, &lead, &rows_k, &total);
    if (rows_k + pitch + 2 >= (1ll << 31)) return fail(FVY_E_INVALID, "fvy_conv_wgrad: %lld rows overflow int32", rows_k);
    const cudaStream_t st = (cudaStream_t)cuda_stream;
    static int num_sms = 0;
    if (!num_sms) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev); }
    __nv_bfloat16* xs = (__nv_bfloat16*)x_scratch + (size_t)lead * cin;
    __nv_bfloat16* ys = (__nv_bfloat16*)dy_scratch + (size_t)lead * cout;
    // rows behind the packed images (the round-up of the K range and the reach of the taps) may hold pixels of an earlier, larger batch
    const long long tail0 = (long long)batch * plane, tail_rows = rows_k - tail0 + pitch + 2;
    CUDA_TRY(cudaMemsetAsync(xs + tail0 * cin, 0, (size_t)tail_rows * cin * 2, st));
    CUDA_TRY(cudaMemsetAsync(ys + tail0 * cout, 0, (size_t)tail_rows * cout * 2, st));
    CUDA_TRY(cudaMemsetAsync(dw_dev, 0, (size_t)cout * cin * ksize * ksize * 4, st));
    // height / width are the OUTPUT's (= dY's) map; a stride-2 layer's input is twice that and is packed as four phase planes, `total` rows apart
    const int s2 = stride == 2 ? 2 : 1;
    const long long gx = (long long)batch * height * s2 * width * s2 * (cin / 8), gy = (long long)batch * height * width * (cout / 8);
    if (stride == 2)
        pack_phase_kernel<<<(unsigned)std::min<long long>((gx + 255) / 256, (long long)num_sms * 32), 256, 0, st>>>(x_dev, batch, 2 * height, 2 * width, cin, pitch, (int)plane, total, xs);
    else
        pack_padded_kernel<<<(unsigned)std::min<long long>((gx + 255) / 256, (long long)num_sms * 32), 256, 0, st>>>(x_dev, batch, height, width, cin, pitch, (int)plane, xs);
    pack_padded_kernel<<<(unsigned)std::min<long long>((gy + 255) / 256, (long long)num_sms * 32), 256, 0, st>>>(dy_dev, batch, height, width, cout, pitch, (int)plane, ys);
    static const int tc_env = [] { const char* v = getenv("FVY_WGRAD_TC"); return v && *v ? atoi(v) : 1; }();
    if (tc_env && cout % 128 == 0) {
        // tcgen05 path: both operands as they are (MN-major descriptors), 128 output channels x up to 256 input channels per item
        CUtensorMap map_dy, map_x;
        if (int e = make_tmap_2d(&map_dy, (const __nv_bfloat16*)dy_scratch, (uint64_t)cout, (uint64_t)total, (uint64_t)cout, 64, 64)) return e;
        if (int e = make_tmap_2d(&map_x, (const __nv_bfloat16*)x_scratch, (uint64_t)cin, (uint64_t)total * (stride == 2 ? 4 : 1), (uint64_t)cin, 64, 64)) return e;
        WgradTcParams p;
        p.cin = cin; p.cout = cout; p.taps = ksize * ksize; p.lead = lead; p.chunks = (int)(rows_k / kWtKC);
        for (int t = 0; t < 9; ++t) {
            const int r = t / 3, q = t % 3;
            if (ksize == 1) p.tap_row[t] = 0;
            else if (stride == 1) p.tap_row[t] = (r - 1) * pitch + (q - 1);
            else {      // phase (r & 1, q & 1), position shifted by (r >> 1, q >> 1) - 1
                const long long v = (long long)(((r & 1) << 1) | (q & 1)) * total + ((r >> 1) - 1) * pitch + ((q >> 1) - 1);
                if (v >= (1ll << 31)) return fail(FVY_E_INVALID, "fvy_conv_wgrad: phase offset overflows int32");
                p.tap_row[t] = (int)v;
            }
        }
        p.n_tile = cin % 256 == 0 ? 256 : (cin % 128 == 0 ? 128 : 64);
        const int base_items = (cout / 128) * (cin / p.n_tile) * p.taps;
        // one round of items: an item's epilogue (its reds) only overlaps the MMAs of a following item, and every extra pixel range is
        // one more red per output element
        p.ksplit = std::max(1, std::min(std::max(1, p.chunks / 4), num_sms / base_items));
        const bool permute = p.taps > 1;
        if (permute && !dw_work) return fail(FVY_E_INVALID, "fvy_conv_wgrad: dw_work is required for a 3 x 3 filter");
        p.dw = permute ? dw_work : dw_dev;
        if (permute) CUDA_TRY(cudaMemsetAsync(dw_work, 0, (size_t)cout * cin * p.taps * 4, st));
        const int items = base_items * p.ksplit;
        const int smem = wt_smem_bytes(p.n_tile);
        static bool attr_tc = false;
        if (!attr_tc) { CUDA_TRY(cudaFuncSetAttribute(wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, wt_smem_bytes(256))); attr_tc = true; }
        wgrad_tc_kernel<<<std::min(items, num_sms), kWtThreads, smem, st>>>(map_dy, map_x, p);
        if (permute) wgrad_permute_kernel<<<num_sms * 4, 256, 0, st>>>(dw_work, cout, cin, p.taps, dw_dev);
        CUDA_TRY(cudaGetLastError());
        return FVY_OK;
    }
    if (stride == 2) return fail(FVY_E_INVALID, "fvy_conv_wgrad: stride 2 runs on the tcgen05 kernel only (FVY_WGRAD_TC=0 is set)");
    const int chunks = (int)(rows_k / kWgKC);
    dim3 grid(cout / 64, cin / 64, 1);
    grid.z = (unsigned)std::max(1, std::min(chunks, (2 * num_sms + (int)(grid.x * grid.y) - 1) / (int)(grid.x * grid.y)));
    if (ksize == 3) {
        static bool attr = false;
        if (!attr) { CUDA_TRY(cudaFuncSetAttribute(conv_wgrad_kernel<9>, cudaFuncAttributeMaxDynamicSharedMemorySize, WgSmem<9>::kBytes)); attr = true; }
        conv_wgrad_kernel<9><<<grid, kWgThreads, WgSmem<9>::kBytes, st>>>(xs, ys, (int)rows_k, pitch, cin, cout, dw_dev);
    } else {
        conv_wgrad_kernel<1><<<grid, kWgThreads, WgSmem<1>::kBytes, st>>>(xs, ys, (int)rows_k, pitch, cin, cout, dw_dev);
    }
    CUDA_TRY(cudaGetLastError());
    return FVY_OK;
}

void* fvy_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void fvy_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
