// Host-side state of one fvy_handle: error channel, TMA tensor-map encoders, the per-layer plan record and the handle itself.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/fvy.h"
#include "conv_igemm_sm100.cuh"
#include "conv_chain_sm100.cuh"
#include "fvy_plan.h"
#include "postproc_kernels.cuh"

#include "stem_kernel.cuh"
#include "stem_conv1_fused.cuh"
#include "prepost_kernels.cuh"
#include "train_kernels.cuh"
#include "wgrad_kernel.cuh"
#include "wgrad_tc_sm100.cuh"

namespace fvy {

// ------------------------------------------------------------------------------------------ errors
static thread_local char g_err[512] = "";
static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CUDA_TRY(expr)                                                                                   \
    do {                                                                                                 \
        cudaError_t _e = (expr);                                                                         \
        if (_e != cudaSuccess) return fail(FVY_E_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

// ------------------------------------------------------------------------------------------ TMA encode
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D bf16 tensor [rows][cols] (cols contiguous, row pitch `pitch_elems`), box = box_cols x box_rows,
// swizzle = box_cols * 2 bytes (64 or 128).  Out-of-bounds elements read as zero.
static int make_tmap_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint64_t pitch_elems, uint32_t box_cols,
                        uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {pitch_elems * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (cols=%llu rows=%llu pitch=%llu box=%ux%u)",
                                       (int)r, (unsigned long long)cols, (unsigned long long)rows, (unsigned long long)pitch_elems, box_cols, box_rows);
    return FVY_OK;
}

// Weights [Cout][taps][Cin] bf16 seen as a 3-D tensor (Cin, Cout, tap): one box = the [box_rows, box_cols] tiles of `box_taps`
// consecutive taps, laid out in shared memory tap after tap - i.e. the B tiles of a whole filter row with ONE TMA instruction
// (a thread issues a TMA instruction every ~200 cycles whatever its size, tools/tma_bench.cu).
static int make_tmap_b3(CUtensorMap* m, const void* base, uint64_t cin, uint64_t cout, uint64_t taps, uint32_t box_cols, uint32_t box_rows,
                        uint32_t box_taps) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[3] = {cin, cout, taps};
    cuuint64_t strides[2] = {taps * cin * 2, cin * 2};
    cuuint32_t box[3] = {box_cols, box_rows, box_taps};
    cuuint32_t estr[3] = {1, 1, 1};
    CUtensorMapSwizzle sw = box_cols * 2 == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled (3-D weights) failed with CUresult %d (cin=%llu cout=%llu taps=%llu box=%ux%ux%u)",
                                       (int)r, (unsigned long long)cin, (unsigned long long)cout, (unsigned long long)taps, box_cols, box_rows, box_taps);
    return FVY_OK;
}

// 4-phase buffer [rp][cp][n][H/2+2][W/2+2][C] (what a stride-2 consumer reads) seen from the PRODUCER's padded compute domain:
// pixel (img, hp, wp) lives at phase (hp&1, wp&1), position (hp>>1, wp>>1).  A run of consecutive pixels of one image row is
// the box (32 channels, cp = 0..1, 64 values of wp>>1) of the 5-D tensor (C, cp, wp>>1, hp>>1, rp*2*nmax + img): its layout in
// shared memory - C fastest, then cp, then wp>>1 - is exactly 2 x box_pairs consecutive domain rows of a staged 32-channel
// chunk.  TMA clips the part of a box that runs past (W+2)/2; a negative start raises "illegal instruction" and a run that
// ends at the tile's end (not the row's) has nothing to clip it, so the store warp covers such a run with two (overlapping)
// boxes of the largest power of two that fits - hence one map per box size 1, 2, 4 ... 64 pairs, kept in global memory
// (tools/phase_tma_test.cu pins these properties).  Replaces 128 threads writing 16-byte pieces (conv_igemm_kernel, tma == 2).
static int make_tmap_phase(CUtensorMap* m, const void* base, int nmax, int H, int W, int dst_w, int dst_plane, int pitch_elems, int box_pairs) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const uint64_t pw = (uint64_t)dst_w, plane = (uint64_t)dst_plane, eb = (uint64_t)pitch_elems * 2;   // the consumer level's stored geometry
    cuuint64_t dims[5] = {(cuuint64_t)pitch_elems, 2, (cuuint64_t)((W + 2) / 2), (cuuint64_t)((H + 2) / 2), (cuuint64_t)(3 * nmax)};
    cuuint64_t strides[4] = {(cuuint64_t)nmax * plane * eb, eb, pw * eb, plane * eb};
    cuuint32_t box[5] = {32, 2, (cuuint32_t)box_pairs, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled (5-D phase view) failed with CUresult %d (H=%d W=%d C=%d)", (int)r, H, W, pitch_elems);
    return FVY_OK;
}

// ------------------------------------------------------------------------------------------ handle
struct Layer {
    ConvSpec s;
    int Hin, Win, Hout, Wout;
    int BN, BK, stages, b_stages = 0, b_resident = 0, num_n_tiles, cout_pad, cin_pad, taps, occ;
    bool deep_k = false;
    int head_slot = -1;           // index of the logit tensor a head layer writes
    bool tap_perm = false;        // column taps stored in the order s = 0, 2, 1 (stride-2 slab pairs)
    int chain = -1, chain_pos = 0;   // index into fvy_handle::chains and position inside it (conv_chain_kernel), or -1
    // cross-layer tile dependencies (see ConvParams::sig_flags)
    bool signals = false;         // every stored form leaves by TMA and a consumer waits on the counters
    int wait_on = -1;             // index (in h->layers) of the producer whose counters gate this layer's tiles, or -1
    int* flags = nullptr;         // this layer's counters
    int flag_blocks = 0;
    bool cta2 = false;            // CTA pair (cta_group::2): 256-row tiles, each CTA stages half of the B tile
    size_t smem_bytes;
    __nv_bfloat16* w = nullptr;   // [cout_pad][taps * cin_pad]
    float* bias = nullptr;        // [cout_pad]
    CUtensorMap tmap_a, tmap_b, tmap_res, tmap_out[2];
    ConvParams p;                 // m_total / num_m_tiles filled per call
    OutDesc primary;              // where fvy_layer_output reads from
    size_t stream_off;            // offset of this layer in the Darknet stream
};

struct DevBuf {
    void* p = nullptr; size_t bytes = 0;
};

}  // namespace fvy

using namespace fvy;

struct fvy_handle {
    fvy_config cfg;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    std::vector<Layer> layers;
    std::vector<void*> allocs;
    bool weights_loaded = false;
    bool use_pdl = true;
    float* d_staged = nullptr;           // fvy_staged_images: [max_batch][net_h][net_w][3] float32, allocated on first use
    unsigned char* d_lb_src = nullptr; size_t lb_src_bytes = 0;   // letterbox source scratch
    int stem_blocks_per_sm = 3;          // resident blocks of stem_strip_kernel (occupancy query)
    __nv_bfloat16* d_stem_w2 = nullptr;  // conv_0 weights in stem_strip_kernel's K order (k = 10 r + 3 q + ci)
    const void* cur_img = nullptr; int cur_dtype = FVY_F32;   // device image of the current forward (layer 0 re-runs)
    // conv_0 + conv_1 in one kernel (stem_conv1_fused_kernel): conv_0's activation stays in shared memory.  Whole forwards only;
    // single-layer runs (profiling, fvy_layer_output of conv_0) use the separate kernels.
    bool fuse_stem = false; bool stem_phase_valid = false;
    CUtensorMap tmap_w1f; FuseParams fuse;
    // single-convolution handle (fvy_conv_create): one stride-1 layer, input packed from the caller's NHWC fp32 tensor, fp32 output
    bool conv_mode = false; int conv_cin = 0, conv_cout = 0, conv_k = 0, conv_stride = 1;
    __nv_bfloat16* d_conv_in = nullptr; float* conv_out = nullptr;
    long long launches = 0;
    long long weight_count = 0;
    // forward
    // Host images are staged through two device slots on a separate copy stream so that, with the async API, the
    // H2D copy of call i+1 overlaps the compute of call i; detections leave on a third stream.
    void* d_input[2] = {nullptr, nullptr}; size_t input_bytes = 0;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_post = nullptr, ev_d2h = nullptr;
    unsigned stage_slot = 0; int last_slot = -1;
    float* d_logits[3] = {nullptr, nullptr, nullptr};
    // Asynchronous detect calls post-process on their own (low-priority) stream: decode / NMS of call i fill the gaps that the
    // persistent conv kernels of call i+1 leave at layer boundaries.  Head logits alternate between two sets for that.
    float* d_logits_alt[3] = {nullptr, nullptr, nullptr};
    cudaStream_t post_stream = nullptr, ps = nullptr;       // ps: the stream post-processing is enqueued on for the current call
    cudaEvent_t ev_fwd_done[2] = {nullptr, nullptr}, ev_post_done[2] = {nullptr, nullptr};
    int logit_set = 0; bool overlap_post = true;
    // The conv stack of one forward is captured once per (batch, dtype, input pointer, logit set) into a CUDA graph (the PDL
    // edges between the layers are kept) and replayed: one graph launch instead of 75 kernel launches.
    struct GraphKey {
        int batch, dtype, set; const void* img;
        bool operator<(const GraphKey& o) const { return std::tie(batch, dtype, set, img) < std::tie(o.batch, o.dtype, o.set, o.img); }
    };
    std::map<GraphKey, cudaGraphExec_t> graphs;
    std::map<GraphKey, long long> graph_launches;   // kernels captured in each graph (what one replay launches)
    bool use_graph = true, capturing = false;
    struct Chain { int first = 0, count = 0; ChainLayer* dev = nullptr; std::vector<ChainLayer> host;
                   int* d_sched = nullptr; int sched_stride = 0; std::vector<int> sched_host; bool sched_on = false; };   // FVY_CHAIN_SCHED: per-pair work lists
    int chain_sched = 0;                 // 0 / 1 / 2 = auto (see prepare_chains)
    std::vector<Chain> chains; bool use_chain = true; int chain_batch = -1;
    int chain_nb = 3, chain_a = 4, chain_b = 6; size_t chain_smem = 0;
    int* d_flags = nullptr; size_t flags_bytes = 0; bool use_flags = true, flags_live = false;
    int gh[3] = {0, 0, 0}, gw[3] = {0, 0, 0}, head_c = 0;
    // post
    int cap = 0, capP = 0, words = 0, np2max = 0, smem_keys = 0;
    double* d_nbox = nullptr;
    int* d_ibox = nullptr; float* d_obj = nullptr; float* d_cls = nullptr; int* d_cand = nullptr; int* d_counts = nullptr;
    int* d_status = nullptr; size_t status_bytes = 16; int* d_image_hw = nullptr;
    int* d_order = nullptr; int4* d_sbox = nullptr; unsigned long long* d_mask = nullptr; unsigned long long* d_gkeys = nullptr;
    unsigned long long* d_rowflag = nullptr;
    uint4* d_srow = nullptr; unsigned char* d_sflag = nullptr;   // half-precision records / per-32 flags of the sorted boxes (nms_mask_kernel)
    int* d_kept = nullptr; int* d_kept_counts = nullptr;
    FvyDet* d_dets = nullptr; int* d_det_counts = nullptr; int dets_cap = 0;
    float last_fwd_ms = 0.f, last_post_ms = 0.f;
    // Asynchronous calls report decode errors late: the status word and the per-image candidate counts of each call are copied to
    // pinned host memory behind its post-processing (one block per logit set) and examined at the next synchronisation point
    // (fvy_sync, any synchronous entry point, or the next asynchronous call that reuses the set).
    // Device time of the forward / post-processing part of every detect call of a timed region (fvy_timer_start ..
    // fvy_timer_breakdown): event pairs in a small ring; a slot's elapsed times are added to the running sums when the slot comes
    // round again (kTimeRing calls later its events have long completed) or when the breakdown is read: a handle
    // holds 128 timing events however long the timed region is.
    static constexpr int kTimeRing = 32;
    cudaEvent_t ev_tf[kTimeRing][2] = {}, ev_tp[kTimeRing][2] = {};
    bool ring_used[kTimeRing] = {};
    double acc_fwd_ms = 0.0, acc_post_ms = 0.0; long long acc_calls = 0;
    long long call_count = 0;
    int time_slot = -1;                      // ring slot of the detect call being enqueued, or -1
    int* h_async[2] = {nullptr, nullptr};    // pinned: [0] status word, [1 .. max_batch] candidate counts
    int async_batch[2] = {0, 0};             // images of the set's pending call; 0 = nothing to examine
};

namespace fvy {

static int dev_alloc(fvy_handle* h, void** p, size_t bytes, bool zero) {
    CUDA_TRY(cudaMalloc(p, bytes ? bytes : 16));
    h->allocs.push_back(*p);
    if (zero) CUDA_TRY(cudaMemsetAsync(*p, 0, bytes ? bytes : 16, h->stream));
    return FVY_OK;
}

static bool is_device_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

static uint16_t f32_to_bf16_rn(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);   // NaN
    u += 0x7fffu + ((u >> 16) & 1u);
    return (uint16_t)(u >> 16);
}

}  // namespace fvy
