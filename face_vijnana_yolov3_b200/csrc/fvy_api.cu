// libfvy.so — C ABI (include/fvy.h) over the sm_100a kernels.  Host-side plan, weight folding,
// buffer arena, TMA tensor maps, launches.  No torch types, no CPU fallback.
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
static int make_tmap_phase(CUtensorMap* m, const void* base, int nmax, int H, int W, int pitch_elems, int box_pairs) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled entry point not available");
    const uint64_t pw = (uint64_t)(W / 2 + 2), plane = (uint64_t)(H / 2 + 2) * pw, eb = (uint64_t)pitch_elems * 2;
    cuuint64_t dims[5] = {(cuuint64_t)pitch_elems, 2, (cuuint64_t)((W + 2) / 2), (cuuint64_t)((H + 2) / 2), (cuuint64_t)(3 * nmax)};
    cuuint64_t strides[4] = {(cuuint64_t)nmax * plane * eb, eb, pw * eb, plane * eb};
    cuuint32_t box[5] = {32, 2, (cuuint32_t)box_pairs, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(FVY_E_CUDA, "cuTensorMapEncodeTiled (5-D phase view) failed with CUresult %d (H=%d W=%d C=%d)", (int)r, H, W, pitch_elems);
    return FVY_OK;
}

// ------------------------------------------------------------------------------------------ small kernels
// Stem operand: im2col of the 3x3 / pad 1 / stride 1 window of the RGB input, K index = (r*3+s)*3 + c,
// padded 27 -> 32 (one 64-byte swizzle row per pixel).  yolov3_detect.py:221 (conv_0), :205 (ZeroPadding2D(1)).
template <typename T>
__global__ void stem_im2col_kernel(const T* __restrict__ img, int batch, int H, int W, __nv_bfloat16* __restrict__ out) {
    const long long total = (long long)batch * H * W;
    for (long long pix = blockIdx.x * (long long)blockDim.x + threadIdx.x; pix < total; pix += (long long)gridDim.x * blockDim.x) {
        const int w = (int)(pix % W);
        const long long t = pix / W;
        const int h = (int)(t % H);
        const long long n = t / H;
        __align__(16) __nv_bfloat16 v[32];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int s = 0; s < 3; ++s) {
                const int hh = h + r - 1, ww = w + s - 1;
                const bool in = (unsigned)hh < (unsigned)H && (unsigned)ww < (unsigned)W;
                const T* src = img + ((n * H + (in ? hh : 0)) * W + (in ? ww : 0)) * 3;
#pragma unroll
                for (int c = 0; c < 3; ++c) v[(r * 3 + s) * 3 + c] = __float2bfloat16_rn(in ? (float)src[c] : 0.f);
            }
#pragma unroll
        for (int k = 27; k < 32; ++k) v[k] = __float2bfloat16_rn(0.f);
        uint4* d = reinterpret_cast<uint4*>(out + pix * 32);
        const uint4* s4 = reinterpret_cast<const uint4*>(v);
#pragma unroll
        for (int j = 0; j < 4; ++j) d[j] = s4[j];
    }
}


// Fused stem: conv_0 (3x3, pad 1, stride 1, 3 -> 32, BN folded, LeakyReLU(0.1)) straight from the fp32/fp64 image to the
// bf16 4-phase activation that conv_1 (stride 2) reads.  yolov3_detect.py:221 (conv_0), :205 (ZeroPadding2D(1)), :212-213.
// K = 27 and N = 32: 1.7 % of the network's FLOPs but its largest activation (11 MB / image), i.e. purely HBM-bound.
// A tcgen05 tile (128 x 32, K = 32) carries too little math per TMEM / mbarrier round trip (measured 284 us + 156 us for
// the im2col operand), so this layer keeps everything in registers: each warp owns 16 consecutive pixels of one image
// row, gathers its A fragment directly from the image (the 3x9 window rows are contiguous in NHWC), runs 8 warp-level
// bf16 MMAs (m16n8k16, fp32 accumulate), and transposes the result inside each quad so that every lane stores 16
// contiguous bytes.  Traffic: image read once (L1 catches the window overlap) + output written once.
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

#ifndef STEM_MIN_BLOCKS
#define STEM_MIN_BLOCKS 4
#endif
template <typename T>
__global__ void __launch_bounds__(256, STEM_MIN_BLOCKS) stem_conv_kernel(const T* __restrict__ img, int batch, int H, int W, int nmax,
                                                        const __nv_bfloat16* __restrict__ wgt /*[32][32] k-major*/,
                                                        const float* __restrict__ bias, __nv_bfloat16* __restrict__ out /*4-phase*/) {
    __shared__ __align__(16) uint32_t stile[8][16][20];     // per warp: 16 pixels x 32 bf16 (16 words) + 4 words of padding
    const int lane = threadIdx.x & 31, quad = lane & 3, grp = lane >> 2, wib = threadIdx.x >> 5;
    const int warps_total = gridDim.x * (blockDim.x >> 5);
    const int warp_id = blockIdx.x * (blockDim.x >> 5) + wib;
    // B fragments (weights) and bias stay in registers for the whole kernel
    uint32_t bfrag[4][2][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint32_t* wr = reinterpret_cast<const uint32_t*>(wgt + (j * 8 + grp) * 32 + t * 16 + quad * 2);
            bfrag[j][t][0] = __ldg(wr);
            bfrag[j][t][1] = __ldg(wr + 4);
        }
    float bia[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bia[j][0] = __ldg(bias + j * 8 + quad * 2); bia[j][1] = __ldg(bias + j * 8 + quad * 2 + 1); }
    const int tiles_per_row = W >> 4;
    const int total_tiles = batch * H * tiles_per_row;      // < 2^31 (checked on the host)
    const int pw = (W >> 1) + 2;      // phase planes have the padded geometry of the stride-2 conv's OUTPUT
    const long long plane = (long long)((H >> 1) + 2) * pw;
    // Per-lane gather table: the 8 K indices this lane feeds (k = ks*16 + half*8 + quad*2 + e) never change, so their
    // element offsets relative to the centre pixel and their edge sensitivities are computed once.
    int koff[8];
    unsigned m_top = 0, m_bot = 0, m_left = 0, m_right = 0, m_none = 0;   // bit i: load i must be skipped at that image edge
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int k = (i >> 2) * 16 + ((i >> 1) & 1) * 8 + quad * 2 + (i & 1);
        const int r = k / 9, j = k - r * 9, dx = j / 3 - 1, dr = r - 1, c = j - (j / 3) * 3;
        koff[i] = k < 27 ? (dr * W + dx) * 3 + c : 0;       // k >= 27: any valid address (the weight row is zero and the value is masked)
        if (k >= 27) m_none |= 1u << i;
        if (dr < 0) m_top |= 1u << i;
        if (dr > 0) m_bot |= 1u << i;
        if (dx < 0) m_left |= 1u << i;
        if (dx > 0) m_right |= 1u << i;
    }
    for (int tile = warp_id; tile < total_tiles; tile += warps_total) {
        const int t2 = tile / tiles_per_row;
        const int tx = tile - t2 * tiles_per_row;
        const int n = t2 / H;
        const int h = t2 - n * H;
        const int w0 = tx << 4;
        // A fragment: rows grp and grp+8 of the tile; register (ks, half*2 + rr) holds k = ks*16 + half*8 + quad*2 + {0,1}
        uint32_t afrag[2][4];
        const unsigned skip_h = m_none | (h == 0 ? m_top : 0u) | (h == H - 1 ? m_bot : 0u);
        const T* centre0 = img + ((long long)(n * H + h) * W + w0 + grp) * 3;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int wpix = w0 + grp + rr * 8;
            const unsigned skip = skip_h | (wpix == 0 ? m_left : 0u) | (wpix == W - 1 ? m_right : 0u);
            const T* centre = centre0 + rr * 24;
            float v[8];
            if (skip == m_none) {       // interior pixel (the common case): unconditional loads
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = (float)__ldg(centre + koff[i]);
#pragma unroll
                for (int i = 0; i < 8; ++i) if ((m_none >> i) & 1u) v[i] = 0.f;
            } else {
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = ((skip >> i) & 1u) ? 0.f : (float)__ldg(centre + koff[i]);
            }
#pragma unroll
            for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    __nv_bfloat162 pk = __floats2bfloat162_rn(v[ks * 4 + half * 2], v[ks * 4 + half * 2 + 1]);
                    afrag[ks][half * 2 + rr] = *reinterpret_cast<uint32_t*>(&pk);
                }
        }
        // accumulators start at the bias (BN folded): saves the separate add
        float acc[4][4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            acc[j][0] = acc[j][2] = bia[j][0];
            acc[j][1] = acc[j][3] = bia[j][1];
            mma_m16n8k16_bf16(acc[j], afrag[0], bfrag[j][0][0], bfrag[j][0][1]);
            mma_m16n8k16_bf16(acc[j], afrag[1], bfrag[j][1][0], bfrag[j][1][1]);
        }
        // LeakyReLU, pack, transpose through shared memory: lane (grp, quad) then owns 16 contiguous bytes of pixel grp / grp+8
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float a = acc[j][rr * 2 + 0], b = acc[j][rr * 2 + 1];
                a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b);
                __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
                stile[wib][grp + rr * 8][j * 4 + quad] = *reinterpret_cast<uint32_t*>(&pk);
            }
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const uint4 o = *reinterpret_cast<const uint4*>(&stile[wib][grp + rr * 8][quad * 4]);
            const int wpix = w0 + grp + rr * 8;
            const int hp = h + 1, wp = wpix + 1;
            const long long row = ((long long)((((hp & 1) << 1) | (wp & 1))) * nmax + n) * plane + (long long)(hp >> 1) * pw + (wp >> 1);
            *reinterpret_cast<uint4*>(out + row * 32 + quad * 8) = o;
        }
    }
}

// Second form of the fused stem (default): the image rows a strip of outputs needs are staged ONCE in shared memory as bf16
// (rolling window of three rows: every input element is read from HBM once, coalesced, and converted once instead of nine
// times), and the MMA fragments are gathered from there with aligned 32-bit loads.  The K order is chosen for that: filter
// row r occupies k = 10 r .. 10 r + 9 (nine contiguous window elements (dx, c) of the NHWC row + one zero slot), so every
// fragment register (k, k+1) is two adjacent bf16 of a staged row; a second copy of each row shifted by one element makes
// the pair 4-byte aligned for odd pixels as well.  Image borders are zeros in the staged rows: no edge tests in the loop.
constexpr int kStemThreads = 416;      // 13 warps: 26 (416 px) / 38 (608 px) 16-pixel tiles per image row
template <typename T>
__global__ void __launch_bounds__(kStemThreads, 2)
stem_rows_kernel(const T* __restrict__ img, int batch, int H, int W, int nmax, const __nv_bfloat16* __restrict__ wgt /*[32][32], k = 10 r + j*/,
                 const float* __restrict__ bias, __nv_bfloat16* __restrict__ out /*4-phase*/, int rows_per_block, int rowlen) {
    extern __shared__ __align__(16) uint16_t srows[];        // [8 slots][2 copies][rowlen]
    __shared__ __align__(16) uint32_t stile[kStemThreads / 32][16][20];
    const int lane = threadIdx.x & 31, quad = lane & 3, grp = lane >> 2, wib = threadIdx.x >> 5;
    const int nwarps = blockDim.x >> 5;
    const long long total_rows = (long long)batch * H;
    const long long row0 = (long long)blockIdx.x * rows_per_block;
    const long long row1 = min(total_rows, row0 + rows_per_block);
    if (row0 >= row1) return;
    uint32_t bfrag[4][2][2];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const uint32_t* wr = reinterpret_cast<const uint32_t*>(wgt + (j * 8 + grp) * 32 + t * 16 + quad * 2);
            bfrag[j][t][0] = __ldg(wr);
            bfrag[j][t][1] = __ldg(wr + 4);
        }
    float bia[4][2];
#pragma unroll
    for (int j = 0; j < 4; ++j) { bia[j][0] = __ldg(bias + j * 8 + quad * 2); bia[j][1] = __ldg(bias + j * 8 + quad * 2 + 1); }
    // this lane's four (k, k+1) pairs of a pixel: k = ks*16 + half*8 + quad*2 -> filter row r = k / 10, window element j = k % 10
    int pr[4], pj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = (i >> 1) * 16 + (i & 1) * 8 + quad * 2;
        pr[i] = k < 30 ? k / 10 : 0;             // k = 30, 31: zero weights, any valid address
        pj[i] = k < 30 ? k % 10 : 0;
    }
    const int par = grp & 1;                     // parity of this lane's pixels (w0 and rr*8 are even): which copy gives aligned pairs
    const int pw = (W >> 1) + 2;
    const long long plane = (long long)((H >> 1) + 2) * pw;
    const int tiles_per_row = W >> 4;
    const int n_elems = (W + 2) * 3 + 2;
    auto load_row = [&](int slot, long long n, int hh) {
        uint16_t* E = srows + (size_t)(slot * 2) * rowlen;
        uint16_t* O = E + rowlen;
        const bool in = hh >= 0 && hh < H;
        const T* src = img + ((n * H + (in ? hh : 0)) * W) * 3;
        for (int e = threadIdx.x; e < n_elems; e += blockDim.x) {
            float v = 0.f;
            if (in && e >= 3 && e < (W + 1) * 3) v = (float)__ldg(src + (e - 3));
            const uint16_t b = __bfloat16_as_ushort(__float2bfloat16_rn(v));
            E[e] = b;
            if (e >= 1) O[e - 1] = b;
        }
    };
    // Two output rows per block-wide barrier; software pipeline: the loads of the next two input rows are in flight while the
    // current two output rows are computed, and are converted and stored (ring of 8 row slots) after them.
    constexpr int kFetch = 5;                    // ceil(((608 + 2) * 3 + 2) / 416)
    float pf[2][kFetch];
    auto fetch_row = [&](int which, long long n, int hh) {
        const bool in = hh >= 0 && hh < H;
        const T* src = img + ((n * H + (in ? hh : 0)) * W) * 3;
#pragma unroll
        for (int k = 0; k < kFetch; ++k) {
            const int e = threadIdx.x + k * kStemThreads;
            pf[which][k] = (in && e >= 3 && e < (W + 1) * 3) ? (float)__ldg(src + (e - 3)) : 0.f;
        }
    };
    auto store_row = [&](int which, int slot) {
        uint16_t* E = srows + (size_t)(slot * 2) * rowlen;
        uint16_t* O = E + rowlen;
#pragma unroll
        for (int k = 0; k < kFetch; ++k) {
            const int e = threadIdx.x + k * kStemThreads;
            if (e < n_elems) {
                const uint16_t b = __bfloat16_as_ushort(__float2bfloat16_rn(pf[which][k]));
                E[e] = b;
                if (e >= 1) O[e - 1] = b;
            }
        }
    };
    long long prev_n = -1; int prev_h = -3;
    for (long long row = row0; row < row1; row += 2) {        // row ranges start on even rows and H is even: (h, h+1) share an image
        const long long n = row / H;
        const int h = (int)(row - n * H);
        if (!(n == prev_n && h == prev_h + 2)) {
            __syncthreads();                     // a new window: nobody may still be reading the slots
            for (int d = -1; d <= 2; ++d) load_row((h + d) & 7, n, h + d);
            __syncthreads();
        }
        prev_n = n; prev_h = h;
        fetch_row(0, n, h + 3);
        fetch_row(1, n, h + 4);
        for (int item = wib; item < 2 * tiles_per_row; item += nwarps) {
            const int second = item >= tiles_per_row ? 1 : 0;
            const int hh = h + second;
            const int w0 = (item - second * tiles_per_row) << 4;
            uint32_t afrag[2][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const uint16_t* rp = srows + (size_t)((((hh - 1 + pr[i]) & 7) * 2 + par)) * rowlen + pj[i] - par;
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                    afrag[i >> 1][(i & 1) * 2 + rr] = *reinterpret_cast<const uint32_t*>(rp + (w0 + grp + rr * 8) * 3);
            }
            float acc[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                acc[j][0] = acc[j][2] = bia[j][0];
                acc[j][1] = acc[j][3] = bia[j][1];
                mma_m16n8k16_bf16(acc[j], afrag[0], bfrag[j][0][0], bfrag[j][0][1]);
                mma_m16n8k16_bf16(acc[j], afrag[1], bfrag[j][1][0], bfrag[j][1][1]);
            }
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    float a = acc[j][rr * 2 + 0], b = acc[j][rr * 2 + 1];
                    a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b);
                    __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
                    stile[wib][grp + rr * 8][j * 4 + quad] = *reinterpret_cast<uint32_t*>(&pk);
                }
            __syncwarp();
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const uint4 o = *reinterpret_cast<const uint4*>(&stile[wib][grp + rr * 8][quad * 4]);
                const int wpix = w0 + grp + rr * 8;
                const int hp = hh + 1, wp = wpix + 1;
                const long long orow = ((long long)((((hp & 1) << 1) | (wp & 1))) * nmax + n) * plane + (long long)(hp >> 1) * pw + (wp >> 1);
                *reinterpret_cast<uint4*>(out + orow * 32 + quad * 8) = o;
            }
        }
        store_row(0, (h + 3) & 7);               // slots of rows h-5, h-4: last read two iterations ago
        store_row(1, (h + 4) & 7);
        __syncthreads();
    }
}

// Debug / parity aid: stored activation -> dense NHWC fp32.
__global__ void unpack_kernel(OutDesc od, int batch, int H, int W, int C, float* __restrict__ dst) {
    const long long total = (long long)batch * H * W * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        long long t = i / C;
        const int w = (int)(t % W); t /= W;
        const int h = (int)(t % H);
        const long long n = t / H;
        float v;
        if (od.kind == OUT_HEAD_F32) {
            v = reinterpret_cast<const float*>(od.ptr)[((n * H + h) * W + w) * od.c_real + c];
        } else {
            long long row;
            if (od.kind == OUT_PADDED) row = (n * (H + 2) + (h + 1)) * (W + 2) + (w + 1);
            else if (od.kind == OUT_PHASE) {
                const int hp = h + 1, wp = w + 1, ph = ((hp & 1) << 1) | (wp & 1), pw = (W >> 1) + 2;
                const long long plane = (long long)((H >> 1) + 2) * pw;
                row = ((long long)ph * od.nmax + n) * plane + (long long)(hp >> 1) * pw + (wp >> 1);
            } else row = (n * (2 * H + 2) + (2 * h + 1)) * (2 * W + 2) + (2 * w + 1);
            v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(od.ptr)[row * od.pitch + od.choff + c]);
        }
        dst[i] = v;
    }
}

// Third form of the fused stem (default, FVY_STEM=3): the same staged-row scheme and K order as stem_rows_kernel, but every
// WARP owns a strip of 16 output columns over a segment of rows and keeps its own ring of staged rows (18 pixels x 3 channels,
// two copies) - no block-wide barrier at all.  stem_rows_kernel spent 35 % of its stall samples in the two-row barrier; here a
// warp only ever waits for its own loads (two rows ahead, in registers while the current row is computed), and 24 independent
// warps per SM hide each other's latency.  The (image, strip, row) space is cut into one equal piece per resident warp.  Halo columns are re-read by the neighbouring strip (12 %, L1 / L2 hits).
__device__ __forceinline__ float stem_px(float v) { return v; }
__device__ __forceinline__ float stem_px(double v) { return (float)v; }                         // Keras casts its input to float32
__device__ __forceinline__ float stem_px(unsigned char v) { return (float)((double)v / 255.0); }   // image / 255 in float64, then that cast
constexpr int kStripWarps = 8;
constexpr int kStripLen = 64;          // staged elements per row copy: 18 pixels x 3 channels = 54, padded
#ifndef FVY_STRIP_MINB
#define FVY_STRIP_MINB 4
#endif
template <typename T>
__global__ void __launch_bounds__(kStripWarps * 32, FVY_STRIP_MINB)
stem_strip_kernel(const T* __restrict__ img, int batch, int H, int W, int nmax, const __nv_bfloat16* __restrict__ wgt /*[32][32], k = 10 r + j*/,
                  const float* __restrict__ bias, __nv_bfloat16* __restrict__ out /*4-phase*/) {
    __shared__ __align__(16) uint16_t srows[kStripWarps][8][2][kStripLen];
    __shared__ __align__(16) uint32_t stile[kStripWarps][16][20];
    const int lane = threadIdx.x & 31, quad = lane & 3, grp = lane >> 2, wib = threadIdx.x >> 5;
    // weight / bias fragments live in shared memory, one 16-byte + one 8-byte entry per (n-tile, lane): 24 registers less per thread
    // buys the fourth resident block per SM (with them in registers: 80 registers and three blocks, or spills)
    __shared__ __align__(16) uint4 swf[4][32];
    __shared__ __align__(8) float2 sbf[4][32];
    if (wib < 4) {
        const int j = wib;
        uint4 f;
        const uint32_t* wr0 = reinterpret_cast<const uint32_t*>(wgt + (j * 8 + grp) * 32 + quad * 2);
        f.x = __ldg(wr0); f.y = __ldg(wr0 + 4); f.z = __ldg(wr0 + 8); f.w = __ldg(wr0 + 12);
        swf[j][lane] = f;
        sbf[j][lane] = make_float2(__ldg(bias + j * 8 + quad * 2), __ldg(bias + j * 8 + quad * 2 + 1));
    }
    __syncthreads();
    // work = (image, strip, row) triples in that order; every warp of the grid takes one contiguous, equally long piece of it
    // (a piece may continue in the next strip / image: the ring of staged rows is simply primed again there)
    const int strips = W >> 4;
    const long long total = (long long)batch * strips * H;
    const long long nwarps = (long long)gridDim.x * kStripWarps;
    const long long piece = (total + nwarps - 1) / nwarps;
    long long pos = ((long long)blockIdx.x * kStripWarps + wib) * piece;
    const long long pos_end = min(total, pos + piece);
    if (pos >= pos_end) return;
    // this lane's four (k, k+1) pairs of a pixel: k = ks*16 + half*8 + quad*2 -> filter row r = k / 10, window element j = k % 10
    int pr[4], pj[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int k = (i >> 1) * 16 + (i & 1) * 8 + quad * 2;
        pr[i] = k < 30 ? k / 10 : 0;             // k = 30, 31: zero weights, any valid address
        pj[i] = k < 30 ? k % 10 : 0;
    }
    const int par = grp & 1;                     // parity of this lane's pixels: which copy gives 4-byte aligned pairs
    const int pw = (W >> 1) + 2;
    const long long plane = (long long)((H >> 1) + 2) * pw;
    while (pos < pos_end) {
    const long long col = pos / H;                           // (image, strip) column of this run of rows
    const int h0 = (int)(pos - col * H);
    const int h1 = (int)min((long long)H, (long long)h0 + (pos_end - pos));
    const long long n = col / strips;
    const int w0 = (int)(col - n * strips) << 4;
    pos += h1 - h0;
    // staged element i of a strip row = image element (w0 - 1) * 3 + i of that row (zero outside the image); lanes own i = lane, lane + 32
    const int c0 = (w0 - 1) * 3 + lane, c1 = c0 + 32;
    const bool ok0 = c0 >= 0 && c0 < W * 3, ok1 = lane + 32 < 54 && c1 < W * 3;
    const T* base = img + (n * H) * (long long)W * 3;
    auto fetch = [&](int hh, float& a, float& b) {
        const bool in = hh >= 0 && hh < H;
        const T* src = base + (long long)(in ? hh : 0) * W * 3;
        a = (in && ok0) ? stem_px(__ldg(src + c0)) : 0.f;
        b = (in && ok1) ? stem_px(__ldg(src + c1)) : 0.f;
    };
    auto stage = [&](int hh, float a, float b) {
        uint16_t* E = srows[wib][hh & 7][0];
        uint16_t* O = srows[wib][hh & 7][1];
        const uint16_t x = __bfloat16_as_ushort(__float2bfloat16_rn(a)), y = __bfloat16_as_ushort(__float2bfloat16_rn(b));
        E[lane] = x; E[lane + 32] = y;
        if (lane >= 1) O[lane - 1] = x;
        O[lane + 31] = y;
    };
    float p0a, p0b, p1a, p1b;                    // the two rows in flight
    {
        float a, b;
        fetch(h0 - 1, a, b); stage(h0 - 1, a, b);
        fetch(h0, a, b); stage(h0, a, b);
        fetch(h0 + 1, a, b); stage(h0 + 1, a, b);
        fetch(h0 + 2, p0a, p0b);
        fetch(h0 + 3, p1a, p1b);
    }
    __syncwarp();
    for (int h = h0; h < h1; ++h) {
        uint32_t afrag[2][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint16_t* rp = srows[wib][(h - 1 + pr[i]) & 7][par] + pj[i] - par;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr)
                afrag[i >> 1][(i & 1) * 2 + rr] = *reinterpret_cast<const uint32_t*>(rp + (grp + rr * 8) * 3);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint4 wf = swf[j][lane];
            const float2 bj = sbf[j][lane];
            float acc[4] = {bj.x, bj.y, bj.x, bj.y};
            mma_m16n8k16_bf16(acc, afrag[0], wf.x, wf.y);
            mma_m16n8k16_bf16(acc, afrag[1], wf.z, wf.w);
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                float a = acc[rr * 2 + 0], b = acc[rr * 2 + 1];
                a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b);
                __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
                stile[wib][grp + rr * 8][j * 4 + quad] = *reinterpret_cast<uint32_t*>(&pk);
            }
        }
        // row h + 2 (loaded two iterations ago) goes into the slot of row h - 6; then the next load is issued
        stage(h + 2, p0a, p0b);
        p0a = p1a; p0b = p1b;
        fetch(h + 4, p1a, p1b);
        __syncwarp();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const uint4 o = *reinterpret_cast<const uint4*>(&stile[wib][grp + rr * 8][quad * 4]);
            const int wpix = w0 + grp + rr * 8;
            const int hp = h + 1, wp = wpix + 1;
            const long long orow = ((long long)((((hp & 1) << 1) | (wp & 1))) * nmax + n) * plane + (long long)(hp >> 1) * pw + (wp >> 1);
            *reinterpret_cast<uint4*>(out + orow * 32 + quad * 8) = o;
        }
        __syncwarp();
    }
    }
}

// ------------------------------------------------------------------------------------------ letterbox (face_detection.py:657-690)
// cv.resize(image / 255, (w_p, h_p), INTER_CUBIC) + zero border, restated operation for operation (OpenCV resizeGeneric_ for
// CV_64F: interpolateCubic with A = -0.75 in float, HResizeCubic then VResizeCubic accumulating in double left to right,
// indices clamped to the image).  One thread per output pixel, three channels; -fmad=false keeps every product and sum
// separately rounded as the C++ reference computes them.
__device__ __forceinline__ void cubic_tab(int d, double scale, int src, int idx[4], float c[4]) {
    const float f0 = (float)(((double)d + 0.5) * scale - 0.5);
    const int s = (int)floorf(f0);
    const float x = f0 - (float)s;
    const float A = -0.75f;
    c[0] = ((A * (x + 1.f) - 5.f * A) * (x + 1.f) + 8.f * A) * (x + 1.f) - 4.f * A;
    c[1] = ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f;
    c[2] = ((A + 2.f) * (1.f - x) - (A + 3.f)) * (1.f - x) * (1.f - x) + 1.f;
    c[3] = 1.f - c[0] - c[1] - c[2];
#pragma unroll
    for (int k = 0; k < 4; ++k) idx[k] = min(max(s + k - 1, 0), src - 1);
}
__global__ void letterbox_u8_kernel(const unsigned char* __restrict__ src, int src_h, int src_w, int w_p, int h_p, int pad_t, int pad_l,
                                    int net_h, int net_w, float* __restrict__ dst) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= net_w) return;
    float* o = dst + ((size_t)y * net_w + x) * 3;
    const int dx = x - pad_l, dy = y - pad_t;
    if (dx < 0 || dx >= w_p || dy < 0 || dy >= h_p) { o[0] = o[1] = o[2] = 0.f; return; }
    int xi[4], yi[4];
    float xa[4], ya[4];
    cubic_tab(dx, 1.0 / ((double)w_p / (double)src_w), src_w, xi, xa);
    cubic_tab(dy, 1.0 / ((double)h_p / (double)src_h), src_h, yi, ya);
    double acc[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const unsigned char* row = src + (size_t)yi[k] * src_w * 3;
        double r[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double v = ((double)row[xi[0] * 3 + c] / 255.0) * (double)xa[0];
            v = v + ((double)row[xi[1] * 3 + c] / 255.0) * (double)xa[1];
            v = v + ((double)row[xi[2] * 3 + c] / 255.0) * (double)xa[2];
            v = v + ((double)row[xi[3] * 3 + c] / 255.0) * (double)xa[3];
            r[c] = v * (double)ya[k];
        }
#pragma unroll
        for (int c = 0; c < 3; ++c) acc[c] = k == 0 ? r[c] : acc[c] + r[c];
    }
    o[0] = (float)acc[0]; o[1] = (float)acc[1]; o[2] = (float)acc[2];
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
    bool fused_stem = true;              // conv_0 straight from the image (stem_conv_kernel) instead of im2col + GEMM
    float* d_staged = nullptr;           // fvy_staged_images: [max_batch][net_h][net_w][3] float32, allocated on first use
    unsigned char* d_lb_src = nullptr; size_t lb_src_bytes = 0;   // letterbox source scratch
    int stem_blocks_per_sm = 3;          // resident blocks of stem_strip_kernel (occupancy query)
    int stem_mode = 2;                   // 3: stem_strip_kernel (per-warp strips), 2: stem_rows_kernel (staged rows), 1: stem_conv_kernel (register gather)
    __nv_bfloat16* d_stem_w2 = nullptr;  // conv_0 weights in stem_rows_kernel's K order
    const void* cur_img = nullptr; int cur_dtype = FVY_F32;   // device image of the current forward (layer 0 re-runs)
    long long launches = 0;
    long long weight_count = 0;
    // forward
    // Host images are staged through two device slots on a separate copy stream so that, with the async API, the
    // H2D copy of call i+1 overlaps the compute of call i; detections leave on a third stream.
    void* d_input[2] = {nullptr, nullptr}; size_t input_bytes = 0;
    cudaStream_t h2d_stream = nullptr, d2h_stream = nullptr;
    cudaEvent_t ev_ready[2] = {nullptr, nullptr}, ev_consumed[2] = {nullptr, nullptr}, ev_post = nullptr, ev_d2h = nullptr;
    unsigned stage_slot = 0; int last_slot = -1;
    __nv_bfloat16* d_stem = nullptr;                        // im2col operand
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
                   int* d_sched = nullptr; int sched_stride = 0; std::vector<int> sched_host; };   // FVY_CHAIN_SCHED: per-pair work lists
    bool chain_sched = false;
    std::vector<Chain> chains; bool use_chain = true; int chain_batch = -1;
    int chain_nb = 3, chain_a = 4, chain_b = 6; size_t chain_smem = 0;
    int* d_flags = nullptr; size_t flags_bytes = 0; bool use_flags = true, flags_live = false;
    int gh[3] = {0, 0, 0}, gw[3] = {0, 0, 0}, head_c = 0;
    // post
    int cap = 0, capP = 0, words = 0, np2max = 0, smem_keys = 0;
    double* d_nbox = nullptr;
    int* d_ibox = nullptr; float* d_obj = nullptr; float* d_cls = nullptr; int* d_cand = nullptr; int* d_counts = nullptr;
    int* d_status = nullptr; int* d_image_hw = nullptr;
    int* d_order = nullptr; int4* d_sbox = nullptr; unsigned long long* d_mask = nullptr; unsigned long long* d_gkeys = nullptr;
    unsigned long long* d_rowflag = nullptr;
    int* d_kept = nullptr; int* d_kept_counts = nullptr;
    FvyDet* d_dets = nullptr; int* d_det_counts = nullptr; int dets_cap = 0;
    float last_fwd_ms = 0.f, last_post_ms = 0.f;
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

template <int BN, int BK, bool CTA2>
static int launch_conv_t(fvy_handle* h, Layer& L, int grid) {
    auto kern = conv_igemm_kernel<BN, BK, CTA2>;   // max dynamic smem was raised in query_occ_t at plan time
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = L.smem_bytes; cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (h->use_pdl) {   // PDL: prologue overlaps the previous layer's tail
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (CTA2) {
        at[na].id = cudaLaunchAttributeClusterDimension;
        at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
        ++na;
    }
    cfg.attrs = at; cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, L.tmap_a, L.tmap_b, L.tmap_res, L.tmap_out[0], L.tmap_out[1], L.p));
    ++h->launches;
    return FVY_OK;
}

template <int BN, int BK, bool CTA2>
static int query_occ_t(size_t smem, int* occ) {
    auto kern = conv_igemm_kernel<BN, BK, CTA2>;
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    if constexpr (CTA2) {
        *occ = 1;       // a cluster of two: one CTA per SM by construction
    } else {
        CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(occ, kern, kThreads, smem));
    }
    return FVY_OK;
}

#define FVY_DISPATCH(BNv, BKv, CALL)                                                  \
    do {                                                                              \
        if (BKv == 64) {                                                              \
            switch (BNv) {                                                            \
                case 32: return CALL(32, 64, false); case 64: return CALL(64, 64, false);           \
                case 128: return CALL(128, 64, false); case 256: return CALL(256, 64, false);       \
            }                                                                         \
        } else {                                                                      \
            switch (BNv) {                                                            \
                case 32: return CALL(32, 32, false); case 64: return CALL(64, 32, false);           \
                case 128: return CALL(128, 32, false); case 256: return CALL(256, 32, false);       \
            }                                                                         \
        }                                                                             \
        return fail(FVY_E_INVALID, "no kernel instance for tile N=%d K=%d", BNv, BKv); \
    } while (0)

static int launch_conv(fvy_handle* h, Layer& L, int grid) {
    if (L.cta2) return L.BN == 256 ? launch_conv_t<256, 64, true>(h, L, grid) : launch_conv_t<128, 64, true>(h, L, grid);
#define CALL(bn, bk, c2) launch_conv_t<bn, bk, c2>(h, L, grid)
    FVY_DISPATCH(L.BN, L.BK, CALL);
#undef CALL
}
static int query_occ(int BN, int BK, bool cta2, size_t smem, int* occ) {
    if (cta2) return BN == 256 ? query_occ_t<256, 64, true>(smem, occ) : query_occ_t<128, 64, true>(smem, occ);
#define CALL(bn, bk, c2) query_occ_t<bn, bk, c2>(smem, occ)
    FVY_DISPATCH(BN, BK, CALL);
#undef CALL
}

struct TensorBufs { __nv_bfloat16* padded = nullptr; __nv_bfloat16* phase = nullptr; };

static int build_plan(fvy_handle* h) {
    const fvy_config& c = h->cfg;
    std::vector<ConvSpec> specs = c.head == FVY_HEAD_YOLO3 ? yolo3_table(c.nb_class) : fd6_table(c.bb_info_c_size);
    const int nmax = c.max_batch;
    // which stored forms does each producer need?
    std::map<int, bool> need_padded, need_phase;
    for (const ConvSpec& s : specs) {
        if (s.src >= 0) { if (s.k == 3 && s.stride == 2) need_phase[s.src] = true; else need_padded[s.src] = true; }
        if (s.res >= 0) need_padded[s.res] = true;
    }
    std::map<int, const ConvSpec*> by_idx;
    for (const ConvSpec& s : specs) by_idx[s.idx] = &s;
    std::map<int, TensorBufs> bufs;
    auto HW = [&](int level, int* H, int* W) { *H = c.net_h >> level; *W = c.net_w >> level; };
    // concat buffers (yolo3 only): A = [up(conv_84) 256 | skip_61 512] at level 4, B = [up(conv_96) 128 | skip_36 256] at level 3
    __nv_bfloat16 *catA = nullptr, *catB = nullptr;
    if (c.head == FVY_HEAD_YOLO3) {
        int H, W;
        HW(4, &H, &W);
        if (int e = dev_alloc(h, (void**)&catA, (size_t)nmax * (H + 2) * (W + 2) * 768 * 2, true)) return e;
        HW(3, &H, &W);
        if (int e = dev_alloc(h, (void**)&catB, (size_t)nmax * (H + 2) * (W + 2) * 384 * 2, true)) return e;
    }
    // stem operand
    {
        const char* v = getenv("FVY_FUSED_STEM");
        h->fused_stem = !(v && *v && atoi(v) == 0);
        const char* m = getenv("FVY_STEM");
        h->stem_mode = (m && *m) ? atoi(m) : 3;
        if (c.net_w % 16) h->stem_mode = 1;
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, stem_strip_kernel<float>, kStripWarps * 32, 0) == cudaSuccess && occ > 0) h->stem_blocks_per_sm = 2 * occ;   // two pieces per resident warp (161 vs 166 us)
        if (const char* sb = getenv("FVY_STEM_BLOCKS")) if (*sb) h->stem_blocks_per_sm = atoi(sb);
    }
    if (int e = dev_alloc(h, (void**)&h->d_stem_w2, 32 * 32 * 2, true)) return e;
    {   // staged rows of stem_rows_kernel: 8 slots x 2 copies x (3 (W + 2) + 2) bf16 - beyond the 48 KB default for wide images
        const size_t need = (size_t)8 * 2 * ((((c.net_w + 2) * 3 + 2) + 7) & ~7) * 2;
        if (need > 100 * 1024) { if (h->stem_mode == 2) h->stem_mode = 1; }
        else {                                   // static (transpose tiles) + dynamic exceed the 48 KB default
            CUDA_TRY(cudaFuncSetAttribute(stem_rows_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
            CUDA_TRY(cudaFuncSetAttribute(stem_rows_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need));
        }
    }
    if (!h->fused_stem)
        if (int e = dev_alloc(h, (void**)&h->d_stem, (size_t)nmax * c.net_h * c.net_w * 32 * 2, false)) return e;
    // activation buffers
    for (const ConvSpec& s : specs) {
        int H, W;
        HW(s.level, &H, &W);
        TensorBufs tb;
        if (need_padded.count(s.idx))
            if (int e = dev_alloc(h, (void**)&tb.padded, (size_t)nmax * (H + 2) * (W + 2) * s.cout * 2, true)) return e;
        if (need_phase.count(s.idx))
            if (int e = dev_alloc(h, (void**)&tb.phase, (size_t)4 * nmax * (H / 2 + 2) * (W / 2 + 2) * s.cout * 2, true)) return e;
        bufs[s.idx] = tb;
    }
    // heads
    int nh = 0;
    for (const ConvSpec& s : specs)
        if (!s.bn) {
            int H, W;
            HW(s.level, &H, &W);
            if (nh >= 3) return fail(FVY_E_INVALID, "more than 3 heads");
            h->gh[nh] = H; h->gw[nh] = W; h->head_c = s.cout;
            if (int e = dev_alloc(h, (void**)&h->d_logits[nh], (size_t)nmax * H * W * s.cout * 4, true)) return e;
            if (int e = dev_alloc(h, (void**)&h->d_logits_alt[nh], (size_t)nmax * H * W * s.cout * 4, true)) return e;
            ++nh;
        }
    auto env_int = [](const char* name, int dflt) { const char* v = getenv(name); return v && *v ? atoi(v) : dflt; };
    const int bn_cap = c.tile_n_max > 0 ? c.tile_n_max : env_int("FVY_BN", 256);
    const int bn_res_cap = std::min(bn_cap, env_int("FVY_BN_RES", 256));
    const int nb_res = env_int("FVY_NB_RES", 4), nb_plain = env_int("FVY_NB", 3);
    const int stages_cap = env_int("FVY_STAGES", kMaxA);
    const int groups_kn = env_int("FVY_GROUPS_KN", 150);   // (K iterations x 32-column chunks) at or below which a layer gets two epilogue groups
    h->use_pdl = env_int("FVY_PDL", 1) != 0;
    size_t stream_off = 0;
    int head_i = 0;
    for (const ConvSpec& s : specs) {
        Layer L;
        L.s = s;
        HW(s.level, &L.Hout, &L.Wout);
        L.Hin = L.Hout * s.stride; L.Win = L.Wout * s.stride;
        L.stream_off = stream_off;
        stream_off += (size_t)(s.bn ? 4 : 1) * s.cout + (size_t)s.cout * s.cin * s.k * s.k;
        const bool stem = s.src == -1;
        L.taps = stem ? 1 : s.k * s.k;
        L.cin_pad = stem ? 32 : s.cin;
        L.BK = (L.cin_pad % 64 == 0) ? 64 : 32;
        if (L.cin_pad % L.BK) return fail(FVY_E_INVALID, "conv_%d: Cin %d not a multiple of %d", s.idx, s.cin, L.BK);
        L.cout_pad = (s.cout + 31) / 32 * 32;
        if (L.cout_pad > kMaxCout) return fail(FVY_E_INVALID, "conv_%d: Cout %d exceeds %d", s.idx, s.cout, kMaxCout);
        const bool has_res = s.res >= 0;
        int cap_n = has_res ? bn_res_cap : bn_cap;
        if (!stem && s.k == 3 && s.stride == 1 && s.cin >= 128) cap_n = std::min(cap_n, env_int("FVY_BN3", 256));   // deep 3x3 layers: narrower tiles = finer waves
        L.BN = 32;
        for (int bn : {256, 128, 64, 32})
            if (bn <= cap_n && L.cout_pad % bn == 0) { L.BN = bn; break; }
        L.num_n_tiles = L.cout_pad / L.BN;
        const int gt = L.taps == 9 ? 3 : 1;                       // column taps per filter row
        // CTA pairs (cta_group::2, 256-row tiles, each CTA stages half of the B tile): every 256-wide layer, and the
        // 128-wide 3x3 layers whose whole weight tile then fits in shared memory (conv_5/7/10)
        L.cta2 = !stem && L.BK == 64 && env_int("FVY_CTA2", 1) != 0 &&
                 (L.BN == 256 || (L.BN == 128 && L.taps == 9 && (L.num_n_tiles == 1 ? env_int("FVY_CTA2_128", 1) != 0 : env_int("FVY_CTA2_128", 1) >= 2)) ||
                  (L.BN == 128 && L.taps == 1 && L.num_n_tiles == 1 && env_int("FVY_CTA2_128_1X1", 0) != 0));
        // stride-1 3x3: the three column taps of a filter row read one A slab at row shifts 0, 1, 2
        const bool slab1 = L.taps == 9 && s.stride == 1 && env_int("FVY_SLAB", 1) != 0;
        // stride-2 3x3: column taps 0 and 2 of a filter row are the same input phase one row apart -> one slab for both,
        // a second box for tap 1 (taps are stored in the order s = 0, 2, 1 for these layers)
        const bool slab2 = !stem && L.taps == 9 && s.stride == 2 && env_int("FVY_SLAB2", 1) != 0;
        L.tap_perm = slab2;
        const bool slab = slab1 || slab2;
        const int srows = L.BK == 64 ? slab_rows<64>() : slab_rows<32>();
        const size_t a_tile = (size_t)kBlockM * L.BK * 2, b_tile = (size_t)(L.cta2 ? L.BN / 2 : L.BN) * L.BK * 2;
        const int a_cover = slab ? gt : (L.BK == 32 ? gt : 1);
        const size_t a_slot = slab ? (size_t)(slab2 ? 2 : 1) * srows * L.BK * 2 : a_cover * a_tile;
        int b_cover = (gt == 3 && a_cover == gt && gt * b_tile <= (size_t)env_int("FVY_B3_MAX", 24576)) ? gt : 1;   // a filter row of B tiles per slot (one 3-D TMA box)
        // Layers with a short K loop are epilogue-bound: two epilogue groups alternate tiles.  Deep-K layers keep one
        // group so that the shared memory goes to the operand pipeline instead of a second staging ring.
        const int k_chunks = L.cin_pad / L.BK;
        const int k_iters = (L.taps / (k_chunks == 1 ? gt : 1)) * k_chunks;
        const int groups = env_int("FVY_GROUPS", 0) > 0 ? env_int("FVY_GROUPS", 0) : 2;
        L.deep_k = k_iters * (L.BN / 32) > groups_kn;        // MMA-bound tiles: the epilogue has slack, its latency is what shows
        int nb = has_res ? nb_res : nb_plain, lead = 0;
        nb = std::max(2, std::min(nb, kMaxRing));
        size_t fixed = 1024 + kSmemRing + (size_t)groups * nb * kChunkBytes;
        size_t budget = 232448 - fixed;
        // Resident weights: with a single N tile per CTA the whole [BN, K] weight tile is loaded once and every later
        // tile of the persistent CTA only streams A (half the operand bytes of a 1x1 layer, a quarter of a slab 3x3 layer).
        const size_t b_total = (size_t)L.taps * k_chunks * b_tile;
        int a_stages = 0, b_stages = 0, b_res = 0;
        // conv_5 (stride 2, 64 -> 128): its 74 KB weight half-tile and three 35 KB A slots miss the budget by 6 KB with three
        // staging buffers per group; streaming the weights instead re-reads 72 KB per tile from L2 next to 104 KB of A
        // (~40 B/clk/SM, the MMA issuer starved 67 % of the time) - two staging buffers buy the residency.
        if (slab2 && L.num_n_tiles == 1 && env_int("FVY_RESIDENT", 1) != 0 && b_total + 3 * a_slot > budget && nb > 2 && !has_res &&
            b_total + 3 * a_slot <= budget + (size_t)groups * (nb - 2) * kChunkBytes && env_int("FVY_RESIDENT_NB2", 1) != 0) {
            nb = 2;
            fixed = 1024 + kSmemRing + (size_t)groups * nb * kChunkBytes;
            budget = 232448 - fixed;
        }
        if (L.num_n_tiles == 1 && env_int("FVY_RESIDENT", 1) != 0 && b_total + 3 * a_slot <= budget) {
            if (L.taps * k_chunks / b_cover > kMaxB && a_cover == gt) b_cover = gt;
            if (L.taps * k_chunks / b_cover <= kMaxB) {
                b_res = 1;
                b_stages = L.taps * k_chunks / b_cover;
                a_stages = (int)std::min<size_t>(std::min(kMaxA, stages_cap), (budget - b_total) / a_slot);
            }
        }
        if (!b_res) {
            // streaming: maximise the taps in flight of the scarcer operand, then the bytes in flight
            const size_t b_slot = b_cover * b_tile;
            // Stride-2 layers: an A slot is two slabs (35 KB at BK = 64) and arrives from the 4-phase planes with the latency of a
            // cold read, while the weights are shared by every CTA and hit L2 - with two A slots the next one can only be requested
            // when the current one is consumed and the MMA issuer starves (FVY_DBG: conv_12 waits for operands 74 % of the time),
            // so these layers take at least three A slots when that leaves two B slots.
            long best = -1;
            for (int min_a = slab2 ? env_int("FVY_S2_MINA", 3) : 2; best < 0 && min_a >= 2; --min_a)
                for (int a = min_a; a <= std::min(kMaxA, stages_cap); ++a)
                    for (int b = 2; b <= std::min(kMaxB, stages_cap * gt); ++b) {
                        const size_t bytes = a * a_slot + b * b_slot;
                        if (bytes > budget) break;
                        const long score = (long)std::min(a * a_cover, b * b_cover) * 1000000 + (long)(bytes >> 10);
                        if (score > best) { best = score; a_stages = a; b_stages = b; }
                    }
            if (best < 0) return fail(FVY_E_INVALID, "conv_%d: no operand pipeline fits in shared memory", s.idx);
        }
        L.stages = a_stages; L.b_stages = b_stages; L.b_resident = b_res;
        L.smem_bytes = fixed + (size_t)a_stages * a_slot + (size_t)b_stages * b_cover * b_tile;
        if (L.smem_bytes > 232448) return fail(FVY_E_INVALID, "conv_%d: shared memory plan %zu exceeds 227 KB", s.idx, L.smem_bytes);
        if (int e = query_occ(L.BN, L.BK, L.cta2, L.smem_bytes, &L.occ)) return e;
        L.occ = 1;   // 320 threads x ~140 registers: one CTA per SM; latency is hidden inside the CTA (stages, two epilogue groups)
        // operands
        const size_t kdim = (size_t)L.taps * L.cin_pad;
        if (int e = dev_alloc(h, (void**)&L.w, (size_t)L.cout_pad * kdim * 2, true)) return e;
        if (int e = dev_alloc(h, (void**)&L.bias, (size_t)L.cout_pad * 4, true)) return e;
        if (b_cover == 3) { if (int e = make_tmap_b3(&L.tmap_b, L.w, L.cin_pad, L.cout_pad, L.taps, L.BK, L.cta2 ? L.BN / 2 : L.BN, 3)) return e; }
        else if (int e = make_tmap_2d(&L.tmap_b, L.w, kdim, L.cout_pad, kdim, L.BK, L.cta2 ? L.BN / 2 : L.BN)) return e;
        ConvParams& p = L.p;
        memset(&p, 0, sizeof(p));
        p.num_taps = L.taps;
        p.k_chunks = L.cin_pad / L.BK;
        p.a_choff = 0;
        p.H = L.Hout; p.W = L.Wout;
        p.leaky = s.leaky ? 1 : 0;
        p.bias = L.bias;
        p.num_n_tiles = L.num_n_tiles;
        p.nb = nb; p.lead = lead; p.epi_groups = groups;
        p.gt = gt; p.a_slab = slab2 ? 2 : (slab1 ? 1 : 0); p.a_cover = a_cover; p.a_stages = a_stages;
        p.b_cover = b_cover; p.b_stages = b_stages; p.b_resident = b_res;
        const void* a_base = nullptr;
        uint64_t a_rows = 0, a_pitch = 0;
        if (stem) {
            if (h->fused_stem) { a_base = L.w; a_rows = (uint64_t)L.cout_pad; a_pitch = 32; }   // placeholder map: conv_0 runs in stem_conv_kernel
            else { a_base = h->d_stem; a_rows = (uint64_t)nmax * L.Hin * L.Win; a_pitch = 32; }
            p.dom_plane = L.Hin * L.Win; p.dom_w = L.Win; p.dom_off = 0; p.tap_off[0] = 0;
        } else if (s.stride == 2) {
            const int Ho = L.Hout, Wo = L.Wout;
            // The compute domain is the padded geometry of the OUTPUT ((Ho+2) x (Wo+2), like a stride-1 layer), and the four
            // phase planes of the input are stored with that same geometry: output pixel (h, w) = domain row m reads phase
            // (r&1, s&1) at position (h + (r>>1), w + (s>>1)) = row m + ((r>>1) - 1) * (Wo+2) + ((s>>1) - 1) of that plane:
            // every tap is a constant row shift AND the rows of the domain are the rows of the padded output (TMA stores).
            const long long plane = (long long)(Ho + 2) * (Wo + 2);
            a_base = bufs[s.src].phase; a_rows = (uint64_t)(4 * nmax * plane); a_pitch = s.cin;
            p.dom_plane = (int)plane; p.dom_w = Wo + 2; p.dom_off = 1;
            for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 3; ++q)
                    p.tap_off[r * 3 + (L.tap_perm ? (q == 0 ? 0 : (q == 2 ? 1 : 2)) : q)] =
                        (int)((((r & 1) << 1) | (q & 1)) * nmax * plane + ((r >> 1) - 1) * (Wo + 2) + ((q >> 1) - 1));
        } else {
            const int H = L.Hin, W = L.Win;
            if (s.src == -2) { a_base = catA; a_pitch = 768; }
            else if (s.src == -3) { a_base = catB; a_pitch = 384; }
            else { a_base = bufs[s.src].padded; a_pitch = s.cin; }
            a_rows = (uint64_t)nmax * (H + 2) * (W + 2);
            p.dom_plane = (H + 2) * (W + 2); p.dom_w = W + 2; p.dom_off = 1;
            if (s.k == 1) p.tap_off[0] = 0;
            else
                for (int r = 0; r < 3; ++r)
                    for (int q = 0; q < 3; ++q) p.tap_off[r * 3 + q] = (r - 1) * (W + 2) + (q - 1);
        }
        if (a_base == nullptr) return fail(FVY_E_INVALID, "conv_%d: input buffer missing", s.idx);
        p.magic_plane = ~0ull / (unsigned long long)p.dom_plane + 1ull;
        p.magic_w = ~0ull / (unsigned long long)p.dom_w + 1ull;
        if ((long long)nmax * p.dom_plane >= (1ll << 31)) return fail(FVY_E_INVALID, "conv_%d: %d x %d rows overflow int32", s.idx, nmax, p.dom_plane);
        if (int e = make_tmap_2d(&L.tmap_a, a_base, a_pitch, a_rows, a_pitch, L.BK, slab ? srows : kBlockM)) return e;
        L.tmap_res = L.tmap_a; L.tmap_out[0] = L.tmap_a; L.tmap_out[1] = L.tmap_a;   // placeholders for unused maps
        // rows of the compute domain coincide with rows of the padded (H, W) output buffer (stride-1 convs on a padded input,
        // stride-2 convs on phase planes of the output's geometry)
        const bool coincident = !stem;
        const uint64_t out_rows = (uint64_t)nmax * (L.Hout + 2) * (L.Wout + 2);
        if (s.res >= 0) {
            p.res = bufs[s.res].padded; p.res_pitch = by_idx[s.res]->cout; p.res_choff = 0;
            if (!p.res) return fail(FVY_E_INVALID, "conv_%d: residual buffer missing", s.idx);
            if (!coincident) return fail(FVY_E_INVALID, "conv_%d: residual on a non stride-1 layer", s.idx);
            if (int e = make_tmap_2d(&L.tmap_res, p.res, p.res_pitch, out_rows, p.res_pitch, 32, kBlockM)) return e;
        }
        // outputs
        int no = 0;
        bool tmap_fail = false;
        const bool use_tma_store = env_int("FVY_TMA_STORE", 1) != 0;
        auto add_out = [&](void* ptr, int kind, int pitch, int choff, int c_real) {
            OutDesc od; od.ptr = ptr; od.aux = nullptr; od.kind = kind; od.pitch = pitch; od.choff = choff; od.nmax = nmax; od.c_real = c_real;
            od.tma = (kind == OUT_PADDED && coincident && use_tma_store) ? 1 : 0;
            if (no < 2 && od.tma && make_tmap_2d(&L.tmap_out[no], ptr, (uint64_t)pitch, out_rows, (uint64_t)pitch, 32, kBlockM)) tmap_fail = true;
            // 4-phase form of a stride-1 layer's output: TMA stores through the 5-D phase view (needs an even padded width, which
            // every level with a stride-2 consumer has)
            if (no < 2 && kind == OUT_PHASE && coincident && use_tma_store && env_int("FVY_TMA_PHASE", 1) != 0 && s.stride == 1 && (L.Wout & 1) == 0 &&
                (L.Hout & 1) == 0 && L.Wout >= env_int("FVY_TMA_PHASE_MIN_W", 32)) {   // narrower rows: too many stores per tile (conv_60 @26: 57 vs 56 us)
                CUtensorMap maps[7];
                bool ok = true;
                for (int k = 0; k < 7 && ok; ++k) ok = make_tmap_phase(&maps[k], ptr, nmax, L.Hout, L.Wout, pitch, 1 << k) == FVY_OK;
                void* dmaps = nullptr;
                if (!ok || dev_alloc(h, &dmaps, sizeof(maps), false) || cudaMemcpy(dmaps, maps, sizeof(maps), cudaMemcpyHostToDevice) != cudaSuccess) tmap_fail = true;
                else { od.tma = 2; od.aux = dmaps; L.tmap_out[no] = maps[6]; }
            }
            if (no < 2) p.out[no] = od;
            ++no;
        };
        if (!s.bn) {
            L.head_slot = head_i;
            add_out(h->d_logits[head_i++], OUT_HEAD_F32, s.cout, 0, s.cout);
        } else {
            if (bufs[s.idx].padded) add_out(bufs[s.idx].padded, OUT_PADDED, s.cout, 0, s.cout);
            if (bufs[s.idx].phase) add_out(bufs[s.idx].phase, OUT_PHASE, s.cout, 0, s.cout);
            if (c.head == FVY_HEAD_YOLO3) {
                if (s.idx == 60) add_out(catA, OUT_PADDED, 768, 256, s.cout);
                if (s.idx == 35) add_out(catB, OUT_PADDED, 384, 128, s.cout);
                if (s.idx == 84) add_out(catA, OUT_UP2_PADDED, 768, 0, s.cout);
                if (s.idx == 96) add_out(catB, OUT_UP2_PADDED, 384, 0, s.cout);
            }
        }
        if (tmap_fail) return FVY_E_CUDA;
        if (no == 0) return fail(FVY_E_INVALID, "conv_%d has no consumer", s.idx);
        if (no > 2) return fail(FVY_E_INVALID, "conv_%d has more than two stored forms", s.idx);
        L.primary = p.out[0];
        h->layers.push_back(L);
    }
    h->weight_count = (long long)stream_off;
    // ---- cross-layer tile dependencies: consumer = stride-1 conv reading the plain padded output of a producer whose stored
    // forms all leave by TMA (same geometry: the consumer's compute-domain rows ARE the producer's output rows)
    {
        h->use_flags = env_int("FVY_FLAGS", 1) != 0;
        std::map<int, int> layer_of;
        for (size_t i = 0; i < h->layers.size(); ++i) layer_of[h->layers[i].s.idx] = (int)i;
        size_t total = 0;
        for (size_t i = 0; i < h->layers.size() && h->use_flags; ++i) {
            Layer& C = h->layers[i];
            if (C.s.src < 0 || C.s.stride != 1 || !layer_of.count(C.s.src)) continue;
            Layer& P = h->layers[layer_of[C.s.src]];
            if (P.s.level != C.s.level) continue;
            bool pure_tma = true;
            for (int o = 0; o < 2; ++o)
                if (P.p.out[o].kind != OUT_NONE && !P.p.out[o].tma) pure_tma = false;
            if (!pure_tma || P.p.out[0].kind != OUT_PADDED) continue;
            C.wait_on = layer_of[C.s.src];
            P.signals = true;
        }
        for (Layer& L : h->layers)
            if (L.signals) { L.flag_blocks = (int)(((size_t)nmax * L.p.dom_plane + 127) / 128) + 2; total += (size_t)L.flag_blocks; }
        if (total) {
            h->flags_bytes = total * sizeof(int);
            if (int e = dev_alloc(h, (void**)&h->d_flags, h->flags_bytes, true)) return e;
            size_t off = 0;
            for (Layer& L : h->layers)
                if (L.signals) { L.flags = h->d_flags + off; off += (size_t)L.flag_blocks; }
        }
    }
    // ---- chains: runs of consecutive layers of the 256-wide CTA-pair instance, each reading its predecessor's TMA-stored
    // output, go out as ONE persistent launch (conv_chain_kernel)
    {
        h->use_chain = h->use_flags && env_int("FVY_CHAIN", 1) != 0;
        h->chain_sched = env_int("FVY_CHAIN_SCHED", 0) != 0;
        auto eligible = [&](const Layer& L) {
            if (!L.cta2 || L.BN != 256 || L.BK != 64 || L.s.stride != 1 || L.s.src < 0 || !L.s.bn) return false;
            if (L.p.b_resident || L.p.b_cover != 1) return false;
            if (L.taps == 9 ? L.p.a_slab != 1 : (L.taps != 1 || L.p.a_slab != 0 || L.p.a_cover != 1)) return false;
            for (int o = 0; o < 2; ++o)
                if (L.p.out[o].kind != OUT_NONE && !(L.p.out[o].kind == OUT_PADDED && L.p.out[o].tma)) return false;
            return L.p.out[0].kind == OUT_PADDED;
        };
        const int n = (int)h->layers.size();
        for (int i = 0; i < n && h->use_chain;) {
            if (!eligible(h->layers[i])) { ++i; continue; }
            int j = i + 1;
            const int max_len = env_int("FVY_CHAIN_MAXLEN", 64);
            while (j < n && j - i < max_len && eligible(h->layers[j]) && h->layers[j].wait_on == j - 1 && h->layers[j - 1].signals) ++j;
            if (j - i >= 2) {
                fvy_handle::Chain ch;
                ch.first = i; ch.count = j - i;
                ch.host.resize(ch.count);
                if (int e = dev_alloc(h, (void**)&ch.dev, sizeof(ChainLayer) * ch.count, true)) return e;
                for (int k = i; k < j; ++k) { h->layers[k].chain = (int)h->chains.size(); h->layers[k].chain_pos = k - i; }
                h->chains.push_back(std::move(ch));
            }
            i = j;
        }
        if (!h->chains.empty()) {
            // one shared-memory carve-up for every layer of a chain: slab-sized A slots, per-tap B slots, nb staging buffers per group
            h->chain_nb = env_int("FVY_CHAIN_NB", 3); h->chain_a = env_int("FVY_CHAIN_A", 4); h->chain_b = env_int("FVY_CHAIN_B", 6);
            h->chain_smem = 1024 + kSmemRing + (size_t)2 * h->chain_nb * kChunkBytes + (size_t)h->chain_a * slab_rows<64>() * 64 * 2 +
                            (size_t)h->chain_b * 128 * 64 * 2;
            if (h->chain_smem > 232448) return fail(FVY_E_INVALID, "chain shared-memory plan %zu exceeds 227 KB", h->chain_smem);
            CUDA_TRY(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
            CUDA_TRY(cudaFuncSetAttribute(conv_chain_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        }
    }
    return FVY_OK;
}

// FVY_CHAIN_SCHED=1: a list schedule of the chain's tiles instead of the static rotation.  Every tile is one item of known length
// (taps x K chunks, all of them 256 x 256 x 64 MMAs) that becomes ready when the row blocks of the previous layer its taps reach
// (and the residual rows) are complete; layers are taken in order and, inside a layer, tiles in ascending order, each going to the
// pair that can start it first.  Per pair the list is layer-monotonic, so a pair only ever waits for items that precede its own
// in the other pairs' lists: no cycle.  Times are in units of one tap; `load` and `drain` model the first-operand latency and the
// epilogue + store + counter visibility of a finished tile.
static int schedule_chain(fvy_handle* h, fvy_handle::Chain& ch, int pairs) {
    static const int load = [] { const char* v = getenv("FVY_SCHED_LOAD"); return v && *v ? atoi(v) : 5; }();
    static const int drain = [] { const char* v = getenv("FVY_SCHED_DRAIN"); return v && *v ? atoi(v) : 10; }();
    std::vector<std::vector<int>> lists(pairs);
    std::vector<long long> free_at(pairs, 0);
    std::vector<std::vector<long long>> done(ch.count);
    size_t total = 0;
    for (int k = 0; k < ch.count; ++k) {
        const ConvParams& p = ch.host[k].p;
        const Layer& L = h->layers[ch.first + k];
        const int nnt = p.num_n_tiles, mt_count = (p.num_m_tiles + 1) / 2, tiles = mt_count * nnt;
        const long long dur = (long long)p.num_taps * p.k_chunks;
        if (tiles >= (1 << 20) || k >= (1 << 10)) return fail(FVY_E_INVALID, "chain schedule: %d tiles / layer %d do not fit the entry format", tiles, k);
        done[k].assign(mt_count, 0);
        int res_k = -1;
        if (L.s.res >= 0)
            for (int q = 0; q < k; ++q)
                if (h->layers[ch.first + q].s.idx == L.s.res) res_k = q;
        for (int t = 0; t < tiles; ++t) {
            const int mt = t / nnt;
            long long ready = 0;
            if (k > 0) {
                const long long m0 = (long long)mt * 2 * kBlockM, margin = p.wait_margin;
                const int lo = (int)std::max<long long>(0, (m0 - margin) / (2 * kBlockM));
                const int hi = (int)std::min<long long>((long long)done[k - 1].size() - 1, (m0 + 2 * kBlockM - 1 + margin) / (2 * kBlockM));
                for (int b = lo; b <= hi; ++b) ready = std::max(ready, done[k - 1][b]);
            }
            if (res_k >= 0 && mt < (int)done[res_k].size()) ready = std::max(ready, done[res_k][mt]);
            ready += load;
            int best = 0;
            long long best_start = -1, best_free = -1;
            for (int q = 0; q < pairs; ++q) {
                const long long st = std::max(free_at[q], ready);
                if (best_start < 0 || st < best_start || (st == best_start && free_at[q] > best_free)) { best = q; best_start = st; best_free = free_at[q]; }
            }
            lists[best].push_back((k << 20) | t);
            free_at[best] = best_start + dur;
            done[k][mt] = std::max(done[k][mt], best_start + dur + drain);
            ++total;
        }
    }
    size_t longest = 0;
    for (const auto& l : lists) longest = std::max(longest, l.size());
    const int stride = (int)longest + 1;
    if (ch.d_sched == nullptr || stride > ch.sched_stride) {
        const int alloc_stride = stride + stride / 4 + 8;
        void* pmem = nullptr;
        if (int e = dev_alloc(h, &pmem, (size_t)pairs * alloc_stride * sizeof(int), false)) return e;
        ch.d_sched = (int*)pmem; ch.sched_stride = alloc_stride;
    }
    ch.sched_host.assign((size_t)pairs * ch.sched_stride, -1);
    for (int q = 0; q < pairs; ++q) std::copy(lists[q].begin(), lists[q].end(), ch.sched_host.begin() + (size_t)q * ch.sched_stride);
    CUDA_TRY(cudaMemcpyAsync(ch.d_sched, ch.sched_host.data(), ch.sched_host.size() * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    (void)total;
    return FVY_OK;
}

// Per-call chain descriptors (batch-dependent fields), uploaded in stream order before the forward that uses them.
static int prepare_chains(fvy_handle* h, int batch) {
    if (h->chains.empty() || h->chain_batch == batch) return FVY_OK;
    const int pairs = (h->num_sms & ~1) / 2;
    for (fvy_handle::Chain& ch : h->chains) {
        for (int k = 0; k < ch.count; ++k) {
            Layer& L = h->layers[ch.first + k];
            ChainLayer& c = ch.host[k];
            memset(&c, 0, sizeof(c));
            c.tmap_a = L.tmap_a; c.tmap_b = L.tmap_b; c.tmap_res = L.tmap_res; c.tmap_out0 = L.tmap_out[0]; c.tmap_out1 = L.tmap_out[1];
            c.p = L.p;
            c.p.m_total = batch * L.p.dom_plane;
            c.p.num_m_tiles = (c.p.m_total + kBlockM - 1) / kBlockM;
            c.p.epi_groups = 2; c.p.epi_split = 1; c.p.dbg = nullptr;
            c.p.split_from = ((c.p.num_m_tiles + 1) / 2) * c.p.num_n_tiles;
            c.p.sig_flags = L.signals ? L.flags : nullptr;
            c.p.wait_flags = nullptr;
            if (k > 0) {
                const Layer& P = h->layers[L.wait_on];
                c.p.wait_flags = P.flags;
                c.p.wait_expected = P.num_n_tiles * 2;              // both epilogue groups store part of every tile
                c.p.wait_margin = L.s.k == 3 ? L.Win + 3 : 0;
                c.p.wait_blocks = (batch * P.p.dom_plane + 127) / 128;
            }
            c.res_flags = nullptr;
            if (L.s.res >= 0)
                for (int q = 0; q < k; ++q) {
                    const Layer& R = h->layers[ch.first + q];
                    if (R.s.idx == L.s.res && R.signals) {       // the residual rows come from a layer of this chain
                        c.res_flags = R.flags; c.res_expected = R.num_n_tiles * 2; c.res_blocks = (batch * R.p.dom_plane + 127) / 128;
                    }
                }
            c.rot = (k * 25) % pairs;
        }
        CUDA_TRY(cudaMemcpyAsync(ch.dev, ch.host.data(), sizeof(ChainLayer) * ch.count, cudaMemcpyHostToDevice, h->stream));
        if (h->chain_sched)
            if (int e = schedule_chain(h, ch, pairs)) return e;
    }
    h->chain_batch = batch;
    return FVY_OK;
}

static int launch_chain(fvy_handle* h, const fvy_handle::Chain& ch) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(h->num_sms & ~1); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = h->chain_smem; cfg.stream = h->stream;
    cudaLaunchAttribute at[2];
    int na = 0;
    if (h->use_pdl) {
        at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    at[na].id = cudaLaunchAttributeClusterDimension;
    at[na].val.clusterDim.x = 2; at[na].val.clusterDim.y = 1; at[na].val.clusterDim.z = 1;
    ++na;
    cfg.attrs = at; cfg.numAttrs = na;
    CUDA_TRY(cudaLaunchKernelEx(&cfg, conv_chain_kernel, (const ChainLayer*)ch.dev, ch.count, h->chain_nb, h->chain_a, h->chain_b,
                                (const int*)(h->chain_sched ? ch.d_sched : nullptr), ch.sched_stride));
    h->launches += 1;
    return FVY_OK;
}

static int total_cands(const fvy_handle* h) {
    if (h->cfg.head == FVY_HEAD_FD6) return (h->cfg.net_h / 32) * (h->cfg.net_w / 32);
    int t = 0;
    for (int lvl : {32, 16, 8}) t += 3 * (h->cfg.net_h / lvl) * (h->cfg.net_w / lvl);
    return t;
}

static int build_post(fvy_handle* h) {
    const fvy_config& c = h->cfg;
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
    if (int e = dev_alloc(h, (void**)&h->d_status, 16, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_image_hw, (size_t)B * 8, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_order, (size_t)B * h->capP * 4, true)) return e;
    if (int e = dev_alloc(h, (void**)&h->d_sbox, (size_t)B * h->capP * 16, true)) return e;
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
    return FVY_OK;
}

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
        if (L.s.src == -1 && h->fused_stem) {
            if (!h->cur_img) return fail(FVY_E_STATE, "no input image resident for conv_0");
            if (h->stem_mode == 3) {
                const long long total = (long long)batch * (h->cfg.net_w / 16) * h->cfg.net_h;     // (image, strip, row) triples
                const int blocks = (int)std::min<long long>((total + kStripWarps - 1) / kStripWarps, (long long)h->num_sms * h->stem_blocks_per_sm);
                if (h->cur_dtype == FVY_F32)
                    stem_strip_kernel<float><<<blocks, kStripWarps * 32, 0, h->stream>>>((const float*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                         h->d_stem_w2, L.bias, (__nv_bfloat16*)L.p.out[0].ptr);
                else if (h->cur_dtype == FVY_U8)
                    stem_strip_kernel<unsigned char><<<blocks, kStripWarps * 32, 0, h->stream>>>((const unsigned char*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w,
                                                                                                 h->cfg.max_batch, h->d_stem_w2, L.bias, (__nv_bfloat16*)L.p.out[0].ptr);
                else
                    stem_strip_kernel<double><<<blocks, kStripWarps * 32, 0, h->stream>>>((const double*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                          h->d_stem_w2, L.bias, (__nv_bfloat16*)L.p.out[0].ptr);
            } else if (h->stem_mode == 2) {
                const long long total_rows = (long long)batch * h->cfg.net_h;
                const int blocks = (int)std::min<long long>(total_rows, (long long)h->num_sms * 2);
                const int rows_per_block = (int)(((total_rows + blocks - 1) / blocks + 1) & ~1LL);     // even: two rows per iteration
                const int rowlen = (((h->cfg.net_w + 2) * 3 + 2) + 7) & ~7;
                const size_t smem = (size_t)8 * 2 * rowlen * 2;
                if (h->cur_dtype == FVY_F32)
                    stem_rows_kernel<float><<<blocks, kStemThreads, smem, h->stream>>>((const float*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                      h->d_stem_w2, L.bias, (__nv_bfloat16*)L.p.out[0].ptr, rows_per_block, rowlen);
                else
                    stem_rows_kernel<double><<<blocks, kStemThreads, smem, h->stream>>>((const double*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                                       h->d_stem_w2, L.bias, (__nv_bfloat16*)L.p.out[0].ptr, rows_per_block, rowlen);
            } else {
            const int blocks = h->num_sms * 8;   // persistent warps, 2 waves of 4 resident blocks per SM
            if (h->cur_dtype == FVY_F32)
                stem_conv_kernel<float><<<blocks, 256, 0, h->stream>>>((const float*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                       L.w, L.bias, (__nv_bfloat16*)L.p.out[0].ptr);
            else
                stem_conv_kernel<double><<<blocks, 256, 0, h->stream>>>((const double*)h->cur_img, batch, h->cfg.net_h, h->cfg.net_w, h->cfg.max_batch,
                                                                        L.w, L.bias, (__nv_bfloat16*)L.p.out[0].ptr);
            }
            CUDA_TRY(cudaGetLastError());
            ++h->launches;
            if (h->last_slot >= 0 && !h->capturing) { CUDA_TRY(cudaEventRecord(h->ev_consumed[h->last_slot], h->stream)); h->last_slot = -1; }
            continue;
        }
        if (!L.s.bn && L.head_slot >= 0)     // head logits of this call's set
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
                L.p.wait_margin = L.s.k == 3 ? L.Win + 3 : 0;
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
    if (!h->weights_loaded) return fail(FVY_E_STATE, "fvy_forward before fvy_load_weights");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d outside [1, %d]", batch, h->cfg.max_batch);
    if (dtype != FVY_F32 && dtype != FVY_F64 && dtype != FVY_U8) return fail(FVY_E_INVALID, "dtype %d", dtype);
    if (dtype == FVY_U8 && !(h->fused_stem && h->stem_mode == 3))
        return fail(FVY_E_INVALID, "uint8 images need the strip stem (FVY_STEM=3, network width a multiple of 16)");
    const void* dimg = nullptr;
    if (int e = stage_input(h, images, dtype, batch, &dimg)) return e;
    h->cur_img = dimg; h->cur_dtype = dtype;
    if (!h->fused_stem) {
        const long long pix = (long long)batch * h->cfg.net_h * h->cfg.net_w;
        const int blocks = (int)std::min<long long>((pix + 255) / 256, (long long)h->num_sms * 16);
        if (dtype == FVY_F32)
            stem_im2col_kernel<float><<<blocks, 256, 0, h->stream>>>((const float*)dimg, batch, h->cfg.net_h, h->cfg.net_w, h->d_stem);
        else
            stem_im2col_kernel<double><<<blocks, 256, 0, h->stream>>>((const double*)dimg, batch, h->cfg.net_h, h->cfg.net_w, h->d_stem);
        CUDA_TRY(cudaGetLastError());
        ++h->launches;
        if (h->last_slot >= 0) { CUDA_TRY(cudaEventRecord(h->ev_consumed[h->last_slot], h->stream)); h->last_slot = -1; }
    }
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
    if (!h->use_graph || !h->fused_stem) return run_all();
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
    CUDA_TRY(cudaMemsetAsync(h->d_status, 0, 4, h->ps));
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
        decode_yolo_kernel<<<batch, 1024, 0, h->ps>>>(a);
    }
    CUDA_TRY(cudaGetLastError());
    ++h->launches;
    return FVY_OK;
}

// NMS over device-resident segments (ibox/classes with stride `seg_stride`, counts on device)
static int nms_enqueue(fvy_handle* h, const int* d_ibox, float* d_cls, const int* d_counts, int batch, int seg_stride, int nb_class,
                       double th) {
    for (int c = 0; c < nb_class; ++c) {
        SortArgs s;
        s.ibox = d_ibox; s.classes = d_cls; s.counts = d_counts; s.seg_stride = seg_stride; s.nb_class = nb_class; s.cls = c;
        s.capP = h->capP; s.descending = 1; s.order = h->d_order; s.sbox = h->d_sbox; s.gkeys = h->d_gkeys;
        s.smem_keys = h->smem_keys; s.np2max = h->np2max;
        sort_scores_kernel<<<batch, 1024, (size_t)h->smem_keys * 8, h->ps>>>(s);
        CUDA_TRY(cudaGetLastError());
        MaskArgs m;
        m.sbox = h->d_sbox; m.counts = d_counts; m.seg_stride = seg_stride; m.batch = batch; m.capP = h->capP; m.words = h->words;
        m.th = th; m.mask = h->d_mask; m.rowflag = h->d_rowflag;
        CUDA_TRY(cudaMemsetAsync(h->d_rowflag, 0, (size_t)batch * h->words * 8, h->ps));
        nms_mask_kernel<<<h->num_sms * 16, 64, 0, h->ps>>>(m);
        CUDA_TRY(cudaGetLastError());
        SweepArgs w;
        w.mask = h->d_mask; w.order = h->d_order; w.counts = d_counts; w.seg_stride = seg_stride; w.capP = h->capP; w.words = h->words;
        w.nb_class = nb_class; w.cls = c; w.classes = d_cls; w.rowflag = h->d_rowflag;
        static const int sweep_threads = [] { const char* v = getenv("FVY_SWEEP_THREADS"); const int t = v && *v ? atoi(v) : 1024; return t >= 128 && t <= 1024 && t % 32 == 0 ? t : 1024; }();
        nms_sweep_kernel<<<batch, sweep_threads, (size_t)h->words * 8, h->ps>>>(w);
        CUDA_TRY(cudaGetLastError());
        h->launches += 3;
    }
    return FVY_OK;
}

static int post_enqueue(fvy_handle* h, const float* dev[3], int batch, const fvy_post_params* pp, const int* d_hw, int max_out) {
    if (int e = decode_enqueue(h, dev, batch, pp, d_hw, false)) return e;
    const int nc = h->cfg.head == FVY_HEAD_FD6 ? 1 : h->cfg.nb_class;
    if (int e = nms_enqueue(h, h->d_ibox, h->d_cls, h->d_counts, batch, h->cap, nc, pp->nms_thresh)) return e;
    AssembleArgs a;
    a.ibox = h->d_ibox; a.objness = h->d_obj; a.classes = h->d_cls; a.cand = h->d_cand; a.counts = h->d_counts;
    a.seg_stride = h->cap; a.nb_class = nc; a.max_out = max_out;
    a.limit = pp->num_cands > 0 ? std::min(pp->num_cands, max_out) : max_out;
    a.kept_idx = nullptr; a.kept_counts = nullptr; a.dets = h->d_dets; a.det_counts = h->d_det_counts;
    if (h->cfg.head == FVY_HEAD_FD6) {
        SortArgs s;
        s.ibox = h->d_ibox; s.classes = h->d_cls; s.counts = h->d_counts; s.seg_stride = h->cap; s.nb_class = 1; s.cls = 0;
        s.capP = h->capP; s.descending = 0; s.order = h->d_order; s.sbox = nullptr; s.gkeys = h->d_gkeys;
        s.smem_keys = h->smem_keys; s.np2max = h->np2max;
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

}  // namespace fvy

// ============================================================================================ C ABI
extern "C" {

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
            h->overlap_post = !(v && *v && atoi(v) == 0);
            const char* g = getenv("FVY_GRAPH");
            h->use_graph = !(g && *g && atoi(g) == 0);
        }
        for (cudaEvent_t* ev : {&h->ev_fwd_done[0], &h->ev_fwd_done[1], &h->ev_post_done[0], &h->ev_post_done[1]})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
        for (auto& ev : h->ev) ok = ok && cudaEventCreate(&ev) == cudaSuccess;
        for (cudaEvent_t* ev : {&h->ev_ready[0], &h->ev_ready[1], &h->ev_consumed[0], &h->ev_consumed[1], &h->ev_post, &h->ev_d2h})
            ok = ok && cudaEventCreateWithFlags(ev, cudaEventDisableTiming) == cudaSuccess;
        if (!ok) { e = fail(FVY_E_CUDA, "cudaEventCreate failed"); break; }
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
        if (stem) {      // the same folded weights in stem_rows_kernel's K order: k = 10 r + (3 q + ci)
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
    if (!h || !images) return fail(FVY_E_INVALID, "NULL argument");
    if (h->cfg.head == FVY_HEAD_NONE) return fail(FVY_E_STATE, "handle has no network");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
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
    if (int e = nms_enqueue(h, d_ibox, d_cls, d_counts, batch, seg_stride, nb_class, nms_thresh)) return e;
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
    if (int e = post_enqueue(h, dev, batch, pp, d_hw, max_out)) return e;
    CUDA_TRY(cudaEventRecord(h->ev[3], h->ps));
    CUDA_TRY(cudaEventRecord(h->ev_post, h->ps));
    CUDA_TRY(cudaStreamWaitEvent(h->d2h_stream, h->ev_post, 0));
    CUDA_TRY(cudaMemcpyAsync(dets, h->d_dets, (size_t)batch * max_out * sizeof(FvyDet), cudaMemcpyDefault, h->d2h_stream));
    CUDA_TRY(cudaMemcpyAsync(det_counts, h->d_det_counts, (size_t)batch * 4, cudaMemcpyDefault, h->d2h_stream));
    CUDA_TRY(cudaEventRecord(h->ev_d2h, h->d2h_stream));
    if (!sync) return FVY_OK;
    std::vector<int> hc(batch);
    CUDA_TRY(cudaMemcpyAsync(hc.data(), h->d_counts, (size_t)batch * 4, cudaMemcpyDeviceToHost, h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->stream));
    CUDA_TRY(cudaStreamSynchronize(h->d2h_stream));
    CUDA_TRY(cudaEventElapsedTime(&h->last_post_ms, h->ev[2], h->ev[3]));
    return check_post_status(h, batch, hc.data());
}

int fvy_postprocess(fvy_handle* h, const float* out0, const float* out1, const float* out2, int batch, const fvy_post_params* pp,
                    const int* image_hw, int max_out, fvy_det* dets, int32_t* det_counts) {
    if (!h) return fail(FVY_E_INVALID, "NULL handle");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    return postprocess_common(h, out0, out1, out2, batch, pp, image_hw, max_out, dets, det_counts, true);
}

static int detect_common(fvy_handle* h, const void* images, int dtype, int batch, const fvy_post_params* pp, const int* image_hw,
                         int max_out, fvy_det* dets, int32_t* det_counts, bool sync) {
    if (!h || !images) return fail(FVY_E_INVALID, "NULL argument");
    if (h->cfg.head == FVY_HEAD_NONE) return fail(FVY_E_STATE, "handle has no network");
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const bool overlap = !sync && h->overlap_post;
    if (overlap) {
        h->logit_set ^= 1;
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[h->logit_set], 0));   // the post-processing that read this set (two calls ago) is done
    } else {
        h->logit_set = 0;
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[0], 0));
        CUDA_TRY(cudaStreamWaitEvent(h->stream, h->ev_post_done[1], 0));              // nothing of an earlier asynchronous call is still in flight
    }
    CUDA_TRY(cudaEventRecord(h->ev[0], h->stream));
    if (int e = forward_enqueue(h, images, dtype, batch)) return e;
    CUDA_TRY(cudaEventRecord(h->ev[1], h->stream));
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
    }
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
    return FVY_OK;
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
    if (!h || !dst_host || layer < 0 || layer >= (int)h->layers.size()) return fail(FVY_E_INVALID, "bad argument");
    if (batch < 1 || batch > h->cfg.max_batch) return fail(FVY_E_INVALID, "batch %d", batch);
    CUDA_TRY(cudaSetDevice(h->cfg.device));
    const Layer& L = h->layers[layer];
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
    if (getenv("FVY_DBG") && !(h->layers[layer].s.src == -1 && h->fused_stem)) {
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
    CUDA_TRY(cudaEventRecord(h->ev[4], h->stream));
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

// Keras Adam over a flat bucket: 16 bytes of each of p, g, m, v per thread and iteration (HBM-bound: 28 B per parameter).
// Separately rounded operations (the library is built with -fmad=false) so that the update equals the torch-op restatement.
__global__ void __launch_bounds__(256) adam_step_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                                        float* __restrict__ v, long long n, float lr_t, float b1, float b2, float eps,
                                                        float gs) {
    const long long n4 = n >> 2;
    const float c1 = 1.0f - b1, c2 = 1.0f - b2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i], mm = reinterpret_cast<float4*>(m)[i], vv = reinterpret_cast<float4*>(v)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float* pa = &pp.x; float* ma = &mm.x; float* va = &vv.x; const float* ga = &gg.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float gk = ga[k] * gs;
            ma[k] = ma[k] * b1 + c1 * gk;
            va[k] = va[k] * b2 + c2 * (gk * gk);
            pa[k] = pa[k] - (lr_t * ma[k]) / (sqrtf(va[k]) + eps);
        }
        reinterpret_cast<float4*>(p)[i] = pp; reinterpret_cast<float4*>(m)[i] = mm; reinterpret_cast<float4*>(v)[i] = vv;
    }
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {          // tail
        const long long i = (n4 << 2) + threadIdx.x;
        const float gk = g[i] * gs;
        m[i] = m[i] * b1 + c1 * gk;
        v[i] = v[i] * b2 + c2 * (gk * gk);
        p[i] = p[i] - (lr_t * m[i]) / (sqrtf(v[i]) + eps);
    }
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

void* fvy_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}
void fvy_host_free(void* p) { if (p) cudaFreeHost(p); }

}  // extern "C"
