// conv_0 (the stem) straight from the image: stem_strip_kernel.  Part of libfvy.so (included by fvy_api.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvy {

// conv_0 (3x3, pad 1, stride 1, 3 -> 32, BN folded, LeakyReLU(0.1)) straight from the fp32 / fp64 / uint8 image to the bf16
// 4-phase activation that conv_1 (stride 2) reads.  yolov3_detect.py:221 (conv_0), :205 (ZeroPadding2D(1)), :212-213.
// K = 27 and N = 32: 1.7 % of the network's FLOPs but its largest activation (11 MB / image), i.e. HBM-bound.  A tcgen05 tile
// (128 x 32, K = 32) carries too little math per TMEM / mbarrier round trip (measured 284 us + 156 us for an im2col operand),
// so this layer runs on warp-level bf16 MMAs (m16n8k16, fp32 accumulate) and transposes the result inside each quad so that
// every lane stores 16 contiguous bytes.
__device__ __forceinline__ void mma_m16n8k16_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// The image rows an output row needs are staged ONCE in shared memory as bf16 (every input element is read from HBM once,
// coalesced, and converted once instead of nine times).  K order: filter row r occupies k = 10 r .. 10 r + 9 (the nine contiguous
// window elements (dx, c) of the NHWC row + one zero slot), so every MMA fragment register (k, k + 1) is two adjacent bf16 of a
// staged row; a second copy of each row shifted by one element makes the pair 4-byte aligned for odd pixels as well; image
// borders are zeros in the staged rows, so the loop has no edge tests.  Every WARP owns a strip of 16 output columns over a
// segment of rows and keeps its own ring of eight staged rows (18 pixels x 3 channels, two copies): no block-wide barrier, a
// warp only waits for its own loads (two rows ahead, in registers while the current row is computed).  The (image, strip, row)
// space is cut into one equal piece per warp.  Halo columns are re-read by the neighbouring strip (12 %, L1 / L2 hits).
// (Round 1 also carried a register-gather form and a block-staged form, 219 us / 195 us against 158 us: deleted; network widths
// are multiples of 32, so the 16-column strips always tile the row.)
__device__ __forceinline__ float stem_px(float v) { return v; }
__device__ __forceinline__ float stem_px(double v) { return (float)v; }                         // Keras casts its input to float32
__device__ __forceinline__ float stem_px(unsigned char v) { return (float)((double)v / 255.0); }   // image / 255 in float64, then that cast
// uint8 frames: the 256 possible values of float32(v / 255.0) come from a per-block table (a double-precision divide per pixel made
// the uint8 stem slower than the float32 one: 12.5k vs 13.0k images/s end to end)
template <typename T> struct StemLut { static constexpr bool on = false; };
template <> struct StemLut<unsigned char> { static constexpr bool on = true; };
constexpr int kStripWarps = 8;
static_assert(kStripWarps * 32 == 256, "the uint8 table is filled by one thread per entry");
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
    __shared__ float s_lut[StemLut<T>::on ? 256 : 1];
    if (StemLut<T>::on) s_lut[threadIdx.x] = stem_px((unsigned char)threadIdx.x);      // kStripWarps * 32 = 256 threads
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
        if (StemLut<T>::on) {
            a = (in && ok0) ? s_lut[(int)__ldg(src + c0)] : 0.f;
            b = (in && ok1) ? s_lut[(int)__ldg(src + c1)] : 0.f;
        } else {
            a = (in && ok0) ? stem_px(__ldg(src + c0)) : 0.f;
            b = (in && ok1) ? stem_px(__ldg(src + c1)) : 0.f;
        }
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

}  // namespace fvy
