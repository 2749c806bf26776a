// conv_0 + conv_1 in ONE kernel (sm_100a): the stem's output never goes to HBM.
//
// Unfused, stem_strip_kernel writes conv_0's 32-channel activation (11.3 MB per 416x416 image: the largest tensor of the network,
// 443 MB at batch 40) as a 4-phase buffer and conv_1 (3x3, stride 2, 32 -> 64; yolov3_detect.py:221-223) reads it straight back:
// 158 us + 124 us of HBM-bound time for 7 % of the FLOPs.  Here a persistent CTA produces the conv_0 rows an output row of conv_1
// needs in SHARED memory, in exactly the form conv_1's tcgen05 MMA reads its A operand in, and consumes them from there:
//
//   producer warps (14)  conv_0 with warp-level bf16 MMAs (m16n8k16), the same staged-row scheme, K order and arithmetic as
//                        stem_strip_kernel (bit-identical values): every warp owns one 16-pixel strip of the tile's column
//                        range and walks down the rows, keeping a private ring of staged image rows; results are written
//                        as bf16 into a ring of conv_0 ROW SLOTS: per row two arrays - even and odd padded columns - of 64-byte
//                        entries (32 channels) in the 64-byte-swizzle K-major layout (the column phases a stride-2 tap reads)
//   MMA warp (1 thread)  per output row tile (<= 128 pixels of one output row): 9 taps x 2 K-steps of tcgen05.mma M = 128, N = 64
//                        into a TMEM accumulator stage; tap (r, s) reads row slot 2y + r, array s & 1, starting s >> 1 entries in
//                        (descriptors swizzle by address: any start row inside a 1 024-byte aligned array is legal); conv_1's
//                        weights (36 KB) are resident in shared memory; same tap order as conv_igemm_kernel (r, then s = 0, 2, 1)
//   epilogue warps (4)   TMEM -> registers -> + bias (BatchNorm folded) -> LeakyReLU(0.1) -> bf16 -> 128 contiguous bytes per pixel
//                        straight into conv_1's padded NHWC output (what conv_2 and conv_3's residual read); same arithmetic as
//                        conv_igemm_kernel's epilogue
//
// Work items: (image, column tile, segment of output rows); a conv_0 row is used by up to two output rows, so walking down a
// segment computes every conv_0 row once (plus one extra row per segment).  HBM traffic: image read once (+ halo re-reads from
// L2) and conv_1's output written once.  Row slots, TMEM stages and all barrier phases carry across tiles and work items.
#pragma once

#include <type_traits>

#include "conv_igemm_sm100.cuh"
#include "stem_kernel.cuh"

namespace fvy {

constexpr int kFuseProducers = 14;                 // producer warps: one 16-pixel strip each (14 strips at 416)
constexpr int kFuseThreads = 32 * (4 + 1 + kFuseProducers);      // warps 0-3 epilogue, 4 MMA / TMEM / weights, 5.. producers
constexpr int kFuseSlots = 6;                      // conv_0 row slots
constexpr int kFuseAcc = 4;                        // accumulator stages of 64 TMEM columns
constexpr int kFuseMaxTileW = (16 * kFuseProducers - 1) / 2;     // 111: 14 strips cover 2 x 111 + 1 conv_0 pixels
constexpr int kFusePf = 4;                         // image rows in flight per strip (registers): ~4 row times cover the HBM latency
constexpr int kFuseArrayBytes = 9 * 1024;          // one column-phase array of a row slot: <= 129 entries x 64 B, 1 024-byte aligned
constexpr int kFuseSlotBytes = 2 * kFuseArrayBytes;
constexpr int kFuseW1Bytes = 9 * 64 * 64;          // conv_1 weights: nine [64 x 32] bf16 tiles
// shared memory: [0, 1024) barriers + TMEM pointer | bias1 (64 floats) | conv_0 fragments | W1 | row slots (+ 2 KB slack: the MMA
// reads 128 entries from an array that holds <= 129 starting at entry 0 or 1) | per-producer staged image rows
constexpr int kFuseOffBias = 1024;
constexpr int kFuseOffFrag = kFuseOffBias + 256;                           // uint4 swf[4][32] + float2 sbf[4][32] = 3 072 B
constexpr int kFuseOffW1 = 5 * 1024;
constexpr int kFuseOffSlots = kFuseOffW1 + kFuseW1Bytes;                   // 41 KB, 1 024-aligned
constexpr int kFuseOffStage = kFuseOffSlots + kFuseSlots * kFuseSlotBytes + 2048;
constexpr int kFuseStageBytes = 4 * 2 * kStripLen * 2;                     // per producer warp: 4 image rows x 2 copies x 64 bf16
constexpr int kFuseSmem = kFuseOffStage + kFuseProducers * kFuseStageBytes + 1024;   // + alignment slack
static_assert(kFuseOffSlots % 1024 == 0 && kFuseSmem <= 232448, "fused stem: shared-memory plan");
static_assert(kFusePf == 4, "the row loop is unrolled by hand over the register slots");

__device__ __forceinline__ void sts16(uint32_t addr, uint16_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory"); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ float2 lds64f(uint32_t addr) {
    float2 v;
    asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr));
    return v;
}
// Long waits of the consumer roles (a tile takes microseconds): sleep between polls so that the spinning warps do not take
// issue slots from the producer warps they share a scheduler with (in a first version 35 % of all issued instructions were polls).
__device__ __forceinline__ void mbar_wait_sleepy(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        if (mbar_try_u32(addr, parity)) return;
        __nanosleep(spin < 4 ? 64 : 256);
        if (spin > (1u << 22)) {
            printf("fvy: fused stem wait timed out (block %d thread %d bar %u parity %u)\n", blockIdx.x, threadIdx.x, addr, parity);
            __trap();
        }
    }
}

struct FuseParams {
    int batch, H, W;              // network input
    int tiles_w, tile_w;          // column tiles per output row and their nominal width (last one may be narrower)
    int seg_rows, segs;           // output rows per work item, work items per (image, column tile)
    const __nv_bfloat16* w0;      // conv_0 weights in stem_strip_kernel's K order [32][32]
    const float* bias0;           // conv_0 folded bias [32]
    const float* bias1;           // conv_1 folded bias [64]
    __nv_bfloat16* out;           // conv_1 output, padded NHWC [n][H/2 + 2][W/2 + 2][64]
};

template <typename T>
__global__ void __launch_bounds__(kFuseThreads, 1)
stem_conv1_fused_kernel(const __grid_constant__ CUtensorMap tmap_w1, const T* __restrict__ img, const FuseParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* row_full = reinterpret_cast<uint64_t*>(smem);          // [kFuseSlots]  conv_0 row complete (count: producer warps)
    uint64_t* row_empty = row_full + kFuseSlots;                       // [kFuseSlots]  every MMA that reads the row has retired
    uint64_t* tmem_full = row_empty + kFuseSlots;                      // [kFuseAcc]
    uint64_t* tmem_empty = tmem_full + kFuseAcc;                       // [kFuseAcc]    count: 4 epilogue warps
    uint64_t* w_full = tmem_empty + kFuseAcc;                          // conv_1 weights landed
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(w_full + 1);
    float* sbias1 = reinterpret_cast<float*>(smem + kFuseOffBias);
    uint4* swf = reinterpret_cast<uint4*>(smem + kFuseOffFrag);       // [4][32]
    float2* sbf = reinterpret_cast<float2*>(smem + kFuseOffFrag + 2048);   // [4][32]
    uint8_t* w1 = smem + kFuseOffW1;
    uint8_t* slots = smem + kFuseOffSlots;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int Ho = p.H >> 1, Wo = p.W >> 1;
    const int items = p.batch * p.tiles_w * p.segs;

    if (threadIdx.x == 0) {
        for (int i = 0; i < kFuseSlots; ++i) { mbar_init(&row_full[i], kFuseProducers); mbar_init(&row_empty[i], 1); }
        for (int i = 0; i < kFuseAcc; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        mbar_init(w_full, 1);
        fence_barrier_init();
    }
    if (warp == 4) { tmem_alloc(tmem_ptr, kFuseAcc * 64); tmem_relinquish(); }
    if (threadIdx.x < 64) sbias1[threadIdx.x] = __ldg(p.bias1 + threadIdx.x);
    if (warp >= 5 && warp < 9) {            // conv_0 weight / bias fragments, as stem_strip_kernel keeps them
        const int j = warp - 5, quad = lane & 3, grp = lane >> 2;
        uint4 f;
        const uint32_t* wr0 = reinterpret_cast<const uint32_t*>(p.w0 + (j * 8 + grp) * 32 + quad * 2);
        f.x = __ldg(wr0); f.y = __ldg(wr0 + 4); f.z = __ldg(wr0 + 8); f.w = __ldg(wr0 + 12);
        swf[j * 32 + lane] = f;
        sbf[j * 32 + lane] = make_float2(__ldg(p.bias0 + j * 8 + quad * 2), __ldg(p.bias0 + j * 8 + quad * 2 + 1));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 4) {
        // ===================== conv_1 weights (once) + MMA issuer =====================
        if (elect_one()) {
            tma_prefetch_desc(&tmap_w1);
            mbar_expect_tx(w_full, kFuseW1Bytes);
            for (int t = 0; t < 9; ++t) tma_load_2d(w1 + t * 4096, &tmap_w1, w_full, t * 32, 0);      // stored tap order: r, then s = 0, 2, 1
            mbar_wait(w_full, 0);
            constexpr uint32_t kIdesc = make_idesc_bf16(kBlockM, 64);
            const uint64_t desc_hi = make_smem_desc<32>(0);
            const uint32_t slots16 = (smem_u32(slots) & 0x3FFFF) >> 4, w116 = (smem_u32(w1) & 0x3FFFF) >> 4;
            long long rowc = 0;           // running index of the first conv_0 row of the current work item
            long long tilec = 0;          // running tile index (accumulator stage / phase)
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                const int seg = item % p.segs;
                const int y0 = seg * p.seg_rows, y1 = min(Ho, y0 + p.seg_rows);
                const int L = y1 - y0;
                for (int i = 0; i < L; ++i, ++tilec) {
                    const int acc = (int)(tilec % kFuseAcc);
                    mbar_wait(&tmem_empty[acc], (uint32_t)(((tilec / kFuseAcc) & 1) ^ 1));
                    const uint32_t tmem_d = tmem_base + acc * 64;
                    uint32_t accum = 0;
#pragma unroll
                    for (int r = 0; r < 3; ++r) {
                        const long long rc = rowc + 2 * i + r;
                        const int slot = (int)(rc % kFuseSlots);
                        mbar_wait_sleepy(&row_full[slot], (uint32_t)((rc / kFuseSlots) & 1));
                        tc_fence_after();
                        const uint32_t srow16 = slots16 + (uint32_t)slot * (kFuseSlotBytes >> 4);
#pragma unroll
                        for (int t = 0; t < 3; ++t) {                         // stored column-tap order s = 0, 2, 1
                            const int s = t == 0 ? 0 : (t == 1 ? 2 : 1);
                            const uint64_t da = desc_hi | (uint64_t)(srow16 + (uint32_t)(s & 1) * (kFuseArrayBytes >> 4) + (uint32_t)(s >> 1) * 4u);
                            const uint64_t db = desc_hi | (uint64_t)(w116 + (uint32_t)(r * 3 + t) * (4096u >> 4));
#pragma unroll
                            for (int k = 0; k < 2; ++k) {
                                umma_bf16(tmem_d, da + 2 * k, db + 2 * k, kIdesc, accum);
                                accum = 1;
                            }
                        }
                    }
                    umma_commit(&tmem_full[acc]);
                    // rows 2i and 2i + 1 of the item are not read again; the item's last tile also frees its third row
                    umma_commit(&row_empty[(int)((rowc + 2 * i) % kFuseSlots)]);
                    umma_commit(&row_empty[(int)((rowc + 2 * i + 1) % kFuseSlots)]);
                    if (i == L - 1) umma_commit(&row_empty[(int)((rowc + 2 * i + 2) % kFuseSlots)]);
                }
                rowc += 2 * L + 1;
            }
        }
    } else if (warp < 4) {
        // ===================== epilogue: TMEM -> bias -> LeakyReLU -> bf16 -> padded NHWC output =====================
        const int r = warp * 32 + lane;                  // row of the tile = TMEM lane
        long long tilec = 0;
        const uint32_t sb = smem_u32(sbias1);
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int seg = item % p.segs;
            const int tw = (item / p.segs) % p.tiles_w;
            const int n = item / (p.segs * p.tiles_w);
            const int y0 = seg * p.seg_rows, y1 = min(Ho, y0 + p.seg_rows);
            const int x0 = tw * p.tile_w, w = min(p.tile_w, Wo - x0);
            for (int y = y0; y < y1; ++y, ++tilec) {
                const int acc = (int)(tilec % kFuseAcc);
                mbar_wait_sleepy(&tmem_full[acc], (uint32_t)((tilec / kFuseAcc) & 1));
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + acc * 64;
                __nv_bfloat16* dst = p.out + (((size_t)n * (Ho + 2) + (y + 1)) * (Wo + 2) + (x0 + r + 1)) * 64;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t a32[32];
                    tmem_ld_32x32(taddr + c * 32, a32);
                    tmem_ld_wait();
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = lds128f(sb + (uint32_t)(c * 32 + 4 * j) * 4u);
                        v[4 * j + 0] = __uint_as_float(a32[4 * j + 0]); v[4 * j + 1] = __uint_as_float(a32[4 * j + 1]);
                        v[4 * j + 2] = __uint_as_float(a32[4 * j + 2]); v[4 * j + 3] = __uint_as_float(a32[4 * j + 3]);
                        add2(v[4 * j + 0], v[4 * j + 1], b.x, b.y);
                        add2(v[4 * j + 2], v[4 * j + 3], b.z, b.w);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float m0, m1;
                        mul2(m0, m1, v[j], v[j + 1], 0.1f);
                        v[j] = fmaxf(v[j], m0); v[j + 1] = fmaxf(v[j + 1], m1);
                    }
                    if (r < w) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 pk;
                            pk.x = pack_bf16x2(v[8 * j + 0], v[8 * j + 1]);
                            pk.y = pack_bf16x2(v[8 * j + 2], v[8 * j + 3]);
                            pk.z = pack_bf16x2(v[8 * j + 4], v[8 * j + 5]);
                            pk.w = pack_bf16x2(v[8 * j + 6], v[8 * j + 7]);
                            *reinterpret_cast<uint4*>(dst + c * 32 + j * 8) = pk;
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            }
        }
    } else {
        // ===================== producers: conv_0 rows into the row slots =====================
        // One strip per warp; every shared-memory access below uses explicit shared-space addresses computed once per work item
        // (generic loads / stores through pointers carved out of the aligned base cost three times the instructions).
        const int pw = warp - 5;                          // strip index 0 .. kFuseProducers - 1
        const int quad = lane & 3, grp = lane >> 2;
        const int par = grp & 1;                          // parity of this lane's pixels: which staged copy gives 4-byte aligned pairs
        const uint32_t stage_u32 = smem_u32(smem + kFuseOffStage + pw * kFuseStageBytes);      // [row & 3][copy][64] bf16
        const uint32_t slots_u32 = smem_u32(slots);
        const uint32_t swf_u32 = smem_u32(swf) + lane * 16, sbf_u32 = smem_u32(sbf) + lane * 8;
        // A fragments: this lane's four (k, k + 1) pairs of a pixel, k = ks * 16 + half * 8 + quad * 2 -> filter row k / 10, window element k % 10
        int frow[4]; uint32_t foff[4][2];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = (i >> 1) * 16 + (i & 1) * 8 + quad * 2;
            frow[i] = k < 30 ? k / 10 : 0;
            const int pjv = k < 30 ? k % 10 : 0;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) foff[i][rr] = (uint32_t)(par * kStripLen + pjv - par + (grp + rr * 8) * 3) * 2u;
        }
        long long rowc = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int seg = item % p.segs;
            const int tw = (item / p.segs) % p.tiles_w;
            const int n = item / (p.segs * p.tiles_w);
            const int y0 = seg * p.seg_rows, y1 = min(Ho, y0 + p.seg_rows);
            const int L = y1 - y0;
            const int x0 = tw * p.tile_w, w = min(p.tile_w, Wo - x0);
            const int npx = 2 * w + 1;                    // conv_0 pixels per row: padded columns 2 x0 .. 2 x0 + 2 w (q = 0 .. 2 w)
            const bool has = 16 * pw < npx;
            const T* base = img + (size_t)n * p.H * p.W * 3;
            const int cfirst = 2 * x0 - 1 + 16 * pw;      // conv_0 column of the strip's first pixel
            const int c0 = (cfirst - 1) * 3 + lane;       // staged element `lane` = image element (cfirst - 1) * 3 + lane of the row
            const bool ok0 = has && c0 >= 0 && c0 < p.W * 3;
            const bool ok1 = has && lane + 32 < 54 && c0 + 32 >= 0 && c0 + 32 < p.W * 3;
            // where this lane's eight results of a row go inside a row slot, and which of its two pixels are real / padding
            uint32_t soff[2][4]; bool live[2], zero[2];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int q = 16 * pw + grp + rr * 8;      // padded column 2 x0 + q
                const int cc = 2 * x0 - 1 + q;             // conv_0 column
                live[rr] = has && q < npx;
                zero[rr] = cc < 0 || cc >= p.W;            // conv_1's zero padding, not conv_0 of a padded image
                const int entry = q >> 1;
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    soff[rr][j] = (uint32_t)((q & 1) * kFuseArrayBytes + entry * 64 + ((j ^ ((entry >> 1) & 3)) << 4) + quad * 4);
            }
            auto fetch = [&](int hh, float& a, float& b) {
                const bool in = hh >= 0 && hh < p.H;
                const T* src = base + (size_t)(in ? hh : 0) * p.W * 3;
                a = (in && ok0) ? stem_px(__ldg(src + c0)) : 0.f;
                b = (in && ok1) ? stem_px(__ldg(src + c0 + 32)) : 0.f;
            };
            auto stage = [&](int hh, float a, float b) {
                const uint32_t E = stage_u32 + (uint32_t)((hh & 3) * 2 * kStripLen * 2), O = E + kStripLen * 2;
                const uint16_t x = __bfloat16_as_ushort(__float2bfloat16_rn(a)), y = __bfloat16_as_ushort(__float2bfloat16_rn(b));
                sts16(E + lane * 2, x); sts16(E + (lane + 32) * 2, y);
                if (lane >= 1) sts16(O + (lane - 1) * 2, x);
                sts16(O + (lane + 31) * 2, y);
            };
            // conv_0 rows of the item: padded rows 2 y0 .. 2 (y1 - 1) + 2, i.e. conv_0 rows ic = 2 y0 - 1 .. 2 y1 - 1
            const int ic0 = 2 * y0 - 1, nrows = 2 * L + 1;
            float pfa[kFusePf], pfb[kFusePf];             // image rows ic + 2 .. ic + 5 in flight: slot (k & 3) holds row ic0 + k + 2
            __syncwarp();
            if (has) {      // prime the staged ring: all loads go out before the first one is used
                float ra[3], rb[3];
#pragma unroll
                for (int d = 0; d < 3; ++d) fetch(ic0 - 1 + d, ra[d], rb[d]);
#pragma unroll
                for (int d = 0; d < kFusePf; ++d) fetch(ic0 + 2 + d, pfa[d], pfb[d]);
#pragma unroll
                for (int d = 0; d < 3; ++d) stage(ic0 - 1 + d, ra[d], rb[d]);
            }
            __syncwarp();
            auto do_row = [&](auto slot_c, int k) {
                constexpr int PS = decltype(slot_c)::value;
                const int ic = ic0 + k;
                const long long rc = rowc + k;
                const int slot = (int)(rc % kFuseSlots);
                mbar_wait(&row_empty[slot], (uint32_t)(((rc / kFuseSlots) & 1) ^ 1));
                if (has) {
                    const uint32_t slot_u32 = slots_u32 + (uint32_t)slot * kFuseSlotBytes;
                    const bool row_in = ic >= 0 && ic < p.H;
                    uint32_t afrag[2][4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t rowb = stage_u32 + (uint32_t)(((ic - 1 + frow[i]) & 3) * 2 * kStripLen * 2);
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) afrag[i >> 1][(i & 1) * 2 + rr] = lds_u32(rowb + foff[i][rr]);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 wf = lds128(swf_u32 + j * 512);
                        const float2 bj = lds64f(sbf_u32 + j * 256);
                        float acc[4] = {bj.x, bj.y, bj.x, bj.y};
                        mma_m16n8k16_bf16(acc, afrag[0], wf.x, wf.y);
                        mma_m16n8k16_bf16(acc, afrag[1], wf.z, wf.w);
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            float m0, m1;
                            mul2(m0, m1, acc[rr * 2 + 0], acc[rr * 2 + 1], 0.1f);        // LeakyReLU(0.1) = max(v, 0.1 v), as stem_strip_kernel
                            uint32_t val = pack_bf16x2(fmaxf(acc[rr * 2 + 0], m0), fmaxf(acc[rr * 2 + 1], m1));
                            if (!row_in || zero[rr]) val = 0u;
                            if (live[rr]) sts32(slot_u32 + soff[rr][j], (int)val);
                        }
                    }
                    fence_proxy_async();    // generic-proxy writes of the row -> visible to the tensor core's async-proxy reads
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&row_full[slot]);
                // Only now the next image row is requested: the fence above waits for the thread's outstanding memory operations, so
                // a load issued in front of it would expose its whole latency in every row.  Image row ic + 2 (loaded four rows ago)
                // replaces row ic - 2 in the staged ring; its register slot is refilled.
                if (has) {
                    stage(ic + 2, pfa[PS], pfb[PS]);
                    fetch(ic + 2 + kFusePf, pfa[PS], pfb[PS]);
                }
                __syncwarp();
            };
            for (int k = 0; k < nrows; k += kFusePf) {
                do_row(std::integral_constant<int, 0>{}, k);
                if (k + 1 < nrows) do_row(std::integral_constant<int, 1>{}, k + 1);
                if (k + 2 < nrows) do_row(std::integral_constant<int, 2>{}, k + 2);
                if (k + 3 < nrows) do_row(std::integral_constant<int, 3>{}, k + 3);
            }
            rowc += nrows;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, kFuseAcc * 64);
    }
}

}  // namespace fvy
