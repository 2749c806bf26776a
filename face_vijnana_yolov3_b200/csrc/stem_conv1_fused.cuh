// conv_0 + conv_1 in ONE kernel (sm_100a): the stem's output never goes to HBM.
//
// Unfused, stem_strip_kernel writes conv_0's 32-channel activation (11.3 MB per 416x416 image: the largest tensor of the network,
// 443 MB at batch 40) as a 4-phase buffer and conv_1 (3x3, stride 2, 32 -> 64; yolov3_detect.py:221-223) reads it straight back:
// 158 us + 124 us of HBM-bound time for 7 % of the FLOPs.  Here a persistent CTA produces the conv_0 rows an output row of conv_1
// needs in SHARED memory, in exactly the form conv_1's tcgen05 MMA reads its A operand in, and consumes them from there:
//
//   producer warps (7)   conv_0 with warp-level bf16 MMAs (m16n8k16), the same staged-row scheme, K order and arithmetic as
//                        stem_strip_kernel (bit-identical values): every warp owns up to three 16-pixel strips of the tile's
//                        column range and walks down the rows, keeping a private ring of staged image rows; results are written
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

#include "conv_igemm_sm100.cuh"
#include "stem_kernel.cuh"

namespace fvy {

constexpr int kFuseProducers = 7;                  // producer warps: 14 strips of 16 pixels at 416 (two each)
constexpr int kFuseThreads = 32 * (4 + 1 + kFuseProducers);      // warps 0-3 epilogue, 4 MMA / TMEM / weights, 5.. producers
constexpr int kFuseSlots = 6;                      // conv_0 row slots
constexpr int kFuseAcc = 4;                        // accumulator stages of 64 TMEM columns
constexpr int kFuseMaxStrips = 3;                  // strips per producer warp (tile width <= 128 pixels -> <= 17 strips)
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
constexpr int kFuseStageBytes = kFuseMaxStrips * 4 * 2 * kStripLen * 2;    // per producer warp: strips x 4 image rows x 2 copies x 64 bf16
constexpr int kFuseSmem = kFuseOffStage + kFuseProducers * kFuseStageBytes + 1024;   // + alignment slack
static_assert(kFuseOffSlots % 1024 == 0 && kFuseSmem <= 232448, "fused stem: shared-memory plan");

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
                        mbar_wait(&row_full[slot], (uint32_t)((rc / kFuseSlots) & 1));
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
                mbar_wait(&tmem_full[acc], (uint32_t)((tilec / kFuseAcc) & 1));
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
        const int pw = warp - 5;                          // 0 .. kFuseProducers - 1
        const int quad = lane & 3, grp = lane >> 2;
        uint16_t* stage_base = reinterpret_cast<uint16_t*>(smem + kFuseOffStage + pw * kFuseStageBytes);     // [strip][row & 3][copy][64]
        int pr[4], pj[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = (i >> 1) * 16 + (i & 1) * 8 + quad * 2;
            pr[i] = k < 30 ? k / 10 : 0;
            pj[i] = k < 30 ? k % 10 : 0;
        }
        const int par = grp & 1;
        long long rowc = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int seg = item % p.segs;
            const int tw = (item / p.segs) % p.tiles_w;
            const int n = item / (p.segs * p.tiles_w);
            const int y0 = seg * p.seg_rows, y1 = min(Ho, y0 + p.seg_rows);
            const int L = y1 - y0;
            const int x0 = tw * p.tile_w, w = min(p.tile_w, Wo - x0);
            const int npx = 2 * w + 1;                    // conv_0 pixels per row: padded columns 2 x0 .. 2 x0 + 2 w (q = 0 .. 2 w)
            const int nstrips = (npx + 15) >> 4;
            const T* base = img + (size_t)n * p.H * p.W * 3;
            // this warp's strips: pw, pw + P, pw + 2 P
            int c0s[kFuseMaxStrips];
            bool ok0[kFuseMaxStrips], ok1[kFuseMaxStrips], has[kFuseMaxStrips];
#pragma unroll
            for (int u = 0; u < kFuseMaxStrips; ++u) {
                const int mt = pw + u * kFuseProducers;
                has[u] = mt < nstrips;
                const int cfirst = 2 * x0 - 1 + 16 * mt;                  // conv_0 column of the strip's first pixel
                c0s[u] = (cfirst - 1) * 3 + lane;                          // staged element `lane` = image element (cfirst - 1) * 3 + lane
                ok0[u] = has[u] && c0s[u] >= 0 && c0s[u] < p.W * 3;
                ok1[u] = has[u] && lane + 32 < 54 && c0s[u] + 32 >= 0 && c0s[u] + 32 < p.W * 3;
            }
            auto fetch = [&](int u, int hh, float& a, float& b) {
                const bool in = hh >= 0 && hh < p.H;
                const T* src = base + (size_t)(in ? hh : 0) * p.W * 3;
                a = (in && ok0[u]) ? stem_px(__ldg(src + c0s[u])) : 0.f;
                b = (in && ok1[u]) ? stem_px(__ldg(src + c0s[u] + 32)) : 0.f;
            };
            auto stage = [&](int u, int hh, float a, float b) {
                uint16_t* E = stage_base + ((u * 4 + (hh & 3)) * 2 + 0) * kStripLen;
                uint16_t* O = E + kStripLen;
                const uint16_t x = __bfloat16_as_ushort(__float2bfloat16_rn(a)), y = __bfloat16_as_ushort(__float2bfloat16_rn(b));
                E[lane] = x; E[lane + 32] = y;
                if (lane >= 1) O[lane - 1] = x;
                O[lane + 31] = y;
            };
            // conv_0 rows of the item: padded rows 2 y0 .. 2 (y1 - 1) + 2, i.e. conv_0 rows ic = 2 y0 - 1 .. 2 y1 - 1
            const int ic0 = 2 * y0 - 1, nrows = 2 * L + 1;
            float p0a[kFuseMaxStrips], p0b[kFuseMaxStrips], p1a[kFuseMaxStrips], p1b[kFuseMaxStrips];   // image rows ic + 2, ic + 3 in flight
            __syncwarp();
            {   // prime the staged ring: all loads go out before the first one is used
                float ra[kFuseMaxStrips][3], rb[kFuseMaxStrips][3];
#pragma unroll
                for (int u = 0; u < kFuseMaxStrips; ++u) {
                    if (!has[u]) continue;
#pragma unroll
                    for (int d = 0; d < 3; ++d) fetch(u, ic0 - 1 + d, ra[u][d], rb[u][d]);
                    fetch(u, ic0 + 2, p0a[u], p0b[u]);
                    fetch(u, ic0 + 3, p1a[u], p1b[u]);
                }
#pragma unroll
                for (int u = 0; u < kFuseMaxStrips; ++u) {
                    if (!has[u]) continue;
#pragma unroll
                    for (int d = 0; d < 3; ++d) stage(u, ic0 - 1 + d, ra[u][d], rb[u][d]);
                }
            }
            __syncwarp();
            for (int k = 0; k < nrows; ++k) {
                const int ic = ic0 + k;
                const long long rc = rowc + k;
                const int slot = (int)(rc % kFuseSlots);
                mbar_wait(&row_empty[slot], (uint32_t)(((rc / kFuseSlots) & 1) ^ 1));
                uint8_t* arrE = slots + slot * kFuseSlotBytes;
                const bool row_in = ic >= 0 && ic < p.H;                  // outside: conv_1's zero padding, not conv_0 of a padded image
#pragma unroll
                for (int u = 0; u < kFuseMaxStrips; ++u) {
                    if (!has[u]) continue;
                    const int mt = pw + u * kFuseProducers;
                    uint32_t afrag[2][4];
                    const uint16_t* sb = stage_base + (u * 4) * 2 * kStripLen;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint16_t* rp = sb + ((((ic - 1 + pr[i]) & 3) * 2 + par) * kStripLen) + pj[i] - par;
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr)
                            afrag[i >> 1][(i & 1) * 2 + rr] = *reinterpret_cast<const uint32_t*>(rp + (grp + rr * 8) * 3);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 wf = swf[j * 32 + lane];
                        const float2 bj = sbf[j * 32 + lane];
                        float acc[4] = {bj.x, bj.y, bj.x, bj.y};
                        mma_m16n8k16_bf16(acc, afrag[0], wf.x, wf.y);
                        mma_m16n8k16_bf16(acc, afrag[1], wf.z, wf.w);
#pragma unroll
                        for (int rr = 0; rr < 2; ++rr) {
                            const int q = 16 * mt + grp + rr * 8;           // padded column 2 x0 + q of the conv_0 row
                            if (q >= npx) continue;
                            const int cc = 2 * x0 - 1 + q;                  // conv_0 column
                            float a = acc[rr * 2 + 0], b = acc[rr * 2 + 1];
                            a = fmaxf(a, 0.1f * a); b = fmaxf(b, 0.1f * b);
                            __nv_bfloat162 pk = __floats2bfloat162_rn(a, b);
                            uint32_t val = *reinterpret_cast<uint32_t*>(&pk);
                            if (!row_in || cc < 0 || cc >= p.W) val = 0u;
                            const int entry = q >> 1;
                            uint8_t* arr = arrE + (q & 1) * kFuseArrayBytes;
                            *reinterpret_cast<uint32_t*>(arr + entry * 64 + ((j ^ ((entry >> 1) & 3)) << 4) + quad * 4) = val;
                        }
                    }
                }
                // image row ic + 2 (loaded two iterations ago) replaces row ic - 2 in the staged ring; the next load goes out
                __syncwarp();
#pragma unroll
                for (int u = 0; u < kFuseMaxStrips; ++u) {
                    if (!has[u]) continue;
                    stage(u, ic + 2, p0a[u], p0b[u]);
                    p0a[u] = p1a[u]; p0b[u] = p1b[u];
                    fetch(u, ic + 4, p1a[u], p1b[u]);
                }
                fence_proxy_async();        // generic-proxy writes of the row -> visible to the tensor core's async-proxy reads
                __syncwarp();
                if (lane == 0) mbar_arrive(&row_full[slot]);
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
