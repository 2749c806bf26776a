// Weight gradient of a stride-1 convolution on tcgen05 (see wgrad_kernel.cuh for the arithmetic and the reference
// lines).  dW_t[co][ci] = sum_p dY[p][co] * X[p + shift_t][ci] is a GEMM whose K dimension is the pixel index, and both operands are
// stored pixel-major ([pixel][channel], the forward path's padded bf16 activations): for tcgen05.mma that is the MN-MAJOR operand
// form.  A TMA box of 64 pixels x 64 channels (128-byte swizzle) lands in shared memory as eight 1024-byte atoms of 8 pixels x 64
// channels - exactly the canonical MN-major SW128 layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units (CUTLASS
// cute/atom/mma_traits_sm100.hpp): SBO = 1024 B between 8-pixel groups, LBO = 8192 B between 64-channel blocks (= one box), and a
// K = 16 step advances the start address by two atoms.  No transposes: dY and X are used as the pack kernel wrote them, and the
// tap shift is a row offset of the B boxes.
//
// One CTA = 192 threads: warp 0 TMA producer, warp 1 MMA issuer (one thread each), warps 2-5 epilogue.  Work item = (128 output
// channels, up to 256 input channels, tap, pixel range); accumulator 128 lanes x N columns in TMEM, double buffered so that the
// epilogue of an item (tcgen05.ld -> 16-byte red.global.add into dW_tap[co][ci]) runs under the MMAs of the next one.
#pragma once
#include "conv_igemm_sm100.cuh"

namespace fvy {

constexpr int kWtThreads = 192, kWtStages = 4, kWtKC = 64;
__host__ __device__ constexpr int wt_stage_bytes(int n) { return (128 + n) * kWtKC * 2; }
__host__ __device__ constexpr int wt_smem_bytes(int n) { return 1024 + kWtStages * wt_stage_bytes(n) + 1024; }

struct WgradTcParams {
    int cin, cout, taps;
    int tap_row[9];                  // row offset of tap t's X operand relative to dY's row (stride 1: (r-1) pitch + (s-1); stride 2: + its phase plane)
    int lead;                        // rows in front of pixel 0 in both buffers
    int chunks;                      // K chunks of 64 pixels
    int ksplit, n_tile;              // pixel ranges per (m, n, tap); input channels per item (64, 128 or 256)
    float* dw;                       // [taps][cout][cin] (the caller permutes to the torch layout; identical for 1 x 1)
};

// MN-major SW128 operand descriptor: start address, LBO = 8192 B (next 64-channel block), SBO = 1024 B (next 8 pixels)
__device__ __forceinline__ uint64_t wt_desc(uint32_t saddr) {
    return (uint64_t)((saddr & 0x3FFFF) >> 4) | ((uint64_t)(8192 >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(kWtThreads, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap map_dy, const __grid_constant__ CUtensorMap map_x, const WgradTcParams p) {
    extern __shared__ uint8_t wt_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(wt_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + kWtStages;
    uint64_t* tmem_full = empty + kWtStages;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint8_t* ring = smem + 1024;
    const int N = p.n_tile;
    const int stage_bytes = wt_stage_bytes(N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = p.cout / 128, n_tiles = p.cin / N;
    const int items = m_tiles * n_tiles * p.taps * p.ksplit;
    const int per = (p.chunks + p.ksplit - 1) / p.ksplit;
    const uint32_t tmem_cols = (uint32_t)(2 * N);

    if (warp == 0 && lane == 0) {
        for (int i = 0; i < kWtStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tmem_full[i], 1); mbar_init(&tmem_empty[i], 4); }
        fence_barrier_init();
        tma_prefetch_desc(&map_dy); tma_prefetch_desc(&map_x);
    }
    if (warp == 1) { tmem_alloc(tmem_ptr, tmem_cols); tmem_relinquish(); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    auto decode = [&](int item, int& mt, int& nt, int& tap, int& c0, int& c1) {
        const int ks = item % p.ksplit; int r = item / p.ksplit;
        tap = r % p.taps; r /= p.taps;
        nt = r % n_tiles; mt = r / n_tiles;
        c0 = ks * per; c1 = min(p.chunks, c0 + per);
    };

    if (warp == 0) {
        if (elect_one()) {
            int s = 0; uint32_t ph = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int mt, nt, tap, c0, c1;
                decode(item, mt, nt, tap, c0, c1);
                const int shift = p.tap_row[tap];
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = ring + (size_t)s * stage_bytes;
                    mbar_expect_tx(&full[s], (uint32_t)stage_bytes);
                    const int row = p.lead + c * kWtKC;
                    tma_load_2d(st, &map_dy, &full[s], mt * 128, row);
                    tma_load_2d(st + 8192, &map_dy, &full[s], mt * 128 + 64, row);
                    for (int b = 0; b < N / 64; ++b) tma_load_2d(st + 16384 + b * 8192, &map_x, &full[s], nt * N + b * 64, row + shift);
                    if (++s == kWtStages) { s = 0; ph ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, N) | (1u << 15) | (1u << 16);       // A and B MN-major
            int s = 0, acc = 0; uint32_t ph = 0, acc_ph = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x) {
                int mt, nt, tap, c0, c1;
                decode(item, mt, nt, tap, c0, c1);
                mbar_wait(&tmem_empty[acc], acc_ph ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc * N);
                uint32_t accum = 0;
                for (int c = c0; c < c1; ++c) {
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
                    const uint64_t da = wt_desc(sa), db = wt_desc(sa + 16384);
#pragma unroll
                    for (int k = 0; k < kWtKC / 16; ++k) {
                        umma_bf16(tmem_d, da + (uint64_t)(k * (2048 >> 4)), db + (uint64_t)(k * (2048 >> 4)), idesc, accum);
                        accum = 1;
                    }
                    umma_commit(&empty[s]);
                    if (++s == kWtStages) { s = 0; ph ^= 1; }
                }
                umma_commit(&tmem_full[acc]);
                if (++acc == 2) { acc = 0; acc_ph ^= 1; }
            }
        }
    } else {
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        const int r = q * 32 + lane;                  // output channel inside the item's 128
        int acc = 0; uint32_t acc_ph = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            int mt, nt, tap, c0, c1;
            decode(item, mt, nt, tap, c0, c1);
            mbar_wait(&tmem_full[acc], acc_ph);
            tc_fence_after();
            if (c1 > c0) {
                // this thread's output channel is one row of dW_tap: 32 consecutive input channels = 128 contiguous bytes, 16 per red
                float* dst = p.dw + ((long long)tap * p.cout + mt * 128 + r) * p.cin + nt * N;
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * N);
                for (int cc = 0; cc < N; cc += 32) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + cc, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        atomicAdd(reinterpret_cast<float4*>(dst + cc + j),
                                  make_float4(__uint_as_float(v[j]), __uint_as_float(v[j + 1]), __uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc]);
            if (++acc == 2) { acc = 0; acc_ph ^= 1; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, tmem_cols); }
}

// dW work layout [taps][cout][cin] -> torch layout [cout][cin][taps]
__global__ void __launch_bounds__(256) wgrad_permute_kernel(const float* __restrict__ work, int cout, int cin, int taps, float* __restrict__ dw) {
    const long long total = (long long)cout * cin * taps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int t = (int)(i % taps);
        const long long oc = i / taps;           // co * cin + ci
        dw[i] = work[(long long)t * cout * cin + oc];
    }
}

}  // namespace fvy
