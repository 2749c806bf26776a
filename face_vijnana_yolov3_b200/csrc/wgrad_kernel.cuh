// Weight gradient of a stride-1, zero-padded k x k convolution (k = 1 or 3) - SURVEY 8 row f-1: what Keras / TensorFlow hand to
// cuDNN for every Conv2D when the reference trains (src/space/yolov3_detect.py:206-211 via src/space/face_detection.py:361-381).
//
//   dW[co][ci][r][s] = sum over pixels p of dY[p][co] * X[p + (r - k/2) * pitch + (s - k/2)][ci]
//
// with X and dY stored as bf16 pixel-major matrices over the SHARED-HALO geometry of the forward path (row pitch W + 1, image
// pitch (H + 1)(W + 1), halo pixels zero): every tap is a constant row shift, halo pixels contribute zeros, so the sum simply runs
// over all rows.  It is a GEMM whose K dimension is the pixel index and whose two operands are both stored K-major-in-rows
// ([pixel][channel]): exactly what ldmatrix.trans feeds to warp-level mma.sync.m16n8k16 (bf16 in, fp32 accumulate).  This first
// version therefore uses the warp-level tensor-core path; the tcgen05 form needs MN-major shared-memory descriptors for both
// operands and is the follow-up.
//
// Block = 256 threads, tile = 64 output channels x 64 input channels x ALL taps (9 x 16 accumulator registers per thread for a
// 3 x 3 filter): one dY tile of 32 pixels serves the nine taps, X comes as three 34-row slabs (one per filter row; the column
// taps are the slab read one / two rows further).  Three-stage cp.async pipeline, 16-byte chunks XOR-swizzled by the row so that
// ldmatrix is conflict-free.  The pixel range is split over gridDim.z; partial sums are added to dW with red.global.add.f32
// (dW is zeroed by the caller), in the torch layout [Co][Ci][k][k].
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace fvy {

constexpr int kWgThreads = 256, kWgKC = 32, kWgStages = 3;
template <int TAPS> struct WgSmem {
    static constexpr int kSlabs = TAPS == 9 ? 3 : 1;
    static constexpr int kSlabRows = TAPS == 9 ? kWgKC + 2 : kWgKC;
    static constexpr int kStageBytes = (kWgKC + kSlabs * kSlabRows) * 128;
    static constexpr int kBytes = kWgStages * kStageBytes;
};

__device__ __forceinline__ void wg_cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void wg_cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void wg_cp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void wg_ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void wg_mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// byte offset of 16-byte chunk `chunk` (0..7) of row `row` of a [rows][64 bf16] tile: chunks XOR-swizzled by the row
__device__ __forceinline__ uint32_t wg_off(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

// x, dy: row 0 of each pointer = pixel index 0 of the shared-halo geometry; rows [-(pitch + 1), rows_total + pitch + 1) are readable
// (zero margins).  rows_k: pixels to sum over, a multiple of kWgKC.
template <int TAPS>
__global__ void __launch_bounds__(kWgThreads, 1) conv_wgrad_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                                                   int rows_k, int pitch, int cin, int cout, float* __restrict__ dw) {
    using S = WgSmem<TAPS>;
    extern __shared__ __align__(128) uint8_t wg_smem[];
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(wg_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                 // warp tile: 32 output channels (wm) x 16 input channels (wn)
    const int co0 = blockIdx.x * 64, ci0 = blockIdx.y * 64;
    // this block's pixel range
    const int chunks = rows_k / kWgKC;
    const int per = (chunks + gridDim.z - 1) / gridDim.z;
    const int c_begin = blockIdx.z * per, c_end = min(chunks, c_begin + per);
    if (c_begin >= c_end) return;

    float acc[TAPS][2][2][4];
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[t][i][j][e] = 0.f;

    auto load_stage = [&](int chunk, int stage) {
        const uint32_t sb = smem0 + (uint32_t)stage * S::kStageBytes;
        const long long p0 = (long long)chunk * kWgKC;
        {   // dY tile: 32 rows x 8 chunks = 256 copies
            const int row = tid >> 3, ch = tid & 7;
            wg_cp_async16(sb + wg_off(row, ch), dy + (p0 + row) * cout + co0 + ch * 8);
        }
        constexpr int kCopies = S::kSlabs * S::kSlabRows * 8;
        for (int i = tid; i < kCopies; i += kWgThreads) {
            const int ch = i & 7, rr = i >> 3;
            const int slab = rr / S::kSlabRows, row = rr - slab * S::kSlabRows;
            const long long src_row = TAPS == 9 ? p0 - 1 + row + (long long)(slab - 1) * pitch : p0 + row;
            wg_cp_async16(sb + kWgKC * 128 + (uint32_t)slab * S::kSlabRows * 128 + wg_off(row, ch), x + src_row * cin + ci0 + ch * 8);
        }
    };

    // prologue
#pragma unroll
    for (int s = 0; s < kWgStages - 1; ++s) {
        if (c_begin + s < c_end) load_stage(c_begin + s, s);
        wg_cp_commit();
    }
    for (int c = c_begin; c < c_end; ++c) {
        const int stage = (c - c_begin) % kWgStages;
        wg_cp_wait<kWgStages - 2>();
        __syncthreads();
        {   // prefetch the chunk kWgStages - 1 ahead into the stage that was consumed in the previous iteration
            const int nc = c + kWgStages - 1;
            if (nc < c_end) load_stage(nc, (nc - c_begin) % kWgStages);
            wg_cp_commit();
        }
        const uint32_t sb = smem0 + (uint32_t)stage * S::kStageBytes;
#pragma unroll
        for (int k16 = 0; k16 < kWgKC / 16; ++k16) {
            // A fragments: dY^T, 32 output channels x 16 pixels (two m16 tiles)
            uint32_t a[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int mat = lane >> 3;
                const int krow = k16 * 16 + (mat >> 1) * 8 + (lane & 7);
                const int mcol = wm * 32 + i * 16 + (mat & 1) * 8;
                wg_ldmatrix_x4_trans(sb + wg_off(krow, mcol >> 3), a[i][0], a[i][1], a[i][2], a[i][3]);
            }
#pragma unroll
            for (int t = 0; t < TAPS; ++t) {
                const int slab = TAPS == 9 ? t / 3 : 0, shift = TAPS == 9 ? t % 3 : 0;
                // B fragments: X, 16 pixels x 16 input channels (two n8 tiles): matrices (k 0-7, n 0-7), (k 8-15, n 0-7), (k 0-7, n 8-15), (k 8-15, n 8-15)
                uint32_t b[4];
                const int mat = lane >> 3;
                const int krow = k16 * 16 + (mat & 1) * 8 + (lane & 7) + shift;
                const int ncol = wn * 16 + (mat >> 1) * 8;
                wg_ldmatrix_x4_trans(sb + kWgKC * 128 + (uint32_t)slab * S::kSlabRows * 128 + wg_off(krow, ncol >> 3), b[0], b[1], b[2], b[3]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    wg_mma_bf16(acc[t][i][0], a[i][0], a[i][1], a[i][2], a[i][3], b[0], b[1]);
                    wg_mma_bf16(acc[t][i][1], a[i][0], a[i][1], a[i][2], a[i][3], b[2], b[3]);
                }
            }
        }
    }
    wg_cp_wait<0>();
    // partial sums -> dW[co][ci][tap] (torch layout: taps contiguous)
    const int g = lane >> 2, tq = lane & 3;
#pragma unroll
    for (int t = 0; t < TAPS; ++t)
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    const int co = co0 + wm * 32 + i * 16 + g + (e >> 1) * 8;
                    const int ci = ci0 + wn * 16 + j * 8 + tq * 2 + (e & 1);
                    atomicAdd(dw + ((long long)co * cin + ci) * TAPS + t, acc[t][i][j][e]);
                }
}

}  // namespace fvy
