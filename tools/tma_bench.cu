// Micro-benchmark: per-SM and chip-wide TMA load throughput / latency as a function of box shape, row pitch and loads in
// flight.  One thread per CTA keeps S bulk-tensor loads in flight into a shared-memory ring and re-issues each as soon as it
// lands (no consumer), so the measured bytes/clk is what the TMA -> L2 path can deliver to one SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/tma_bench tools/tma_bench.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c)); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0;
    while (!done)
        asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}

// mode 0: every CTA walks its own region; mode 1: CTAs c and c^1 read the same boxes (pair sharing)
__global__ void __launch_bounds__(128, 1) tma_kernel(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int S, int iters,
                                                      int rows_total, int cols_boxes, int share, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + 1024;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int cta = share ? (blockIdx.x >> (share - 1)) : blockIdx.x;
        const int boxes_per_col = rows_total / box_rows;
        unsigned idx = (unsigned)cta * 977u;             // scatter the start positions
        // lean issue loop: no divisions, running slot / phase / coordinates
        const int c0 = (int)((idx / boxes_per_col) % cols_boxes) * 64;
        int c1 = (int)(idx % boxes_per_col) * box_rows;
        const int c1_end = boxes_per_col * box_rows;
        long long t0 = clock64();
        long long lat_sum = 0;
        int s = 0; uint32_t ph = 0;
#pragma unroll 1
        for (int i = 0; i < iters + S; ++i) {
            if (i >= S) mbar_wait(&bars[s], ph ^ 1);
            if (i < iters) {
                mbar_expect(&bars[s], box_bytes);
                tma_load_2d(ring + (size_t)s * box_bytes, &map, &bars[s], c0, c1);
                c1 += box_rows; if (c1 >= c1_end) c1 = 0;
            }
            if (++s == S) { s = 0; ph ^= 1; }
        }
        long long t1 = clock64();
        // single-load latency, idle pipe
        for (int i = 0; i < 8; ++i) {
            const unsigned b = (idx + 13 * i) % (unsigned)(boxes_per_col * cols_boxes);
            long long a = clock64();
            mbar_expect(&bars[0], box_bytes);
            tma_load_2d(ring, &map, &bars[0], (int)(b / boxes_per_col) * 64, (int)(b % boxes_per_col) * box_rows);
            mbar_wait(&bars[0], (((iters + S - 1) / S) + i) & 1);
            lat_sum += clock64() - a;
        }
        out[blockIdx.x * 2] = (unsigned long long)(t1 - t0);
        out[blockIdx.x * 2 + 1] = (unsigned long long)(lat_sum / 8);
    }
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
// variant: W warps issue (warp w owns slots w, w+W, ...); uniform = 1: the whole warp runs the loop and one elected lane issues
__global__ void __launch_bounds__(128, 1) tma_kernel2(const __grid_constant__ CUtensorMap map, int box_bytes, int box_rows, int S, int iters,
                                                       int rows_total, int W, int uniform, unsigned long long* out) {
    extern __shared__ uint8_t raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);
    uint8_t* ring = smem + 1024;
    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp >= W) return;
    if (!uniform && lane != 0) return;
    const int boxes_per_col = rows_total / box_rows;
    const int c1_end = boxes_per_col * box_rows;
    int c1 = (int)(((unsigned)blockIdx.x * 977u + warp * 131u) % boxes_per_col) * box_rows;
    const int my_slots = (S - warp + W - 1) / W;
    const int my_iters = iters / W;
    long long t0 = clock64();
    int k = 0; uint32_t ph = 0;
#pragma unroll 1
    for (int i = 0; i < my_iters + my_slots; ++i) {
        const int s = warp + k * W;
        if (i >= my_slots) mbar_wait(&bars[s], ph ^ 1);
        if (i < my_iters) {
            if (!uniform || elect_one()) {
                mbar_expect(&bars[s], box_bytes);
                tma_load_2d(ring + (size_t)s * box_bytes, &map, &bars[s], 0, c1);
            }
            c1 += box_rows; if (c1 >= c1_end) c1 = 0;
        }
        if (++k == my_slots) { k = 0; ph ^= 1; }
    }
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x * 8 + warp] = (unsigned long long)(t1 - t0);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    void* fnp = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fnp;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("SMs %d, max clock %d MHz\n", sms, clk_khz / 1000);
    const size_t buf_bytes = 64ull << 20;    // L2 resident (126 MB L2)
    void* buf;
    CK(cudaMalloc(&buf, buf_bytes));
    CK(cudaMemset(buf, 1, buf_bytes));
    unsigned long long* out;
    CK(cudaMalloc(&out, sms * 16));
    CK(cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    printf("%8s %8s %4s %6s %6s | %10s %10s %10s %8s\n", "pitchB", "boxrows", "S", "CTAs", "share", "B/clk/SM", "GB/s/SM", "TB/s chip", "lat clk");
    struct Cfg { int pitch, rows, S, ctas, share, l2promo; };
    std::vector<Cfg> cfgs;
    for (int pitch : {128, 512, 1024, 4608})
        for (int rows : {128})
            for (int S : {1, 2, 4, 8, 12})
                for (int ctas : {1, sms}) cfgs.push_back({pitch, rows, S, ctas, 0, 2});
    for (int rows : {32, 64, 136, 256})
        for (int S : {4, 8}) cfgs.push_back({512, rows, S, sms, 0, 2});
    for (int share : {2, 3, 4})               // 2, 4, 8 CTAs read the same boxes
        for (int S : {8}) cfgs.push_back({512, 128, S, sms, share, 2});
    for (int promo : {0, 1, 3})
        cfgs.push_back({512, 128, 8, sms, 0, promo});
    for (const Cfg& c : cfgs) {
        const int box_bytes = c.rows * 128;
        if ((size_t)c.S * box_bytes + 2048 > 232448) continue;
        const uint64_t cols = c.pitch / 2;                      // bf16 elements per row
        const uint64_t rows_total = buf_bytes / c.pitch;
        CUtensorMap map;
        cuuint64_t dims[2] = {cols, rows_total};
        cuuint64_t strides[1] = {(cuuint64_t)c.pitch};
        cuuint32_t box[2] = {64, (cuuint32_t)c.rows};
        cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         (CUtensorMapL2promotion)c.l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); continue; }
        const int iters = 2000;
        const size_t smem = 2048 + (size_t)c.S * box_bytes;
        for (int rep = 0; rep < 2; ++rep) {
            tma_kernel<<<c.ctas, 128, smem>>>(map, box_bytes, c.rows, c.S, iters, (int)rows_total, (int)(cols / 64), c.share, out);
            CK(cudaDeviceSynchronize());
        }
        std::vector<unsigned long long> h(c.ctas * 2);
        CK(cudaMemcpy(h.data(), out, c.ctas * 16, cudaMemcpyDeviceToHost));
        double clk = 0, lat = 0;
        for (int i = 0; i < c.ctas; ++i) { clk += (double)h[2 * i]; lat += (double)h[2 * i + 1]; }
        clk /= c.ctas; lat /= c.ctas;
        const double bpc = (double)iters * box_bytes / clk;
        printf("%8d %8d %4d %6d %6d | %10.1f %10.1f %10.2f %8.0f   promo=%d\n", c.pitch, c.rows, c.S, c.ctas, c.share, bpc, bpc * clk_khz / 1e6,
               bpc * clk_khz / 1e6 * c.ctas / 1e3, lat, c.l2promo);
    }
    // ---- issue-rate variants
    CK(cudaFuncSetAttribute(tma_kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448));
    unsigned long long* out2;
    CK(cudaMalloc(&out2, sms * 64));
    printf("\nissue variants (pitch 512, all SMs): boxrows S W uniform -> B/clk/SM, clk per TMA instruction per warp\n");
    for (int rows : {32, 64, 128})
        for (int W : {1, 2, 4})
            for (int uniform : {0, 1}) {
                const int S = 8, pitch = 512, box_bytes = rows * 128, iters = 2000;
                const uint64_t rows_total = buf_bytes / pitch;
                CUtensorMap map;
                cuuint64_t dims[2] = {(cuuint64_t)pitch / 2, rows_total};
                cuuint64_t strides[1] = {(cuuint64_t)pitch};
                cuuint32_t box[2] = {64, (cuuint32_t)rows};
                cuuint32_t es[2] = {1, 1};
                if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) continue;
                for (int rep = 0; rep < 2; ++rep) {
                    tma_kernel2<<<sms, 128, 2048 + S * box_bytes>>>(map, box_bytes, rows, S, iters, (int)rows_total, W, uniform, out2);
                    CK(cudaDeviceSynchronize());
                }
                std::vector<unsigned long long> h(sms * 8);
                CK(cudaMemcpy(h.data(), out2, sms * 64, cudaMemcpyDeviceToHost));
                double clk = 0;
                for (int i = 0; i < sms; ++i) { double m = 0; for (int w = 0; w < W; ++w) m = std::max(m, (double)h[i * 8 + w]); clk += m; }
                clk /= sms;
                printf("%6d %3d %2d %2d -> %8.1f B/clk/SM  %7.0f clk/instr\n", rows, S, W, uniform, (double)iters * box_bytes / clk, clk / (iters / W));
            }
    return 0;
}
